"""Back-to-back ppo_gae launches through the C ABI on rotating buffers (> L2), per-launch device time (dev aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import _lib

def run(steps, envs, strided, stats, sets=8, iters=64):
    lib = _lib.load()
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    bufs = []
    for _ in range(sets):
        r = torch.randint(-1, 3, (steps, envs), device=dev, generator=g).float()
        d = (torch.rand((steps, envs), device=dev, generator=g) < 0.03).view(torch.uint8)
        if strided:
            vs = torch.randn((steps + 1, envs, 128), device=dev, generator=g)
            v = vs[:, :, -1]
        else:
            v = torch.randn((steps + 1, envs), device=dev, generator=g)
        bufs.append((r, d, v, torch.empty((steps, envs), device=dev), torch.empty((steps, envs), device=dev)))
    st = torch.zeros(16, dtype=torch.float64, device=dev)
    ws = torch.zeros(int(lib.ppo_gae_workspace_bytes(envs)), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    def launch(i):
        r, d, v, a, t = bufs[i % sets]
        rc = lib.ppo_gae(r.data_ptr(), envs, v.data_ptr(), v.stride(0), v.stride(1), v[steps].data_ptr(), v.stride(1),
                         d.data_ptr(), 1, envs, steps, envs, 0.999, 0.95, 0.95, a.data_ptr(), t.data_ptr(),
                         st.data_ptr() if stats else 0, ws.data_ptr(), ws.numel(), s)
        _lib.check(rc, "gae")
    for i in range(8): launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): launch(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    nbytes = steps * envs * 21
    print(f"gae {steps}x{envs} strided_v={int(strided)} stats={int(stats)} tile={os.environ.get('PPO_GAE_TILE','1')}: "
          f"{us:7.1f} us/launch  {nbytes / us / 1e3:6.0f} GB/s algorithmic  frac {nbytes / us / 1e3 / 6548.8:.3f}")

if __name__ == "__main__":
    for envs, sets in ((4096, 8), (32768, 2)):
        for strided in (False, True):
            if strided and envs > 4096:
                continue
            for stats in (False, True):
                run(256, envs, strided, stats, sets=sets)
