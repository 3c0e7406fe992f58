"""Dev aid: ppo_gae (no stats / stats) and ppo_gae_normalized at 256 x 4096, back-to-back launches on rotating buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import _lib

def main(envs=4096, steps=256, sets=8, iters=100):
    lib = _lib.load()
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    bufs = []
    for _ in range(sets):
        r = torch.randint(-1, 3, (steps, envs), device=dev, generator=g).float()
        d = (torch.rand((steps, envs), device=dev, generator=g) < 0.03).view(torch.uint8)
        v = torch.randn((steps + 1, envs), device=dev, generator=g)
        bufs.append((r, d, v, [torch.empty((steps, envs), device=dev) for _ in range(3)]))
    st = torch.zeros(16, dtype=torch.float64, device=dev)
    ws = torch.zeros(int(lib.ppo_gae_workspace_bytes(envs)), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def gae(i, stats):
        r, d, v, o = bufs[i % sets]
        _lib.check(lib.ppo_gae(r.data_ptr(), envs, v.data_ptr(), envs, 1, v[steps].data_ptr(), 1, d.data_ptr(), 1, envs,
                               steps, envs, 0.999, 0.95, 0.95, o[0].data_ptr(), o[1].data_ptr(),
                               st.data_ptr() if stats else 0, ws.data_ptr(), ws.numel(), s), "gae")

    def fused(i):
        r, d, v, o = bufs[i % sets]
        rc = lib.ppo_gae_normalized(r.data_ptr(), envs, v.data_ptr(), envs, 1, v[steps].data_ptr(), 1, d.data_ptr(), 1,
                                    envs, steps, envs, 0.999, 0.95, 0.95, 1e-8, 0.0, o[0].data_ptr(), o[1].data_ptr(),
                                    o[2].data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), s)
        assert rc >= 0

    def norm(i):
        r, d, v, o = bufs[i % sets]
        _lib.check(lib.ppo_adv_normalize(o[0].data_ptr(), steps * envs, st.data_ptr(), 1e-8, 0.0, o[2].data_ptr(), s), "n")

    def timed(name, fn):
        for i in range(8): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:40s} {e0.elapsed_time(e1) * 1e3 / iters:7.2f} us", flush=True)

    for ew in ("16", "32"):
        os.environ["PPO_GAE_EW"] = ew
        timed(f"EW={ew} ppo_gae, no stats", lambda i: gae(i, False))
        timed(f"EW={ew} ppo_gae, stats", lambda i: gae(i, True))
        timed(f"EW={ew} ppo_gae_normalized (fused)", fused)
    timed("ppo_adv_normalize", norm)
    def empty(i):
        torch.cuda._sleep(0) if False else None
    # launch floor: an empty-ish kernel (ppo_adv_normalize on 1 element)
    def tiny(i):
        r, d, v, o = bufs[i % sets]
        lib.ppo_adv_normalize(o[0].data_ptr(), 1, st.data_ptr(), 1e-8, 0.0, o[2].data_ptr(), s)
    timed("1-element kernel (launch floor)", tiny)

if __name__ == "__main__":
    main()
