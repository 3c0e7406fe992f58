"""Development aid: device time of the transposed slab kernel (or, with PPO_TIME_PATH=qslab, the quad slab kernel) at
the bench shape (256 x 4096 x 128, exponential / advanced 8 samples) under the kernel's PPO_TVF_DEBUG bits (read per
launch).  PPO_TIME_CHECK=1 also compares the targets of the two kernels."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import tvf as _tvf_mod
from ppo_b200 import returns_truncated as rt, _lib


def main():
    steps, envs = [int(x) for x in os.environ.get("PPO_TIME_SHAPE", "256,4096").split(",")]
    lib = _lib.load()
    hz = _tvf_mod.get_value_head_horizons(128, 30000)
    device = torch.device("cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    r = torch.randint(-1, 3, (steps, envs), device=device, generator=g).float()
    d = (torch.rand((steps, envs), device=device, generator=g) < 0.03).view(torch.uint8)
    off = int(os.environ.get("PPO_TIME_OFFSET", "0")) // 4  # start the two big buffers this many bytes into their allocations
    vs_all = torch.empty((steps + 1) * envs * 128 + off, device=device)
    vs = vs_all[off:].view(steps + 1, envs, 128)
    vs.copy_(torch.randn((steps + 1, envs, 128), device=device, generator=g) * 0.4 + 0.1)
    out_all = torch.empty(steps * envs * 128 + off, device=device)
    out = out_all[off:].view(steps, envs, 128)
    print("alignment of the buffers within a 2 MB page:", vs.data_ptr() % (1 << 21), out.data_ptr() % (1 << 21))
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    tabs = rt.build_tables(0.99997, steps, hz, hz, table)
    ctab, blob = rt._upload_tables(tabs, device)
    s = torch.cuda.current_stream().cuda_stream

    path = {"tslab": rt.PATH_TSLAB, "qslab": rt.PATH_QSLAB, "cslab": rt.PATH_CSLAB}[os.environ.get("PPO_TIME_PATH", "tslab")]

    def launch(path=path, out=out):
        rc = lib.ppo_tvf_targets(r.data_ptr(), envs, d.data_ptr(), 1, envs, vs.data_ptr(), envs * 128, 128,
                                 steps, envs, 128, 128, ctypes.byref(ctab), path, out.data_ptr(), 0, 0, 0, s)
        _lib.check(rc, "tvf")
    if os.environ.get("PPO_TIME_CHECK"):
        ref = torch.empty_like(out)
        launch(rt.PATH_TSLAB, ref)
        launch()
        torch.cuda.synchronize()
        print("max |diff| vs tslab:", float((out - ref).abs().max()), "equal:", bool(torch.equal(out, ref)))
    for dbg in [int(x) for x in (sys.argv[1:] or ["0"])]:
        os.environ["PPO_TVF_DEBUG"] = str(dbg)
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            launch()
        e1.record()
        torch.cuda.synchronize()
        print(f"dbg {dbg:5d}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us")


if __name__ == "__main__":
    main()
