"""tslab kernel time against the number of envs per SM (fixed cost vs per-env cost), development aid."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_b200 import tvf as _tvf_mod
from ppo_b200 import returns_truncated as rt, _lib

def main():
    lib = _lib.load()
    hz = _tvf_mod.get_value_head_horizons(128, 30000)
    device = torch.device("cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for steps in (128, 256):
        table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
        tabs = rt.build_tables(0.99997, steps, hz, hz, table)
        ctab, blob = rt._upload_tables(tabs, device)
        for per_sm in (1, 2, 4, 8, 16):
            envs = 148 * per_sm
            g = torch.Generator(device="cuda"); g.manual_seed(0)
            r = torch.randint(-1, 3, (steps, envs), device=device, generator=g).float()
            d = (torch.rand((steps, envs), device=device, generator=g) < 0.03).view(torch.uint8)
            vs = torch.randn((steps + 1, envs, 128), device=device, generator=g) * 0.4 + 0.1
            out = torch.empty((steps, envs, 128), device=device)
            def fn():
                rc = lib.ppo_tvf_targets(r.data_ptr(), envs, d.data_ptr(), 1, envs, vs.data_ptr(), envs * 128, 128,
                                         steps, envs, 128, 128, ctypes.byref(ctab), rt.PATH_TSLAB, out.data_ptr(), 0, 0, 0,
                                         torch.cuda.current_stream().cuda_stream)
                _lib.check(rc, "tvf")
            for _ in range(3): fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); fn(); e.record(); torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3)
            ts.sort()
            print(f"N={steps} envs/SM={per_sm:2d}: {ts[len(ts)//2]:8.1f} us  ({ts[len(ts)//2]/per_sm:6.1f} us per env)")

if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def stamps():
    """PPO_TVF_DEBUG=64: print the hand-off time line of CTA 0 for one env per SM and eight envs per SM."""
    os.environ["PPO_TVF_DEBUG"] = "64"
    lib = _lib.load()
    import ctypes as C
    hz = _tvf_mod.get_value_head_horizons(128, 30000)
    device = torch.device("cuda")
    names = ["start", "smem init", "R/lim ready", "chunk 0 in slab", "env 0 in slab", "consumers start", "env 0 done",
             "last step starts", "consumers done", "last store done"]
    for steps, per_sm in ((128, 1), (128, 8), (256, 1), (256, 8)):
        envs = 148 * per_sm
        table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
        tabs = rt.build_tables(0.99997, steps, hz, hz, table)
        ctab, blob = rt._upload_tables(tabs, device)
        r = torch.randint(-1, 3, (steps, envs), device=device).float()
        d = (torch.rand((steps, envs), device=device) < 0.03).view(torch.uint8)
        vs = torch.randn((steps + 1, envs, 128), device=device)
        out = torch.empty((steps, envs, 128), device=device)
        for _ in range(3):
            lib.ppo_tvf_targets(r.data_ptr(), envs, d.data_ptr(), 1, envs, vs.data_ptr(), envs * 128, 128, steps, envs,
                                128, 128, ctypes.byref(ctab), rt.PATH_TSLAB, out.data_ptr(), 0, 0, 0,
                                torch.cuda.current_stream().cuda_stream)
        buf = (C.c_ulonglong * 16)()
        lib.ppo_tvf_debug_times(buf)
        t0 = buf[0]
        print(f"N={steps} envs/SM={per_sm}: " + ", ".join(f"{n} {(buf[i] - t0) / 1e3:.1f}" for i, n in enumerate(names)))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "stamps":
    stamps()
