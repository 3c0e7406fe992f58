"""Scratch micro-benchmark of the TVF target kernels (development aid, not the judged bench)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_b200 import returns_truncated as rt, synthetic, _lib, _device as dev, _tvf_tables

def heads():
    from ppo_b200 import tvf as _tvf_mod
    return _tvf_mod.get_value_head_horizons(128, 30000)

def main():
    lib = _lib.load()
    hz = heads()
    device = torch.device("cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    cfgs = [(128, 1024), (256, 4096)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cfgs.append((256, 32768))
    for steps, envs in cfgs:
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        r = torch.randint(-1, 3, (steps, envs), device=device, generator=g).float()
        for done_p in (0.03, 0.0):
            d = (torch.rand((steps, envs), device=device, generator=g) < done_p).view(torch.uint8)
            vs = torch.randn((steps + 1, envs, 128), device=device, generator=g) * 0.4 + 0.1
            vs[:, :, 0] = 0
            out = torch.empty((steps, envs, 128), device=device)
            for mode in ("advanced8", "fixed20", "standard8"):
                pass
                if mode == "advanced8":
                    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
                elif mode == "standard8":
                    table = rt.draw_n_step_samples("exponential", "standard", hz, steps, 20, 8, seed=1)
                else:
                    table = np.zeros([128, 1], dtype=np.int32) + 20
                tabs = rt.build_tables(0.99997, steps, hz, hz, table)
                ctab, blob = rt._upload_tables(tabs, device)
                for path, name in ((rt.PATH_TSLAB, "tslab"), (rt.PATH_RING, "ring"), (rt.PATH_GENERIC, "generic")):
                    if name == "generic" and (envs > 4096 or len(sys.argv) > 1):
                        continue
                    def fn():
                        rc = lib.ppo_tvf_targets(r.data_ptr(), envs, d.data_ptr(), 1, envs, vs.data_ptr(), envs * 128, 128,
                                                 steps, envs, 128, 128, ctypes.byref(ctab), path, out.data_ptr(), 0, 0, 0,
                                                 torch.cuda.current_stream().cuda_stream)
                        _lib.check(rc, "tvf")
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(10):
                        flush.zero_()
                        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        s.record(); fn(); e.record(); torch.cuda.synchronize()
                        ts.append(s.elapsed_time(e) * 1e3)
                    ts.sort()
                    nbytes = 4 * (steps + 1) * envs * 128 + 4 * steps * envs * 128 + 5 * steps * envs
                    med = ts[len(ts) // 2]
                    print(f"tvf {steps}x{envs}x128 done_p={done_p} {mode:10s} E={tabs.n_entries} nmax={tabs.nmax} {name:8s}: "
                          f"med {med:9.1f}us best {ts[0]:9.1f}us  {nbytes / med / 1e3:7.0f} GB/s  "
                          f"{steps * envs * 128 / med / 1e3:7.1f} Gtargets/s  frac {nbytes / med / 1e3 / 6548.8:.3f}")

if __name__ == "__main__":
    main()
