"""One TVF ring-kernel launch per mode, for ncu (development aid)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_b200 import tvf as _tvf_mod
from ppo_b200 import returns_truncated as rt, _lib, _tvf_tables

def main():
    steps, envs = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    mode = sys.argv[2] if len(sys.argv) > 2 else "advanced8"
    path = {"ring": rt.PATH_RING, "tslab": rt.PATH_TSLAB}[sys.argv[3] if len(sys.argv) > 3 else "tslab"]
    lib = _lib.load()
    hz = _tvf_mod.get_value_head_horizons(128, 30000)
    device = torch.device("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    r = torch.randint(-1, 3, (steps, envs), device=device, generator=g).float()
    d = (torch.rand((steps, envs), device=device, generator=g) < 0.03).view(torch.uint8)
    vs = torch.randn((steps + 1, envs, 128), device=device, generator=g) * 0.4 + 0.1
    out = torch.empty((steps, envs, 128), device=device)
    if mode == "advanced8":
        table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    else:
        table = np.zeros([128, 1], dtype=np.int32) + 20
    tabs = rt.build_tables(0.99997, steps, hz, hz, table)
    ctab, blob = rt._upload_tables(tabs, device)
    for _ in range(3):
        rc = lib.ppo_tvf_targets(r.data_ptr(), envs, d.data_ptr(), 1, envs, vs.data_ptr(), envs * 128, 128,
                                 steps, envs, 128, 128, ctypes.byref(ctab), path, out.data_ptr(), 0, 0, 0,
                                 torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "tvf")
    torch.cuda.synchronize()
    print("ok", float(out.sum()))

if __name__ == "__main__":
    main()
