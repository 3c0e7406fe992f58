"""Scratch micro-benchmark of the scan kernels (development aid, not the judged bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import returns, synthetic, _lib, _device as dev

def timeit(fn, iters=50, flush=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for steps, envs in [(256, 4096), (256, 32768), (128, 128)]:
        r, d, v = synthetic.rollout(steps, envs, seed=1)
        r, d, v = (torch.from_numpy(x).cuda() for x in (r, d, v))
        for stats in (False, True):
            fn = lambda: returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.95, with_stats=stats)
            med, best = timeit(fn)
            medc, bestc = timeit(fn, flush=flush)
            nbytes = steps * envs * (4 + 4 + 1 + 4 + 4)
            print(f"gae+tdl {steps}x{envs} stats={stats}: warm med {med:.1f}us best {best:.1f}us "
                  f"({nbytes / best / 1e3:.0f} GB/s) | cold med {medc:.1f}us best {bestc:.1f}us "
                  f"({nbytes / bestc / 1e3:.0f} GB/s)")

if __name__ == "__main__":
    main()
