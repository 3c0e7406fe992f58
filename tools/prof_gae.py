"""A few fused GAE + TD(lambda) + statistics launches for ncu (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import returns, synthetic

def main():
    steps, envs = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    r, d, v = synthetic.rollout(steps, envs, seed=1)
    r, d, v = (torch.from_numpy(x).cuda() for x in (r, d, v))
    for _ in range(3):
        out = returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.95, with_stats=True)
    torch.cuda.synchronize()
    print("ok", float(out[0].sum()))

if __name__ == "__main__":
    main()
