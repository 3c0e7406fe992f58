"""Host-side time line of one end-to-end calculate_returns call with pinned buffers (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_b200 import rollout, returns_truncated as rt, tvf as tvf_mod

def main():
    steps, envs, n_h = 256, 4096, 128
    hz = tvf_mod.get_value_head_horizons(n_h, 30000)
    g = torch.Generator().manual_seed(0)
    rewards = torch.randint(-1, 3, (steps, envs), generator=g).float().pin_memory()
    dones = (torch.rand((steps, envs), generator=g) < 0.03).pin_memory()
    vs = (torch.randn((steps + 1, envs, n_h), generator=g) * 0.4 + 0.1).pin_memory()
    values = vs[:, :, -1].contiguous().pin_memory()
    out = {"advantage": torch.empty((steps, envs)).pin_memory(), "ext_returns": torch.empty((steps, envs)).pin_memory(),
           "advantages": torch.empty(steps * envs).pin_memory(), "tvf_returns": torch.empty((steps, envs, n_h)).pin_memory()}
    kw = dict(value_samples=vs, horizons=hz, gamma=0.99997, mode="advanced", distribution="exponential", n_step=20,
              max_samples=8, seed=1)
    for _ in range(3):
        rollout.calculate_returns(rewards, values, dones, 0.999, 0.95, 0.95, tvf=kw, out=out)
    # pieces
    def t(fn, n=10):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    print("draw samples       %.3f ms" % t(lambda: rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)))
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    print("build tables       %.3f ms" % t(lambda: rt.build_tables(0.99997, steps, hz, hz, table)))
    tabs = rt.build_tables(0.99997, steps, hz, hz, table)
    print("upload tables      %.3f ms" % t(lambda: rt._upload_tables(tabs, torch.device("cuda", 0))))
    print("gae part only      %.3f ms" % t(lambda: rollout.calculate_returns(rewards, values, dones, 0.999, 0.95, 0.95, out=out)))
    print("tvf host blocks    %.3f ms" % t(lambda: (rt._launch_host_blocks(0.99997, rewards, dones, hz, hz, vs, table, False, out["tvf_returns"]), torch.cuda.current_stream().synchronize())))
    print("full call          %.3f ms" % t(lambda: rollout.calculate_returns(rewards, values, dones, 0.999, 0.95, 0.95, tvf=kw, out=out)))
    # enqueue time of the full call (host returns after the final sync, so measure the TVF part without sync)
    t0 = time.perf_counter(); j = rt._launch_host_blocks(0.99997, rewards, dones, hz, hz, vs, table, False, out["tvf_returns"], defer_join=True); t1 = time.perf_counter()
    j(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("tvf enqueue %.3f ms, until complete %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))

if __name__ == "__main__":
    main()
