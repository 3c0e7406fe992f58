"""Development check + timing of the column slab TVF kernel (PATH_CSLAB) against the generic and transposed slab kernels.
Not the judged bench.  usage: python tools/dev_cslab.py [check] [time] [full]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ppo_b200 import _lib
from ppo_b200 import returns_truncated as rt
from ppo_b200 import tvf as tvf_mod


def launch(lib, path, r, d, vs, ctab, out, stats=None, ws=None, extras=None):
    steps, envs = r.shape
    k, v = out.shape[2], vs.shape[2]
    rc = lib.ppo_tvf_targets_ex(r.data_ptr(), r.stride(0), d.data_ptr(), 1, d.stride(0), vs.data_ptr(), vs.stride(0),
                                vs.stride(1), steps, envs, k, v, ctypes.byref(ctab), path, out.data_ptr(),
                                stats.data_ptr() if stats is not None else 0, ws.data_ptr() if ws is not None else 0,
                                ws.numel() if ws is not None else 0,
                                ctypes.byref(extras) if extras is not None else None,
                                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ppo_tvf_targets_ex")


def check():
    lib = _lib.load()
    dev = torch.device("cuda")
    bad = 0
    cases = [
        # steps, envs, heads, max_h, dist, mode, n_step, samples, done_p, log
        (256, 300, 128, 30000, "exponential", "advanced", 20, 8, 0.03, False),
        (256, 5, 128, 30000, "exponential", "advanced", 20, 8, 0.0, False),
        (128, 200, 128, 30000, "exponential", "advanced", 20, 8, 0.03, False),
        (256, 149, 128, 30000, "exponential", "standard", 20, 8, 0.03, True),
        (256, 64, 128, 30000, "uniform", "advanced", 20, 8, 0.1, False),
        (128, 64, 128, 30000, "uniform", "advanced", 20, 12, 0.01, True),
        (256, 33, 64, 5000, "exponential", "advanced", 40, 5, 0.02, False),
        (128, 33, 32, 1000, "hyperbolic", "clipped", 30, 4, 0.1, False),
        (256, 20, 128, 30000, "fixed", "standard", 20, 1, 0.03, False),
        (256, 20, 128, 30000, "fixed", "standard", 256, 1, 0.003, False),
        (256, 400, 8, 60, "exponential", "advanced", 10, 3, 0.05, False),
        (128, 33, 32, 1000, "exponential", "advanced", 30, 4, 0.1, False),
        (256, 2, 32, 1000, "exponential", "standard", 10, 4, 0.0, False),
        (256, 700, 96, 9000, "exponential", "advanced", 10, 6, 0.05, True),
    ]
    for steps, envs, heads, max_h, dist, mode, n_step, samples, done_p, use_log in cases:
        hz = tvf_mod.get_value_head_horizons(heads, max_h)
        g = torch.Generator(device=dev)
        g.manual_seed(steps + envs)
        r = torch.randint(-1, 3, (steps, envs), device=dev, generator=g).float()
        d = (torch.rand((steps, envs), device=dev, generator=g) < done_p).view(torch.uint8)
        vs = torch.randn((steps + 1, envs, len(hz)), device=dev, generator=g) * 0.4 + 0.1
        table = rt.draw_n_step_samples(dist, mode, hz, steps, n_step, samples, seed=3)
        tabs = rt.build_tables(0.99997, steps, hz, hz, table, use_log)
        ctab, blob = rt._upload_tables(tabs, dev)
        ref = torch.empty((steps, envs, len(hz)), device=dev)
        got = torch.full((steps, envs, len(hz)), float("nan"), device=dev)
        stats_r = torch.zeros(16, dtype=torch.float64, device=dev)
        stats_g = torch.zeros(16, dtype=torch.float64, device=dev)
        ws = torch.zeros(int(lib.ppo_tvf_workspace_bytes(steps, envs)), dtype=torch.uint8, device=dev)
        vo_r = torch.zeros((steps + 1, envs), device=dev)
        vo_g = torch.full((steps + 1, envs), float("nan"), device=dev)
        vw = np.linspace(-0.2, 0.3, len(hz)).astype(np.float32)
        ex_r, keep_r = rt._extras(vo_r, vw)
        ex_g, keep_g = rt._extras(vo_g, vw)
        launch(lib, rt.PATH_GENERIC, r, d, vs, ctab, ref, stats_r, ws, ex_r)
        torch.cuda.synchronize()
        try:
            launch(lib, rt.PATH_CSLAB, r, d, vs, ctab, got, stats_g, ws, ex_g)
            torch.cuda.synchronize()
        except ValueError as exc:
            print("UNSUPPORTED", (steps, envs, heads, dist, mode, n_step, samples), exc)
            continue
        err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        nan = int(torch.isnan(got).sum().item())
        verr = (vo_g - vo_r).abs().max().item() / max(vo_r.abs().max().item(), 1e-30)
        sg = stats_g.cpu().numpy()
        last = got[:, :, -1].double()
        sr = np.zeros(16)
        sr[0], sr[1], sr[2], sr[3], sr[4] = last.numel(), last.sum().item(), (last * last).sum().item(), last.min().item(), last.max().item()
        sr[12] = (last == 0).sum().item()
        serr = np.abs(sr - sg).max() / max(np.abs(sr).max(), 1e-30)
        ok = err <= 1e-5 and nan == 0 and verr <= 1e-5 and serr < 1e-9
        bad += 0 if ok else 1
        print(("ok  " if ok else "BAD ") + f"{steps}x{envs}x{heads} {dist}/{mode} n={n_step} c={samples} p={done_p} log={use_log}: "
              f"rel err {err:.2e} nan {nan} value err {verr:.2e} stats err {serr:.1e} E={tabs.n_entries} nmax={tabs.nmax} "
              f"path {lib.ppo_tvf_last_path():#x}")
        if nan:
            m = torch.isnan(got)
            print("   nan envs:", torch.nonzero(m.any(0).any(1)).flatten().tolist()[:40], " nan k:",
                  torch.nonzero(m.any(0).any(0)).flatten().tolist()[:40], " nan t:",
                  torch.nonzero(m.any(1).any(1)).flatten().tolist()[:12])
        if not ok and nan == 0:
            diff = (got - ref).abs()
            badm = diff > 1e-5 * ref.abs().max()
            idx = torch.nonzero(badm)[:8]
            print("   first bad (t, env, k):", idx.tolist(), " count", int(badm.sum()))
            print("   bad t:", torch.nonzero(badm.any(1).any(1)).flatten().tolist()[:40])
            print("   bad k:", torch.nonzero(badm.any(0).any(0)).flatten().tolist()[:40])
            print("   bad env:", torch.nonzero(badm.any(0).any(1)).flatten().tolist()[:40])
    print("check:", "all ok" if bad == 0 else f"{bad} BAD")
    return bad


def timing(full):
    lib = _lib.load()
    dev = torch.device("cuda")
    hz = tvf_mod.get_value_head_horizons(128, 30000)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    cfgs = [(256, 4096), (128, 1024)] if full else [(256, 4096)]
    for steps, envs in cfgs:
        g = torch.Generator(device=dev)
        g.manual_seed(0)
        r = torch.randint(-1, 3, (steps, envs), device=dev, generator=g).float()
        d = (torch.rand((steps, envs), device=dev, generator=g) < 3 / 101).view(torch.uint8)
        vs = torch.randn((steps + 1, envs, 128), device=dev, generator=g) * 0.4 + 0.1
        out = torch.empty((steps, envs, 128), device=dev)
        stats = torch.zeros(16, dtype=torch.float64, device=dev)
        ws = torch.zeros(int(lib.ppo_tvf_workspace_bytes(steps, envs)), dtype=torch.uint8, device=dev)
        vo = torch.zeros((steps + 1, envs), device=dev)
        vw = np.linspace(-0.2, 0.3, 128).astype(np.float32)
        ex, keep = rt._extras(vo, vw)
        table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
        tabs = rt.build_tables(0.99997, steps, hz, hz, table)
        ctab, blob = rt._upload_tables(tabs, dev)
        nbytes = 4 * (steps + 1) * envs * 128 + 4 * steps * envs * 128 + 5 * steps * envs
        for path, name in ((rt.PATH_CSLAB, "cslab"), (rt.PATH_TSLAB, "tslab")):
            for _ in range(3):
                launch(lib, path, r, d, vs, ctab, out, stats, ws, ex)
            torch.cuda.synchronize()
            ts, ts3 = [], []
            for _ in range(10):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                launch(lib, path, r, d, vs, ctab, out, stats, ws, ex)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e) * 1e3)
            for _ in range(5):  # back to back: the launch overhead of all but the first hides behind the previous kernel
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(4):
                    launch(lib, path, r, d, vs, ctab, out, stats, ws, ex)
                e.record()
                torch.cuda.synchronize()
                ts3.append(s.elapsed_time(e) * 1e3 / 4)
            ts.sort(), ts3.sort()
            med = ts[len(ts) // 2]
            print(f"tvf {steps}x{envs}x128 adv8 {name:6s} dbg={os.environ.get('PPO_TVF_DEBUG', '0'):3s}: med {med:8.1f} us best {ts[0]:8.1f} us "
                  f"{nbytes / med / 1e3:6.0f} GB/s  frac {nbytes / med / 1e3 / 6548.8:.3f} | 4 back to back: {ts3[len(ts3) // 2]:8.1f} us each", flush=True)


def stamps():
    """One launch with PPO_TVF_DEBUG bit 256: time line of CTA 0 (us since the first stamp)."""
    lib = _lib.load()
    dev = torch.device("cuda")
    hz = tvf_mod.get_value_head_horizons(128, 30000)
    steps, envs = 256, 4096
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    r = torch.randint(-1, 3, (steps, envs), device=dev, generator=g).float()
    d = (torch.rand((steps, envs), device=dev, generator=g) < 3 / 101).view(torch.uint8)
    vs = torch.randn((steps + 1, envs, 128), device=dev, generator=g) * 0.4 + 0.1
    out = torch.empty((steps, envs, 128), device=dev)
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    tabs = rt.build_tables(0.99997, steps, hz, hz, table)
    ctab, blob = rt._upload_tables(tabs, dev)
    for _ in range(2):
        launch(lib, rt.PATH_CSLAB, r, d, vs, ctab, out)
    buf = (ctypes.c_ulonglong * 832)()
    lib.ppo_tvf_cslab_debug_times.argtypes = [ctypes.c_void_p]
    lib.ppo_tvf_cslab_debug_times(buf)
    t = np.array(buf[:], dtype=np.float64)
    t0 = t[:512][t[:512] > 0].min()
    us = lambda i: (t[i] - t0) / 1e3 if t[i] > 0 else float("nan")
    print("coordinators done", us(500), us(501), " consumers done", us(502), us(503))
    st, en = t[512:512 + 148], t[672:672 + 148]
    print("CTA start (us): min %.1f max %.1f | CTA end: min %.1f median %.1f max %.1f | duration: min %.1f median %.1f max %.1f" % (
        (st.min() - t0) / 1e3, (st.max() - t0) / 1e3, (en.min() - t0) / 1e3, (np.median(en) - t0) / 1e3, (en.max() - t0) / 1e3,
        (en - st).min() / 1e3, np.median(en - st) / 1e3, (en - st).max() / 1e3))
    dur = (en - st) / 1e3
    print("slowest CTAs:", np.argsort(-dur)[:12].tolist(), np.sort(-dur)[:12].round(1).tolist())
    print("fastest CTAs:", np.argsort(dur)[:12].tolist(), np.sort(dur)[:12].round(1).tolist())
    for e in range(4):
        b = 64 * e
        print(f"env {e}: transform start {us(b):7.2f} rfull {us(b + 1):7.2f} | " +
              " | ".join(f"g{g_} free {us(b + 2 + 2 * g_):7.2f} full {us(b + 3 + 2 * g_):7.2f}" for g_ in range(4)))
        print("       g1 chunks: " + " | ".join(f"wait {us(b + 54 + 3 * c_):7.2f} got {us(b + 55 + 3 * c_):7.2f} done {us(b + 56 + 3 * c_):7.2f}" for c_ in range(2)))
        for hg in range(4):
            c = b + 12 + 12 * hg
            print(f"       job hg{hg}: grab {us(c):7.2f} ready {us(c + 1):7.2f} | entries done " +
                  " ".join(f"{us(c + 5 + u_):7.2f}" for u_ in range(4)) + f" | computed {us(c + 2):7.2f} "
                  f"before grant {us(c + 9):7.2f} granted {us(c + 10):7.2f} written {us(c + 3):7.2f} stored {us(c + 4):7.2f}")


if __name__ == "__main__":
    args = sys.argv[1:] or ["check", "time"]
    rc = 0
    if "check" in args:
        rc = check()
    if "time" in args:
        timing("full" in args)
    if "stamps" in args:
        stamps()
    sys.exit(1 if rc else 0)
