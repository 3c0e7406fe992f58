#!/usr/bin/env python
"""Benchmark of the return / advantage hot path (BASELINE.json metric: return targets/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" is one full pass of the return stage (rl/rollout.py:1182-1285 + :1886-1897) over one synthetic rollout
of 256 steps x 4096 envs x 128 horizons PER GPU (weak scaling; 8 GPUs = BASELINE config 5, 32768 envs), in the
reference's own data flow: 128-horizon TVF targets (exponential / advanced, 8 samples, n_step 20) + statistics of
the longest horizon + the value estimate GAE reads, rediscounted from the TVF gamma 0.99997 to gamma 0.999
(rl/rollout.py:1196 -> rl/tvf.py:303-433) -> GAE advantages + TD(lambda) returns + their statistics -> all-reduce
of {count, sum, sumsq} (N > 1) -> advantage normalisation.  ``value`` counts steps x envs x horizons TVF targets
per second over all ranks (the 2 x steps x envs GAE outputs of the same pass are not counted).

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU arm: the NumPy oracle (a restatement of
the reference, pinned to it by tests/golden) on the host cores; nothing else in this file executes ``oracle/``
except the ``cpu_baseline`` leg.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

STEPS, ENVS_PER_GPU, HORIZONS, MAX_HORIZON = 256, 4096, 128, 30000
GAMMA, LAMBDA_POLICY, LAMBDA_VALUE = 0.999, 0.95, 0.95
TVF_GAMMA, N_STEP, SAMPLES, SAMPLE_SEED = 0.99997, 20, 8, 1
METRIC = "return targets (steps x envs x horizons) per second, full pass: GAE + adv-norm + 128-horizon TVF targets"
UNIT = "targets/s"


STRONG_TOTAL_ENVS = 32768  # BASELINE config 5 as stated: 256 x 32768 x 128 over the whole job


def envs_per_gpu(n_gpus, scaling):
    return ENVS_PER_GPU if scaling == "weak" else STRONG_TOTAL_ENVS // n_gpus


def config(n_gpus, scaling="weak"):
    per_gpu = envs_per_gpu(n_gpus, scaling)
    return {
        "workload": f"full return pass, {STEPS} steps x {per_gpu} envs x {HORIZONS} horizons per GPU "
                    f"(total {per_gpu * n_gpus} envs; " +
                    ("8 GPUs = BASELINE config 5" if scaling == "weak" else "BASELINE config 5 at every N, --scaling strong") +
                    "): TVF targets + rediscounted value estimate + GAE / TD(lambda) + advantage normalisation",
        "steps": STEPS, "envs_per_gpu": per_gpu, "envs_total": per_gpu * n_gpus, "horizons": HORIZONS,
        "max_horizon": MAX_HORIZON, "gae": {"gamma": GAMMA, "lambda_policy": LAMBDA_POLICY, "lambda_value": LAMBDA_VALUE},
        "tvf": {"gamma": TVF_GAMMA, "distribution": "exponential", "mode": "advanced", "n_step": N_STEP,
                "samples": SAMPLES, "seed": SAMPLE_SEED},
        "sharding": f"env axis over {n_gpus} rank(s); the 3 float64 advantage moments are exchanged inside the GAE launch "
                    f"over NVLink peer mailboxes (NCCL all-reduce with PPO_PEER_EXCHANGE=0)" if n_gpus > 1
                    else "single GPU, no collective",
        "l2": "per-step working set 1.1 GB per GPU exceeds the 126 MB L2, no explicit flush",
        "sample_table": "a fresh table per pass, drawn on the host with the reference's NumPy RNG calls, resolved "
                        "(ppo_tvf_tables_build) and uploaded inside both timed regions (device-resident and e2e)",
    }


# ------------------------------------------------------------------------------------------ CPU arm
def _oracle_pass(rewards, dones, values, value_samples, horizons):
    """One full pass with the NumPy oracle (reference algorithm)."""
    import oracle
    n1, envs, k = value_samples.shape
    values = oracle.get_rediscounted_value_estimate(value_samples.reshape(n1 * envs, k), TVF_GAMMA, GAMMA,
                                                    horizons).reshape(n1, envs)
    adv = oracle.gae(rewards, values[:-1], values[-1], dones, GAMMA, LAMBDA_POLICY)
    ret = oracle.td_lambda(rewards, values[:-1], values[-1], dones, GAMMA, LAMBDA_VALUE)
    norm = oracle.normalize_advantages(adv)
    tvf = oracle.get_return_estimate("exponential", "advanced", TVF_GAMMA, rewards, dones, horizons, horizons,
                                     value_samples, n_step=N_STEP, max_samples=SAMPLES, seed=SAMPLE_SEED)
    return adv, ret, norm, tvf


def _cpu_inputs(envs, seed):
    import oracle
    from ppo_b200 import synthetic
    hz = oracle.get_value_head_horizons(HORIZONS, MAX_HORIZON)
    r, d, v = synthetic.rollout(STEPS, envs, seed=seed)
    vs = synthetic.value_samples(STEPS, envs, len(hz), seed=seed)
    return r, d, v, vs, hz


_WORKER = {}


def _worker_init(envs, seed_base):
    import multiprocessing
    ident = multiprocessing.current_process()._identity
    _WORKER["data"] = _cpu_inputs(envs, seed_base + (ident[0] if ident else 0))


def _worker_step(_):
    r, d, v, vs, hz = _WORKER["data"]
    t0 = time.perf_counter()
    _oracle_pass(r, d, v, vs, hz)
    return time.perf_counter() - t0


def cpu_baseline_single(envs=1024):
    """The oracle on ONE host core over a bounded sample (256 steps x `envs` envs x 128 horizons)."""
    r, d, v, vs, hz = _cpu_inputs(envs, seed=0)
    t0 = time.perf_counter()
    _oracle_pass(r, d, v, vs, hz)
    dt = time.perf_counter() - t0
    # BASELINE config 1: the reference's own CPU case, GAE over 128 steps x 128 envs (rl/returns.py:58-67)
    import oracle
    from ppo_b200 import synthetic
    r1, d1, v1 = synthetic.rollout(128, 128, seed=0)
    oracle.gae(r1, v1[:-1], v1[-1], d1, GAMMA, LAMBDA_POLICY)
    c0 = time.perf_counter()
    for _ in range(50):
        oracle.gae(r1, v1[:-1], v1[-1], d1, GAMMA, LAMBDA_POLICY)
    c1_us = (time.perf_counter() - c0) / 50 * 1e6
    return {"value": STEPS * envs * HORIZONS / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "c1_gae_128x128": {"us_per_call": c1_us, "step_env_per_s": 128 * 128 / (c1_us * 1e-6),
                               "what": "oracle.gae (NumPy restatement of rl/returns.py:58-67), one core"},
            "sample": f"one full pass of the NumPy oracle over {STEPS} steps x {envs} envs x {HORIZONS} horizons "
                      f"({envs}/{ENVS_PER_GPU} of the per-GPU workload), {dt:.2f} s, single-threaded like the reference; "
                      f"host has {os.cpu_count()} logical cores"}


def run_reference(args):
    """CPU arm: the oracle port of the reference on all usable host cores (one process per core, envs are
    independent).  Each step is a bounded sample: every worker runs one full pass over its own 256-env rollout
    (256 envs = the reference's default agent count, rl/config.py:791)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, 16))
    warmup = args.warmup  # honoured as given: the sample below is sized for steps + warm-up together
    # bounded sample: ~3.4 M targets/s per core -> 256 steps x e envs x 128 horizons take ~e / 104 s; keep the whole
    # run near two minutes, never more than the reference's default 256 agents per call (rl/config.py:791)
    envs_per_worker = int(104 * 120 / (args.steps + warmup))
    envs_per_worker = max(16, min(256, envs_per_worker // 16 * 16))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_worker_init, initargs=(envs_per_worker, 1000)) as pool:
        for _ in range(warmup):
            pool.map(_worker_step, range(workers), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_step, range(workers), chunksize=1)
        dt = time.perf_counter() - t0
    sample_envs = envs_per_worker * workers
    value = STEPS * sample_envs * HORIZONS * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config(args.gpus, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"each step = one full oracle pass over {STEPS} steps x {sample_envs} envs x {HORIZONS} "
                                   f"horizons ({workers} processes x {envs_per_worker} envs), host has {cores} usable cores; "
                                   f"{warmup} warm-up step(s) as requested"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while a timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: report it rather than guessing
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def other_configs(device, peak_gbs, iters=40):
    """Kernel-level figures of the other BASELINE.json configurations (parity-test cases, not bench lines):
    back-to-back launches through the C ABI on rotating buffer sets larger than L2, device time per launch from
    CUDA events, algorithmic bytes as in DESIGN.md."""
    import ctypes
    import torch
    from ppo_b200 import _lib
    from ppo_b200 import returns_truncated as rt
    from ppo_b200 import tvf as tvf_mod
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=device)
    g.manual_seed(123)

    def timed(launch, nbytes, units):
        for i in range(6):
            launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            launch(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        return {"us_per_launch": us, "targets_per_s": units / (us * 1e-6), "achieved_gbs": nbytes / us / 1e3,
                "frac_of_hbm_peak": nbytes / us / 1e3 / peak_gbs, "algorithmic_bytes": nbytes}

    out = {}
    # ---- c2: GAE + TD(lambda) + statistics, then advantage normalisation, 256 x 4096, contiguous values
    n, a, sets = 256, 4096, 8
    bufs = []
    for _ in range(sets):
        r = torch.randint(-1, 3, (n, a), device=device, generator=g).float()
        d = (torch.rand((n, a), device=device, generator=g) < 3 / 101).view(torch.uint8)
        v = torch.randn((n + 1, a), device=device, generator=g) * 0.4 + 0.1
        bufs.append((r, d, v, torch.empty((n, a), device=device), torch.empty((n, a), device=device),
                     torch.empty(n * a, device=device)))
    st = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=device)
    ws = torch.zeros(int(lib.ppo_gae_workspace_bytes(a)), dtype=torch.uint8, device=device)

    fused = []

    def c2(i):
        r, d, v, adv, ret, norm = bufs[i % sets]
        rc = lib.ppo_gae_normalized(r.data_ptr(), a, v.data_ptr(), a, 1, v[n].data_ptr(), 1, d.data_ptr(), _lib.DONE_U8,
                                    a, n, a, GAMMA, LAMBDA_POLICY, LAMBDA_VALUE, 1e-8, 0.0, adv.data_ptr(),
                                    ret.data_ptr(), norm.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "ppo_gae_normalized")
        fused.append(rc)
    out["c2_gae_tdlambda_stats_advnorm_256x4096"] = dict(
        timed(c2, n * a * 21, n * a), launches_per_pass=1 if fused and fused[-1] == 1 else 2,
        note="GAE + TD(lambda) + statistics + advantage normalisation in one cooperative launch; 21 B per (t, env): "
        "r 4 + V 4 + done 1 in, raw adv 4 + TD(lambda) return 4 + normalised adv 4 out; launch latency is a large "
        "part of the time at 22 MB")
    del bufs

    # ---- c3: TVF targets 128 steps x 1024 envs x 128 horizons, exponential / advanced 8 samples and fixed n-step
    n, a, k = 128, 1024, HORIZONS
    hz = tvf_mod.get_value_head_horizons(k, MAX_HORIZON)
    sets = 6
    bufs = []
    for _ in range(sets):
        r = torch.randint(-1, 3, (n, a), device=device, generator=g).float()
        d = (torch.rand((n, a), device=device, generator=g) < 3 / 101).view(torch.uint8)
        vs = torch.randn((n + 1, a, k), device=device, generator=g) * 0.4 + 0.1
        bufs.append((r, d, vs, torch.empty((n, a, k), device=device)))
    nbytes = 4 * (n + 1) * a * k + 4 * n * a * k + 5 * n * a
    for name, table in (("exponential_advanced_8", rt.draw_n_step_samples("exponential", "advanced", hz, n, N_STEP,
                                                                            SAMPLES, seed=SAMPLE_SEED)),
                        ("fixed_n_step_20", rt.draw_n_step_samples("fixed", "advanced", hz, n, N_STEP, SAMPLES))):
        ctab, keep = rt._upload_tables(rt.build_tables(TVF_GAMMA, n, hz, hz, table), device)

        def c3(i):
            r, d, vs, tg = bufs[i % sets]
            _lib.check(lib.ppo_tvf_targets(r.data_ptr(), a, d.data_ptr(), _lib.DONE_U8, a, vs.data_ptr(), a * k, k,
                                           n, a, k, k, ctypes.byref(ctab), rt.PATH_AUTO, tg.data_ptr(), 0, 0, 0,
                                           stream), "ppo_tvf_targets")
        res = timed(c3, nbytes, n * a * k)
        res["kernel"] = {1: "generic", 2: "ring", 3: "tslab", 4: "cslab", 5: "qslab"}.get(lib.ppo_tvf_last_path() & 0xff)
        out[f"c3_tvf_128x1024x128_{name}"] = res
        del keep
    del bufs

    # ---- c5 shape (256 x 4096 x 128) with a FIXED n-step table ("fixed" distribution: one n-step per horizon): one
    # gather pair per target instead of eight, so the HBM transfers dominate -- through AUTO and on the ring kernel
    n, a, k = 256, 4096, HORIZONS
    r = torch.randint(-1, 3, (n, a), device=device, generator=g).float()
    d = (torch.rand((n, a), device=device, generator=g) < 3 / 101).view(torch.uint8)
    vs = torch.randn((n + 1, a, k), device=device, generator=g) * 0.4 + 0.1
    tg = torch.empty((n, a, k), device=device)
    nbytes = 4 * (n + 1) * a * k + 4 * n * a * k + 5 * n * a
    table = rt.draw_n_step_samples("fixed", "advanced", hz, n, N_STEP, SAMPLES)
    ctab, keep = rt._upload_tables(rt.build_tables(TVF_GAMMA, n, hz, hz, table), device)
    for name, path in (("auto", rt.PATH_AUTO), ("ring", rt.PATH_RING)):
        def c5(i):
            _lib.check(lib.ppo_tvf_targets(r.data_ptr(), a, d.data_ptr(), _lib.DONE_U8, a, vs.data_ptr(), a * k, k,
                                           n, a, k, k, ctypes.byref(ctab), path, tg.data_ptr(), 0, 0, 0, stream),
                       "ppo_tvf_targets")
        try:
            res = timed(c5, nbytes, n * a * k)
            res["kernel"] = {1: "generic", 2: "ring", 3: "tslab", 4: "cslab", 5: "qslab"}.get(lib.ppo_tvf_last_path() & 0xff)
        except Exception as e:  # a path that declines this shape is reported, not hidden
            res = {"error": repr(e)}
        res["note"] = "1.07 GB per launch (> L2), targets only (no statistics, no value estimate)"
        out[f"c5_tvf_256x4096x128_fixed_n_step_20_{name}"] = res
    del keep
    # ---- the same shape with the headline's sample table (exponential / advanced, 8 samples), targets only, on each of
    # the three slab kernels explicitly: PATH_AUTO takes the transposed slab kernel; the other two are kept honest here
    table = rt.draw_n_step_samples("exponential", "advanced", hz, n, N_STEP, SAMPLES, seed=SAMPLE_SEED)
    ctab, keep = rt._upload_tables(rt.build_tables(TVF_GAMMA, n, hz, hz, table), device)
    for name, path in (("tslab", rt.PATH_TSLAB), ("qslab", rt.PATH_QSLAB), ("cslab", rt.PATH_CSLAB)):
        def c5s(i):
            _lib.check(lib.ppo_tvf_targets(r.data_ptr(), a, d.data_ptr(), _lib.DONE_U8, a, vs.data_ptr(), a * k, k,
                                           n, a, k, k, ctypes.byref(ctab), path, tg.data_ptr(), 0, 0, 0, stream),
                       "ppo_tvf_targets")
        try:
            res = timed(c5s, nbytes, n * a * k)
        except Exception as e:
            res = {"error": repr(e)}
        res["note"] = "targets only (no statistics, no value estimate), explicit kernel path"
        out[f"c5_tvf_256x4096x128_exponential_advanced_8_{name}"] = res
    del r, d, vs, tg, keep

    # ---- c1: GAE over 128 steps x 128 envs (the reference's own CPU-sized case): one launch, latency bound
    n, a = 128, 128
    r = torch.randint(-1, 3, (n, a), device=device, generator=g).float()
    d = (torch.rand((n, a), device=device, generator=g) < 3 / 101).view(torch.uint8)
    v = torch.randn((n + 1, a), device=device, generator=g) * 0.4 + 0.1
    adv = torch.empty((n, a), device=device)
    ws1 = torch.zeros(max(int(lib.ppo_gae_workspace_bytes(a)), 4096), dtype=torch.uint8, device=device)

    def c1(i):
        _lib.check(lib.ppo_gae(r.data_ptr(), a, v.data_ptr(), a, 1, v[n].data_ptr(), 1, d.data_ptr(), _lib.DONE_U8, a, n, a,
                               GAMMA, LAMBDA_POLICY, LAMBDA_POLICY, adv.data_ptr(), 0, 0, ws1.data_ptr(), ws1.numel(),
                               stream), "ppo_gae")
    res = timed(c1, n * a * 13, n * a)
    res["note"] = ("GAE advantages only (rl/returns.py:58-67), 13 B per (t, env), 213 KB in all: pure launch latency; "
                   "the CPU time of the same call is in cpu_baseline.c1_gae_128x128")
    out["c1_gae_128x128"] = res

    # ---- c4: V-trace 256 steps x 4096 envs x 18 actions (one GPU's worth of the 8-way sharded config is 512 envs)
    n, b, act, sets = 256, 4096, 18, 4
    bufs = []
    for _ in range(sets):
        beh = torch.log_softmax(torch.randn((n, b, act), device=device, generator=g), dim=-1)
        tgt = torch.log_softmax(torch.randn((n, b, act), device=device, generator=g), dim=-1)
        actions = torch.randint(0, act, (n, b), device=device, generator=g)
        r = torch.randint(-1, 3, (n, b), device=device, generator=g).float()
        d = (torch.rand((n, b), device=device, generator=g) < 3 / 101).view(torch.uint8)
        v = torch.randn((n + 1, b), device=device, generator=g) * 0.4 + 0.1
        bufs.append((beh, tgt, actions, r, d, v, [torch.empty((n, b), device=device) for _ in range(3)]))

    def c4(i):
        beh, tgt, actions, r, d, v, o = bufs[i % sets]
        _lib.check(lib.ppo_vtrace(beh.data_ptr(), tgt.data_ptr(), actions.data_ptr(), 8, r.data_ptr(), d.data_ptr(),
                                  _lib.DONE_U8, v.data_ptr(), v[n].data_ptr(), n, b, act, GAMMA, 1.0,
                                  o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), None, stream), "ppo_vtrace")
    out["c4_vtrace_256x4096x18"] = dict(
        timed(c4, n * b * 37, n * b), note="37 useful B per (t, env); each gathered 4-byte log-probability costs a 64-byte "
        "DRAM burst: ncu measures 169 MB of DRAM reads per launch (161 B per (t, env), profiles/r01c_vtrace_*), i.e. "
        "about 4.1 TB/s = 63 % of the measured HBM peak in traffic terms")
    del bufs

    # ---- intrinsic-return normalisation + GAE (rl/rollout.py:1140-1180), 256 x 4096, device tensors
    from ppo_b200 import normalization as nz
    n, a = 256, 4096
    ir = nz.IntrinsicReturns(a, device=device)
    r = torch.rand((n, a), device=device, generator=g) * 0.5
    d = torch.rand((n, a), device=device, generator=g) < 3 / 101
    v = torch.randn((n + 1, a), device=device, generator=g) * 0.4 + 0.5

    def f3(i):
        ir.calculate_intrinsic_returns(r, v, d, 0.99, LAMBDA_POLICY, normalize=True, center=False)
    # r 4 + done 1 in, tracker rows 8 out + 8 + 8 re-read (two row passes), r 4 in + scaled r 4 out, GAE 13 in + 8 out
    res = timed(f3, n * a * 58, n * a)
    res["note"] = ("tracker scan, row moments + fold (one launch), scale, GAE (4 launches) through the Python API; "
                   "58 B per (t, env) incl. the float64 tracker rows; the working set (60 MB) stays in L2")
    out["f3_intrinsic_returns_256x4096"] = res
    return out


def sharded_check(rollout, samples, make_pass, rewards, dones, value_samples, device, world, rank):
    """N > 1, before timing: one pass with the in-kernel peer exchange and one with PPO_PEER_EXCHANGE=0 (NCCL
    all-gather + merge between two launches) must give the same bits, and the merged moments must be the moments of the
    all-gathered raw advantages."""
    import torch
    import torch.distributed as dist
    from ppo_b200 import _lib
    res, outs = {}, {}
    saved = os.environ.get("PPO_PEER_EXCHANGE")
    for name, env in (("peers", "1"), ("nccl", "0")):
        os.environ["PPO_PEER_EXCHANGE"] = env
        rp = make_pass()
        rp.set_samples(samples)
        rp.run(rewards, None, dones, value_samples)
        torch.cuda.synchronize()
        status = 0
        try:
            rp.check_peers()
        except TimeoutError:
            status = 1
        outs[name] = dict(norm=rp.advantages.clone(), merged=rp.global_stats.clone(), raw=rp.advantage.clone(),
                          active=rp.peers is not None, status=status, launches=rp.kernel_launches_per_pass)
        dist.barrier()
        rp.close()
    if saved is None:
        os.environ.pop("PPO_PEER_EXCHANGE", None)
    else:
        os.environ["PPO_PEER_EXCHANGE"] = saved
    pe, nc = outs["peers"], outs["nccl"]
    flags = torch.tensor([int(torch.equal(pe["norm"], nc["norm"])), int(torch.equal(pe["merged"], nc["merged"])),
                          int(pe["active"]), int(pe["status"] == 0 and nc["status"] == 0)], device=device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    # the merged record holds the same bits on every rank (sums are added in rank order everywhere)
    records = torch.empty((world, _lib.STATS_LEN), dtype=torch.float64, device=device)
    dist.all_gather(list(records.unbind(0)), pe["merged"].contiguous())
    identical = all(torch.equal(records[0].view(torch.int64), records[r].view(torch.int64)) for r in range(world))
    # checksums of the normalised advantages of every rank, both paths
    sums = torch.empty((world, 2), dtype=torch.int64, device=device)
    mine = torch.stack([pe["norm"].view(torch.int32).to(torch.int64).sum(), nc["norm"].view(torch.int32).to(torch.int64).sum()])
    dist.all_gather(list(sums.unbind(0)), mine)
    # moments of the all-gathered raw advantages against the merged record
    raw = torch.empty((world,) + tuple(pe["raw"].shape), dtype=torch.float32, device=device)
    dist.all_gather(list(raw.unbind(0)), pe["raw"].contiguous())
    x = raw.double()
    want = [float(x.numel()), float(x.sum().item()), float((x * x).sum().item())]
    got = [float(records[0][i].item()) for i in (_lib.STAT_COUNT, _lib.STAT_ADV_SUM, _lib.STAT_ADV_SUMSQ)]
    rel = max(abs(g - w) / max(abs(w), 1e-30) for g, w in zip(got[1:], want[1:]))
    mean = want[1] / want[0]
    std = max(want[2] / want[0] - mean * mean, 0.0) ** 0.5
    renorm = ((pe["raw"].reshape(-1).double() - mean) / (std + 1e-8)).float()
    norm_err = float((renorm - pe["norm"]).abs().max().item())
    res = {"peers_active": bool(flags[2].item()), "peers_vs_nccl_advantages_bit_equal": bool(flags[0].item()),
           "merged_stats_bit_equal_across_paths": bool(flags[1].item()), "merged_stats_identical_on_all_ranks": identical,
           "advantage_checksums_equal_on_all_ranks": bool((sums[:, 0] == sums[:, 1]).all().item()),
           "count_exact": got[0] == want[0], "moments_vs_allgathered_raw_rel_err": rel,
           "normalised_vs_recomputed_max_abs_err": norm_err, "no_peer_timeout": bool(flags[3].item()),
           "launches_per_pass": {"peers": pe["launches"], "nccl": nc["launches"]}}
    res["ok"] = bool(res["peers_vs_nccl_advantages_bit_equal"] and res["merged_stats_bit_equal_across_paths"]
                     and identical and res["advantage_checksums_equal_on_all_ranks"] and res["count_exact"]
                     and rel < 1e-11 and norm_err < 1e-5 and res["no_peer_timeout"])
    return res


def link_floor(h2d_bytes, d2h_bytes, device, world, iters=5):
    """What the host link alone allows for one e2e step: every rank at once copies as many bytes host -> device and
    device -> host as its e2e step moves (pinned buffers, two streams, both directions together), nothing else."""
    import torch
    import torch.distributed as dist
    chunk = 256 << 20
    hin = torch.empty(min(h2d_bytes, chunk), dtype=torch.uint8).pin_memory()
    hout = torch.empty(min(d2h_bytes, chunk), dtype=torch.uint8).pin_memory()
    din, dout = torch.empty(hin.numel(), dtype=torch.uint8, device=device), torch.empty(hout.numel(), dtype=torch.uint8, device=device)
    s_in, s_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

    def one():
        with torch.cuda.stream(s_in):
            left = h2d_bytes
            while left > 0:
                n = min(left, hin.numel())
                din[:n].copy_(hin[:n], non_blocking=True)
                left -= n
        with torch.cuda.stream(s_out):
            left = d2h_bytes
            while left > 0:
                n = min(left, hout.numel())
                hout[:n].copy_(dout[:n], non_blocking=True)
                left -= n
        s_in.synchronize()
        s_out.synchronize()

    one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        one()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / iters
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt * 1e3


def vtrace_sharded(device, world, total_envs=4096, iters=40):
    """BASELINE config 4 (V-trace 256 x 4096 x 18) with the env axis over the ranks: no exchange at all, every rank
    launches its block; device time per launch, max over ranks."""
    import torch
    import torch.distributed as dist
    from ppo_b200 import _lib
    lib = _lib.load()
    n, act = 256, 18
    b = total_envs // world
    g = torch.Generator(device=device)
    g.manual_seed(77 + (dist.get_rank() if world > 1 else 0))
    sets = max(2, min(8, (400 << 20) // (n * b * act * 8)))  # rotate through more than L2's worth of log-policies
    bufs = []
    for _ in range(sets):
        beh = torch.log_softmax(torch.randn((n, b, act), device=device, generator=g), dim=-1)
        tgt = torch.log_softmax(torch.randn((n, b, act), device=device, generator=g), dim=-1)
        bufs.append((beh, tgt, torch.randint(0, act, (n, b), device=device, generator=g),
                     torch.randint(-1, 3, (n, b), device=device, generator=g).float(),
                     (torch.rand((n, b), device=device, generator=g) < 3 / 101).view(torch.uint8),
                     torch.randn((n + 1, b), device=device, generator=g) * 0.4 + 0.1,
                     [torch.empty((n, b), device=device) for _ in range(3)]))
    stream = torch.cuda.current_stream().cuda_stream

    def launch(i):
        beh, tgt, a, r, d, v, o = bufs[i % sets]
        _lib.check(lib.ppo_vtrace(beh.data_ptr(), tgt.data_ptr(), a.data_ptr(), 8, r.data_ptr(), d.data_ptr(), _lib.DONE_U8,
                                  v.data_ptr(), v[n].data_ptr(), n, b, act, GAMMA, 1.0, o[0].data_ptr(), o[1].data_ptr(),
                                  o[2].data_ptr(), None, stream), "ppo_vtrace")
    for i in range(6):
        launch(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    if world > 1:
        t = torch.tensor([us], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item())
    return {"workload": f"V-trace {n} steps x {total_envs} envs x {act} actions, {b} envs per GPU, no exchange",
            "us_per_launch_max_over_ranks": us, "step_env_per_s": n * total_envs / (us * 1e-6),
            "useful_gbs_per_gpu": n * b * 37 / us / 1e3}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ppo_b200 import _lib, rollout
    from ppo_b200 import returns_truncated as rt
    from ppo_b200 import tvf as tvf_mod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    steps, envs, n_h = STEPS, envs_per_gpu(world, args.scaling), HORIZONS
    hz = tvf_mod.get_value_head_horizons(n_h, MAX_HORIZON)
    assert len(hz) == n_h

    # ---- synthetic rollout of this rank (seed = rank), distributions of the reference's tests (tests/test_tvf.py:17-24)
    g = torch.Generator(device=device)
    g.manual_seed(rank)
    rewards = torch.randint(-1, 3, (steps, envs), device=device, generator=g).float()
    dones = (torch.rand((steps, envs), device=device, generator=g) < 3 / 101).view(torch.uint8)
    value_samples = torch.randn((steps + 1, envs, n_h), device=device, generator=g) * 0.4 + 0.1
    value_samples[:, :, 0] = 0
    # the value estimate GAE uses comes off the value curves inside the pass (rl/rollout.py:1196, rl/tvf.py:303-433)

    samples = rt.draw_n_step_samples("exponential", "advanced", hz, steps, N_STEP, SAMPLES, seed=SAMPLE_SEED)
    rp = rollout.ReturnPass(steps, envs, hz, GAMMA, TVF_GAMMA, LAMBDA_POLICY, LAMBDA_VALUE)
    tabs = rp.set_samples(samples)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing.  Like the reference (rl/returns_truncated.py:85-129 draws a new n-step table in every
    # call) each timed pass draws a fresh sample table, resolves it on the host and uploads it: that host work runs
    # ahead of the GPU (all launches are asynchronous), so it costs GPU time only if it is slower than the pass itself.
    def timed_passes(redraw):
        for _ in range(args.warmup):
            if redraw:
                rp.set_samples(rt.draw_n_step_samples("exponential", "advanced", hz, steps, N_STEP, SAMPLES))
            rp.run(rewards, None, dones, value_samples)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        host_s = 0.0
        t_begin.record()
        for i in range(args.steps):
            if redraw:
                h0 = time.perf_counter()
                rp.set_samples(rt.draw_n_step_samples("exponential", "advanced", hz, steps, N_STEP, SAMPLES))
                host_s += time.perf_counter() - h0
            rp.run(rewards, None, dones, value_samples, tvf_events=evs[i])
        t_end.record()
        barrier()
        total = t_begin.elapsed_time(t_end)
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total / args.steps, [e[0].elapsed_time(e[1]) for e in evs], host_s / args.steps * 1e3

    check = None
    if world > 1:
        check = sharded_check(rollout, samples,
                              lambda: rollout.ReturnPass(steps, envs, hz, GAMMA, TVF_GAMMA, LAMBDA_POLICY, LAMBDA_VALUE),
                              rewards, dones, value_samples, device, world, rank)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    np.random.seed(SAMPLE_SEED + 1 + rank)
    ms_per_step, tvf_ms, redraw_host_ms = timed_passes(redraw=True)
    tabs = rp.set_samples(samples)  # back to the seeded table: the end-to-end leg below must reproduce these targets
    fixed_ms_per_step, fixed_tvf_ms, _ = timed_passes(redraw=False)
    targets_per_step = steps * envs * n_h * world
    value = targets_per_step / (ms_per_step * 1e-3)

    # ---- what sharding adds to a step: the same pass on this rank's envs WITHOUT the exchange (a group of one), so
    # that the wait in front of the in-kernel moment exchange (rank skew + one NVLink round trip) has a number
    exchange = None
    if world > 1:
        solo_groups = [dist.new_group([r]) for r in range(world)]  # (collective: every rank creates every group)
        solo = rollout.ReturnPass(steps, envs, hz, GAMMA, TVF_GAMMA, LAMBDA_POLICY, LAMBDA_VALUE, group=solo_groups[rank])
        solo.set_samples(samples)
        for _ in range(args.warmup):
            solo.run(rewards, None, dones, value_samples)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        for i in range(args.steps):
            solo.run(rewards, None, dones, value_samples, tvf_events=evs[i])
        t1.record()
        barrier()
        solo_step = t0.elapsed_time(t1) / args.steps
        solo_rest = solo_step - sum(e[0].elapsed_time(e[1]) for e in evs) / len(evs)
        shard_rest = fixed_ms_per_step - sum(fixed_tvf_ms) / len(fixed_tvf_ms)
        t = torch.tensor([solo_rest, solo_step], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exchange = {"rest_of_step_us": shard_rest * 1e3, "rest_of_step_unsharded_us": float(t[0].item()) * 1e3,
                    "exchange_wait_us": (shard_rest - float(t[0].item())) * 1e3,
                    "unsharded_ms_per_step": float(t[1].item()),
                    "note": "rest of step = step - target kernel (GAE / TD(lambda) / statistics / normalisation launch and "
                            "launch gaps), fixed sample table, max over ranks; unsharded = the same pass on this rank's "
                            "envs in a group of one; the difference is the wait in front of the in-kernel moment "
                            "exchange (rank skew of the target launch + one NVLink store / poll round trip)"}
        del solo

    # ---- end to end through the public API with pinned host buffers
    host_in = {
        "rewards": rewards.cpu().pin_memory(),
        "dones": dones.view(torch.bool).cpu().pin_memory(), "value_samples": value_samples.cpu().pin_memory(),
    }
    host_out = {
        "advantage": torch.empty((steps, envs), dtype=torch.float32).pin_memory(),
        "ext_returns": torch.empty((steps, envs), dtype=torch.float32).pin_memory(),
        "advantages": torch.empty(steps * envs, dtype=torch.float32).pin_memory(),
        "tvf_returns": torch.empty((steps, envs, n_h), dtype=torch.float32).pin_memory(),
    }
    tvf_kw = dict(value_samples=host_in["value_samples"], horizons=hz, gamma=TVF_GAMMA, mode="advanced",
                  distribution="exponential", n_step=N_STEP, max_samples=SAMPLES, seed=SAMPLE_SEED)

    group = dist.group.WORLD if world > 1 else None  # sharded: the all-reduce of the advantage moments is inside

    def e2e_step():
        return rollout.calculate_returns(host_in["rewards"], None, host_in["dones"], GAMMA,
                                         LAMBDA_POLICY, LAMBDA_VALUE, tvf=tvf_kw, group=group, out=host_out)

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # returns after the results have landed in host_out (stream synchronised inside)
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop()
    h2d = sum(t.numel() * t.element_size() for t in host_in.values()) * world
    d2h = sum(t.numel() * t.element_size() for t in host_out.values()) * world
    floor_ms = link_floor(h2d // world, d2h // world, device, world)
    vt_leg = vtrace_sharded(device, world) if world > 1 else None
    e2e_value = targets_per_step * e2e_steps / e2e_s
    # the e2e results must equal the device-resident ones (same inputs, same sample table)
    same = bool(torch.equal(host_out["tvf_returns"], rp.tvf_returns.cpu()))

    # ---- roofline of the dominant kernel (TVF targets), from CUDA events inside the timed region
    alg_bytes = 4 * (steps + 1) * envs * n_h + 4 * steps * envs * n_h + 5 * steps * envs  # + 4 (N+1) A value estimate, not counted
    tvf_avg_ms = sum(tvf_ms) / len(tvf_ms)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    achieved = alg_bytes / (tvf_avg_ms * 1e-3) / 1e9
    # DRAM bytes per launch from the newest ncu --set full capture of THIS kernel (tools/ncu_traffic.py writes the file
    # from the capture named in it); null when the file describes another kernel than the one that just ran
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "tvf_traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath))
        if rp.tvf_kernel in rec.get("kernel", ""):
            traffic, traffic_src = rec.get("dram_bytes_per_launch"), f"profiles/tvf_traffic.json <- {rec.get('report')}"
    roofline = {"bound": "hbm", "kernel": f"ppo_tvf_targets_ex, {rp.tvf_kernel} kernel (targets + fused last-horizon "
                                          f"statistics + fused rediscounted value estimate)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tvf_avg_ms,
                "kernel_share_of_step": tvf_avg_ms / ms_per_step,
                "units_per_launch": f"{steps} x {envs} x {n_h} targets at 8.06 B/target",
                # SURVEY 8(d) asks for both denominators: the measured copy bandwidth above and the 8 TB/s spec figure
                "frac_of_8tbs_spec": achieved / 8000.0,
                # the whole step (target launch + GAE / TD(lambda) / normalisation at 21 B per (t, env)) against the same peak
                "whole_step_frac": (alg_bytes + 21 * steps * envs) / (ms_per_step * 1e-3) / 1e9 / peak}
    # informational: the same kernel against the shared-memory gather bound that actually limits it:
    # every target gathers 2 x (distinct n-steps per horizon) float32 value samples, one 128-byte wavefront per clock per SM
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    ent = float(tabs.layout.offset_words[15] and tabs.host[int(tabs.layout.offset_words[15]) + 1].item())
    gather_bytes = steps * envs * ent * 2 * 4 if ent > 0 else None
    if gather_bytes:
        smem_peak = sms * 128 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e9
        roofline["shared_memory_gather_bound"] = {
            "gather_bytes_per_launch": gather_bytes, "peak_gbs": smem_peak,
            "floor_ms": gather_bytes / smem_peak / 1e6, "frac_of_floor": gather_bytes / smem_peak / 1e6 / tvf_avg_ms,
            "note": "2 gathers x 4 B per (target, distinct n-step); HBM traffic is 8.06 B per target, so at 8 samples "
                    "per horizon the gathers take longer than the HBM transfers on any shared-memory formulation"}

    other = other_configs(device, peak) if world == 1 and not args.no_other_configs else None
    if rank == 0:
        cpu = cpu_baseline_single() if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(world, args.scaling), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "matches_resident": same,
                    # the host link alone: all ranks at once copy one step's bytes in both directions, nothing else
                    "link_floor_ms": floor_ms, "frac_of_link_floor": floor_ms / (e2e_s / e2e_steps * 1e3),
                    "link_gbs_at_floor": (h2d + d2h) / floor_ms / 1e6,
                    "api": "ppo_b200.rollout.calculate_returns(pinned host tensors, out=pinned host tensors)"},
            "gpu_launches": rp.kernel_launches_per_pass * args.steps,
            "roofline": roofline,
            "tvf_tables": {"entries_per_horizon": int(tabs.n_entries), "nmax": int(tabs.nmax)},
            # `value` / `ms_per_step` include a fresh n-step table per pass (draw + ppo_tvf_tables_build + upload, like
            # the reference's per-call draw); the same passes on one fixed table beside it
            "table_redraw": {"per_pass": True, "host_ms_per_pass": redraw_host_ms,
                             "fixed_table_ms_per_step": fixed_ms_per_step,
                             "fixed_table_value": targets_per_step / (fixed_ms_per_step * 1e-3),
                             "fixed_table_kernel_ms": sum(fixed_tvf_ms) / len(fixed_tvf_ms),
                             "redraw_over_fixed": ms_per_step / fixed_ms_per_step},
        }
        if check is not None:
            line["sharded_check"] = check
        if exchange is not None:
            line["sharded_step"] = exchange
        if vt_leg is not None:
            line["c4_vtrace_sharded"] = vt_leg
        if other is not None:
            line["other_configs"] = other
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 4096 envs per GPU (default); strong: 32768 envs over the whole job (BASELINE config 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~15 s oracle timing (development)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the kernel-level figures of the other BASELINE configs (clean launch lists under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
