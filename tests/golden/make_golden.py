"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [/root/reference]

The reference package cannot be imported as a whole (``rl/__init__.py`` pulls in
``gym``), so the four hot-path files are loaded by path with a stub ``rl``
package.  Nothing here is copied into the repo: only the
seeded inputs and the reference's outputs are stored, as ``*.npz``.  The GPU box
never runs this script and never sees the reference.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ppo_b200 import synthetic  # noqa: E402


def load_reference(root="/root/reference"):
    """Import rl.returns / rl.returns_truncated / rl.vtrace / rl.tvf from the reference checkout."""
    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg = types.ModuleType("rl")
    pkg.__path__ = []
    sys.modules["rl"] = pkg
    cfg = types.ModuleType("rl.config")
    cfg.args = types.SimpleNamespace()
    sys.modules["rl.config"] = cfg
    sys.modules["rl.utils"] = types.ModuleType("rl.utils")
    ro = types.ModuleType("rl.rollout")

    class RunnerModule:  # stand-in for rl/rollout.py:91-113
        def __init__(self, parent):
            self.runner = parent

    ro.RunnerModule = RunnerModule
    sys.modules["rl.rollout"] = ro
    pkg.config, pkg.utils, pkg.rollout = cfg, sys.modules["rl.utils"], ro
    returns = by_path("rl.returns", "rl/returns.py")
    truncated = by_path("rl.returns_truncated", "rl/returns_truncated.py")
    vtrace = by_path("rl.vtrace", "rl/vtrace.py")
    tvf = by_path("rl.tvf", "rl/tvf.py")
    pkg.returns, pkg.returns_truncated, pkg.vtrace, pkg.tvf = returns, truncated, vtrace, tvf
    return types.SimpleNamespace(returns=returns, truncated=truncated, vtrace=vtrace, tvf=tvf)


def tvf_test_inputs(steps, envs, n_req, n_val, seed):
    """Seeded version of the reference's own estimator test data (tests/test_tvf.py:14-24)."""
    rng = np.random.RandomState(seed)
    return {
        'gamma': 0.9997,
        'rewards': rng.randint(-1, 3, size=[steps, envs]).astype('float32'),
        'dones': rng.randint(0, 101, size=[steps, envs]) >= 98,
        'required_horizons': np.geomspace(1, 1024, num=n_req).astype('int32'),
        'value_sample_horizons': np.geomspace(1, 1024, num=n_val).astype('int32') - 1,
        'value_samples': _zero_h0(rng.normal(0.1, 0.4, [steps + 1, envs, n_val]).astype('float32')),
    }


def _zero_h0(vs):
    vs[:, :, 0] *= 0
    return vs


def reference_source(root, rel, name, method=None):
    """Source text of a top-level function / class (or one method of a class) of a reference file that cannot be
    imported as a module (``rl/utils.py``, ``rl/wrappers.py`` and ``rl/rollout.py`` pull in gym, cv2 and removed
    NumPy aliases).  The text is executed as is, never stored."""
    import ast
    import textwrap
    text = open(os.path.join(root, rel)).read()
    for node in ast.parse(text).body:
        if getattr(node, "name", None) != name:
            continue
        if method is None:
            return ast.get_source_segment(text, node)
        for sub in node.body:
            if getattr(sub, "name", None) == method:
                lines = text.splitlines()[sub.lineno - 1:sub.end_lineno]
                return textwrap.dedent("\n".join(lines))
    raise KeyError(f"{name}.{method} not found in {rel}")


def normalization_vectors(root, ref):
    """Return trackers + running moments: RunningMeanStd (rl/utils.py:379-443), Runner.calculate_intrinsic_returns
    (rl/rollout.py:1140-1180) and VecNormalizeRewardWrapper (rl/wrappers.py:795-919), run from the reference's own
    source text with stand-in runner / env objects."""
    import math
    import typing
    out = {}
    ns = {"np": np}
    exec(reference_source(root, "rl/utils.py", "update_mean_var_count_from_moments"), ns)
    exec(reference_source(root, "rl/utils.py", "RunningMeanStd"), ns)
    rms_cls = ns["RunningMeanStd"]

    # ---- RunningMeanStd on its own
    rng = np.random.RandomState(5)
    rms = rms_cls(shape=())
    batches = [rng.normal(3.0, 2.0, size=n) for n in (1, 7, 64, 5)]
    states = []
    for b in batches:
        rms.update(b)
        states.append([rms.mean, rms.var, rms.count])
    out["rms_in_batches"] = np.concatenate(batches)
    out["rms_in_sizes"] = np.asarray([len(b) for b in batches])
    out["rms_states"] = np.asarray(states, dtype=np.float64)

    # ---- Runner.calculate_intrinsic_returns over two consecutive rollouts (the carry and the moments persist)
    steps, envs = 40, 24
    src = reference_source(root, "rl/rollout.py", "Runner", "calculate_intrinsic_returns")
    for tag, normalize, center, propagation in [("norm", True, False, False), ("center", True, True, False),
                                                ("prop", True, False, True), ("raw", False, True, False)]:
        args = types.SimpleNamespace(use_intrinsic_rewards=True, gamma_int=0.99, lambda_policy=0.95,
                                     ir=types.SimpleNamespace(normalize=normalize, center=center,
                                                              propagation=propagation))
        fns = {"np": np, "args": args, "gae": ref.returns.gae}
        exec(src, fns)
        runner = types.SimpleNamespace(N=steps, intrinsic_returns_rms=rms_cls(shape=()), ems_norm=np.zeros([envs]),
                                       intrinsic_reward_norm_scale=1, int_returns=np.zeros([steps, envs], np.float32),
                                       prev_obs=np.zeros([steps, envs, 1]))
        for roll in range(2):
            rng = np.random.RandomState(100 + roll)
            runner.int_rewards = np.abs(rng.normal(0.2, 0.3, [steps, envs])).astype(np.float32)
            runner.terminals = rng.random_sample([steps, envs]) < 0.05
            runner.int_value = rng.normal(0.5, 0.4, [steps + 1, envs]).astype(np.float32)
            if tag == "norm":
                out[f"ir_in_rewards_{roll}"], out[f"ir_in_terminals_{roll}"] = runner.int_rewards, runner.terminals
                out[f"ir_in_value_{roll}"] = runner.int_value
            adv = fns["calculate_intrinsic_returns"](runner)
            out[f"ir_{tag}_adv_{roll}"] = np.asarray(adv)
            out[f"ir_{tag}_returns_{roll}"] = runner.int_returns.copy()
            out[f"ir_{tag}_rewards_{roll}"] = np.asarray(runner.int_rewards)
            out[f"ir_{tag}_scale_{roll}"] = np.asarray(runner.intrinsic_reward_norm_scale, dtype=np.float64)
            out[f"ir_{tag}_ems_{roll}"] = runner.ems_norm.copy()
            r = runner.intrinsic_returns_rms
            out[f"ir_{tag}_rms_{roll}"] = np.asarray([r.mean, r.var, r.count], dtype=np.float64)

    # ---- VecNormalizeRewardWrapper.step over a rollout, fed by a stand-in vector env
    gym = types.SimpleNamespace(Wrapper=type("Wrapper", (), {"__init__": lambda self, env: setattr(self, "env", env)}),
                                Env=object)
    wns = {"np": np, "math": math, "gym": gym, "utils": types.SimpleNamespace(RunningMeanStd=rms_cls),
           "VectorEnv": object, "Optional": typing.Optional, "EpisodicDiscounting": None}
    exec(reference_source(root, "rl/wrappers.py", "EpisodicDiscounting"), wns)  # (its static normalisation constants)
    exec(reference_source(root, "rl/wrappers.py", "VecNormalizeRewardWrapper"), wns)
    rng = np.random.RandomState(9)
    steps, envs = 50, 16
    rewards = (rng.randint(-1, 3, size=[steps, envs]) * rng.choice([1.0, 30.0], size=[steps, envs], p=[0.9, 0.1]))
    rewards = rewards.astype(np.float32)
    dones = rng.random_sample([steps, envs]) < 0.04
    out["vn_in_rewards"], out["vn_in_dones"] = rewards, dones
    # env times for the episodic-discounting normalisers (rl/wrappers.py:870-874): steps since the env's last reset
    times = np.zeros([steps, envs], dtype=np.int64)
    clock = rng.randint(0, 40, size=envs)
    for t in range(steps):
        times[t] = clock
        clock = np.where(dones[t], 0, clock + 1)
    out["vn_in_times"] = times

    class Env:
        num_envs = envs

        def __init__(self):
            self.t = 0

        def step(self, actions):
            t = self.t
            self.t += 1
            return None, rewards[t], dones[t], [{"time": int(times[t, a])} for a in range(envs)]

    for tag, kwargs in [("default", dict(gamma=0.99)), ("noclip", dict(gamma=0.999, clip=-1, scale=0.5)),
                        ("tight", dict(gamma=0.9, clip=1.0)),
                        ("ema", dict(gamma=0.99, mode="ema", ema_horizon=200.0)),
                        ("custom", dict(gamma=0.99, mode="custom", clip=3.0)),
                        ("edquad", dict(gamma=0.99, ed_type="quadratic")),
                        ("edharm", dict(gamma=0.99, ed_type="harmonic", ed_bias=2.0, mode="ema", ema_horizon=300.0)),
                        ("edpower", dict(gamma=0.995, ed_type="power", returns_transform=lambda x: x * 0.5))]:
        w = wns["VecNormalizeRewardWrapper"](Env(), **kwargs)
        if tag == "custom":
            w.ret_var = 6.25  # "custom requires setting of ret_std" (rl/wrappers.py:821)
        rows, clips = [], 0
        for t in range(steps):
            _, scaled, _, infos = w.step(None)
            rows.append(np.asarray(scaled, dtype=np.float64))
            clips += int(infos[0].get("reward_clips", 0))
        out[f"vn_{tag}_scaled"] = np.stack(rows)
        out[f"vn_{tag}_clips"] = np.asarray(clips)
        out[f"vn_{tag}_returns"] = np.asarray(w.current_returns, dtype=np.float64)
        out[f"vn_{tag}_rms"] = np.asarray([w.ret_rms.mean, w.ret_rms.var, w.ret_rms.count], dtype=np.float64)
        out[f"vn_{tag}_ret_var"] = np.asarray(w.ret_var, dtype=np.float64)

    # ---- ReturnTracker.update (rl/wrappers.py:737-755) over a rollout with per-step env masks (two env types, as the
    # multi-env reward normaliser :946-971 would call it; one step has no env of the type at all)
    tns = {"np": np, "utils": types.SimpleNamespace(RunningMeanStd=rms_cls)}
    exec(reference_source(root, "rl/wrappers.py", "ReturnTracker"), tns)
    rng = np.random.RandomState(21)
    steps, envs = 30, 12
    rt_rewards = (rng.randint(-1, 3, size=[steps, envs]) * rng.choice([1.0, 10.0], size=[steps, envs])).astype(np.float32)
    rt_dones = rng.random_sample([steps, envs]) < 0.08
    rt_mask = rng.random_sample([steps, envs]) < 0.6
    rt_mask[7] = False
    out["rt_in_rewards"], out["rt_in_dones"], out["rt_in_mask"] = rt_rewards, rt_dones, rt_mask
    tracker = tns["ReturnTracker"](envs, 0.99)
    rows, states = [], []
    for t in range(steps):
        tracker.update(rt_rewards[t], rt_dones[t], rt_mask[t])
        rows.append(tracker.current_returns.copy())
        states.append([tracker.ret_rms.mean, tracker.ret_rms.var, tracker.ret_rms.count])
    out["rt_returns"] = np.stack(rows)                       # float32, after every step
    out["rt_states"] = np.asarray(states, dtype=np.float64)  # running moments after every step

    # ---- TVFRunnerModule.on_train_value_minibatch (rl/tvf.py:32-77): per-sample loss and its gradient (autograd)
    import torch
    import rl.config as cfg
    hz, hw = ref.tvf.get_value_head_horizons(24, 1000, include_weight=True)
    rng = np.random.RandomState(13)
    batch = 19
    targets = rng.normal(1.0, 2.0, [batch, len(hz)]).astype(np.float32)
    preds = rng.normal(1.0, 2.0, [batch, len(hz), 2]).astype(np.float32)
    out["loss_in_horizons"], out["loss_in_weights"] = hz, hw
    out["loss_in_targets"], out["loss_in_predictions"] = targets, preds
    log = types.SimpleNamespace(watch_mean=lambda *a, **k: None)
    runner = types.SimpleNamespace(ext_rewards=np.zeros([2, 3], np.float32), value_heads=["ext", "int"],
                                   tvf_horizons=list(hz), tvf_weights=hw, log=log,
                                   model=types.SimpleNamespace(device="cpu"))
    mod = ref.tvf.TVFRunnerModule(runner)
    for tag, weighting, required, single in [("all", "off", None, None), ("hweighted", "h_weighted", None, None),
                                             ("single", "h_weighted", [5], 5), ("prefix", "off", list(range(8)), -7)]:
        cfg.args.tvf = types.SimpleNamespace(coef=0.7, head_weighting=weighting, max_horizon=1000, horizon_dropout=0.0)
        cfg.args.tvf_return_n_step = 20
        cfg.args.value_opt = types.SimpleNamespace(epochs=2)
        ref.tvf.args = cfg.args
        sel = slice(None) if required is None else required
        p = torch.tensor(preds[:, sel, :], requires_grad=True)
        loss = mod.on_train_value_minibatch({"tvf_value": p}, {"tvf_returns": torch.from_numpy(targets)},
                                            required_tvf_heads=required, single_value_head=single)
        (loss * 0.25).mean().backward()  # rl/rollout.py:1553-1555 with loss_scale = 0.25
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"loss_{tag}_grad"] = p.grad.numpy()
    return out


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    ref = load_reference(root)
    out = {}

    # ---- GAE / TD(lambda) / bootstrapped returns (rl/returns.py)
    r, d, v = synthetic.rollout(24, 12, seed=11)
    out["gae_in_rewards"], out["gae_in_dones"], out["gae_in_values"] = r, d, v
    out["gae_bool"] = ref.returns.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    out["gae_f32done"] = ref.returns.gae(r, v[:-1], v[-1], d.astype(np.float32), 0.999, 0.95)
    out["gae_noterm"] = ref.returns.gae(r, v[:-1], v[-1], None, 0.99, 0.9)
    out["tdl_bool"] = ref.returns.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.8)
    out["boot_scalar"] = ref.returns.calculate_bootstrapped_returns(r, d, v[-1], 0.999)
    garr = np.random.RandomState(5).uniform(0.9, 1.0, size=r.shape).astype(np.float32)
    out["boot_in_gamma_arr"] = garr
    out["boot_array"] = ref.returns.calculate_bootstrapped_returns(r, d, v[-1], garr)
    # known answer from rl/unit_tests.py:203-210
    kr = np.asarray([1, 0, 2, 4, 6], dtype=np.float32)[:, None]
    kd = np.asarray([0, 0, 1, 0, 0], dtype=np.float32)[:, None]
    kv = np.asarray([0, 0.5, 0.5, 3, 4], dtype=np.float32)[:, None]
    out["gae_kat"] = ref.returns.gae(kr, kv, np.asarray(5, dtype=np.float32), kd, gamma=0.5, lamb=1.0)

    # ---- advantage normalisation: rl/rollout.py:1887-1897 is inline trainer code that cannot be
    # imported (rl/rollout.py needs gym and removed NumPy aliases); the three lines are evaluated here.
    adv = out["gae_bool"]
    flat = adv.reshape(-1).copy()
    out["advnorm"] = (flat - flat.mean()) / (flat.std() + 1e-8)
    out["advnorm_clip"] = np.clip(out["advnorm"], -1.5, +1.5)

    # ---- TVF estimator (rl/returns_truncated.py), the reference's own test shapes scaled down
    args = tvf_test_inputs(48, 6, 16, 32, seed=3)
    for key, val in args.items():
        out["tvf_in_" + key] = np.asarray(val)
    fast = ref.truncated._calculate_sampled_return_multi_fast
    slow = ref.truncated._calculate_sampled_return_multi_reference
    for n in (1, 8, 48):
        out[f"tvf_fast_n{n}"] = fast(n_step_list=[n], **args)
        out[f"tvf_slow_n{n}"] = slow(n_step_list=[n], **args)
    lamb = 1 - 1 / 20
    w = np.asarray([lamb ** x for x in range(48)], dtype=np.float32)
    samples = np.random.RandomState(9).choice(range(1, 49), [16, 8], replace=True, p=w / w.sum())
    out["tvf_in_samples"] = samples
    out["tvf_fast_samples"] = fast(n_step_samples=samples, **args)
    out["tvf_slow_samples"] = slow(n_step_samples=samples, **args)
    out["tvf_fast_samples_log"] = fast(n_step_samples=samples, use_log_interpolation=True, **args)
    for dist, mode in [("fixed", "standard"), ("exponential", "standard"), ("exponential", "advanced"),
                       ("uniform", "advanced"), ("hyperbolic", "clipped"), ("quadratic", "adaptive"),
                       ("exponential", "mcx"), ("exponential", "full")]:
        out[f"tvf_gre_{dist}_{mode}"] = ref.truncated.get_return_estimate(
            dist, mode, n_step=6, max_samples=5, seed=17, **args)

    # production-like horizons: required == value horizons == get_value_head_horizons (rl/tvf.py:250-262)
    hz = ref.tvf.get_value_head_horizons(32, 1000)
    r2, d2, _ = synthetic.rollout(40, 5, seed=21)
    vs2 = synthetic.value_samples(40, 5, len(hz), seed=21)
    out["prod_in_horizons"], out["prod_in_rewards"], out["prod_in_dones"], out["prod_in_vs"] = hz, r2, d2, vs2
    out["prod_adv"] = ref.truncated.get_return_estimate(
        "exponential", "advanced", 0.99997, r2, d2, hz, hz, vs2, n_step=20, max_samples=8, seed=1)
    out["prod_fixed"] = ref.truncated.get_return_estimate(
        "fixed", "advanced", 0.99997, r2, d2, hz, hz, vs2, n_step=40, max_samples=8, seed=1)
    out["prod_adv_log"] = ref.truncated.get_return_estimate(
        "exponential", "advanced", 0.99997, r2, d2, hz, hz, vs2, n_step=20, max_samples=8,
        use_log_interpolation=True, seed=1)

    # ---- rl/tvf.py free functions
    out["heads_128_30000"] = ref.tvf.get_value_head_horizons(128, 30000)
    hw, cw = ref.tvf.get_value_head_horizons(16, 100, include_weight=True)
    out["heads_16_100"], out["heads_16_100_w"] = hw, cw
    out["heads_lin_9_64"] = ref.tvf.get_value_head_horizons(9, 64, spacing="linear")
    kh = np.asarray([0, 1, 2, 10, 100])
    kvals = np.asarray([0, 5, 10, -1, 2])[None, :].repeat(11, axis=0)
    ktgt = np.asarray([-100, -1, 0, 1, 2, 3, 4, 99, 100, 101, 200])
    out["hinterp_kat"] = ref.tvf.horizon_interpolate(kh, kvals, ktgt)  # tests/test_tvf.py:121-129
    rng = np.random.RandomState(31)
    hv = rng.normal(0, 1, [7, 9, len(hz)]).astype(np.float32)
    ht = rng.randint(-5, 1100, size=[7, 9])
    out["hinterp_in_values"], out["hinterp_in_targets"] = hv, ht
    out["hinterp"] = ref.tvf.horizon_interpolate(hz, hv, ht)
    rv = np.cumsum(np.abs(rng.normal(0, 1, [20, len(hz)])), axis=1).astype(np.float32)
    rv[:, 0] = 0
    out["redis_in_values"] = rv
    out["redis_down"] = ref.tvf.get_rediscounted_value_estimate(rv, 0.9999, 0.999, hz)
    out["redis_up"] = ref.tvf.get_rediscounted_value_estimate(rv, 0.99, 0.9999, hz)
    out["redis_same"] = np.asarray(ref.tvf.get_rediscounted_value_estimate(rv, 0.99, 0.99, hz))

    # ---- TVFRunnerModule.trim_horizons (rl/tvf.py:91-208) with a stand-in runner and config
    import types as _types
    import rl.config as _cfg
    hz_t = ref.tvf.get_value_head_horizons(32, 1000)
    _cfg.args.env = _types.SimpleNamespace(timeout=900)
    _cfg.args.tvf = _types.SimpleNamespace(eta_percentile=90, eta_buffer=32, eta_minh=128, trim_clip=-1)
    ref.tvf.args = _cfg.args
    log = _types.SimpleNamespace(watch_stats=lambda *a, **k: None, watch_mean=lambda *a, **k: None)
    n_envs_t = 37
    runner = _types.SimpleNamespace(ext_rewards=np.zeros([4, n_envs_t], np.float32), value_heads=["ext", "int"],
                                    tvf_horizons=list(hz_t), K=len(hz_t), log=log,
                                    episode_length_buffer=[300, 420, 510, 640, 880, 200])
    mod = ref.tvf.TVFRunnerModule(runner)
    rng = np.random.RandomState(77)
    tv = np.cumsum(np.abs(rng.normal(0, 1, [n_envs_t, len(hz_t), 2])), axis=1).astype(np.float32)
    tt = rng.randint(0, 1000, size=[n_envs_t])
    tt[:4] = [0, 899, 900, 950]  # nothing trimmed, one step left, at the limit, past the limit
    out["trim_in_horizons"], out["trim_in_values"], out["trim_in_time"] = hz_t, tv, tt
    out["trim_in_ep_lengths"] = np.asarray(runner.episode_length_buffer)
    for method in ("timelimit", "est_term"):
        for mode in ("interpolate", "average", "substitute"):
            new, fin, ttt = mod.trim_horizons(tv, tt, method=method, mode=mode)
            out[f"trim_{method}_{mode}_values"], out[f"trim_{method}_{mode}_final"] = new, fin
            out[f"trim_{method}_{mode}_ttt"] = np.asarray(ttt)
    _cfg.args.tvf.trim_clip = 0.25
    new, fin, _ = mod.trim_horizons(tv, tt, method="timelimit", mode="average")
    out["trim_clip_values"], out["trim_clip_final"] = new, fin
    _cfg.args.tvf.trim_clip = -1
    np.random.seed(123)
    new, fin, _ = mod.trim_horizons(tv, tt, method="timelimit", mode="random")
    out["trim_random_values"], out["trim_random_final"] = new, fin
    new, fin, ttt = mod.trim_horizons(tv, tt, method="off", mode="average")
    out["trim_off_values"] = new
    assert fin == 0 and ttt is None

    # ---- V-trace (rl/vtrace.py)
    beh, tgt, act = synthetic.policies(14, 6, n_actions=5, seed=41)
    r3, d3, v3 = synthetic.rollout(14, 6, seed=41)
    out["vt_in_beh"], out["vt_in_tgt"], out["vt_in_act"] = beh, tgt, act
    out["vt_in_rewards"], out["vt_in_dones"], out["vt_in_values"] = r3, d3, v3
    vs, adv, cs = ref.vtrace.importance_sampling_v_trace(beh, tgt, act, r3, d3, v3[:-1], v3[-1], 0.99, 0.9)
    out["vt_vs"], out["vt_adv"], out["vt_cs"] = vs, adv, cs
    vs, adv, cs = ref.vtrace.importance_sampling_v_trace(beh, beh, act, r3, d3, v3[:-1], v3[-1], 0.99, 1.0)
    out["vt_onpolicy_vs"], out["vt_onpolicy_adv"] = vs, adv
    out["vt_trust"] = ref.vtrace.v_trace_trust_region(beh, tgt)

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")

    norm = normalization_vectors(root, ref)
    path = os.path.join(HERE, "normalization_vectors.npz")
    np.savez_compressed(path, **norm)
    print(f"wrote {path}: {len(norm)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
