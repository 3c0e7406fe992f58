"""GPU parity of the value estimates read off the truncated value curves (rl/tvf.py:303-433): standalone
rediscounting kernels against the golden vectors and the oracle, and the estimate fused into the target launch."""
import types

import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import _lib, returns, rollout, synthetic
from ppo_b200 import returns_truncated as rt
from ppo_b200 import tvf
from util import TOL, rel_err

pytestmark = pytest.mark.gpu


def test_rediscount_golden(golden):
    hz, rv = golden["prod_in_horizons"], golden["redis_in_values"]
    got = tvf.get_rediscounted_value_estimate(rv, 0.9999, 0.999, hz)
    assert got.dtype == np.float32 and rel_err(got, golden["redis_down"]) <= TOL
    assert rel_err(tvf.get_rediscounted_value_estimate(rv, 0.99, 0.9999, hz), golden["redis_up"]) <= TOL
    assert np.array_equal(tvf.get_rediscounted_value_estimate(rv, 0.99, 0.99, hz), golden["redis_same"])
    t = tvf.get_rediscounted_value_estimate(torch.from_numpy(rv).cuda(), 0.9999, 0.999, hz)
    assert isinstance(t, torch.Tensor) and t.is_cuda and rel_err(t.cpu().numpy(), golden["redis_down"]) <= TOL
    with pytest.raises(AssertionError):  # rl/tvf.py:409 `first horizon must be 0`
        tvf.get_rediscounted_value_estimate(rv[:, 1:], 0.99, 0.999, hz[1:])


@pytest.mark.parametrize("rows,heads,max_h", [(3000, 128, 30000), (17, 16, 200), (1, 8, 50), (0, 8, 50)])
def test_rediscount_vs_oracle(rows, heads, max_h):
    hz = oracle.get_value_head_horizons(heads, max_h)
    rng = np.random.RandomState(rows + heads)
    v = np.cumsum(np.abs(rng.normal(0, 1, [rows, len(hz)])), axis=1).astype(np.float32)
    v[:, 0] = 0
    for old, new in ((0.99997, 0.999), (0.99, 0.9999)):
        ref = oracle.get_rediscounted_value_estimate(v, old, new, hz)
        assert rel_err(tvf.get_rediscounted_value_estimate(v, old, new, hz), ref) <= TOL
        ref = oracle.rediscount_horizons(v, old, new, hz)
        got = tvf.rediscount_horizons(v, old, new, hz)
        assert got.shape == ref.shape and rel_err(got, ref) <= TOL
    assert tvf.rediscount_horizons(v, 0.99, 0.99, hz) is v


def test_rediscount_weights_are_the_reference_sum():
    """sum_i (V_i - V_{i-1}) rho_i == sum_j w_j V_j in exact arithmetic (float64 check of the re-association)."""
    hz = oracle.get_value_head_horizons(128, 30000)
    rho = tvf.rediscount_ratios(0.99997, 0.999, [int(h) for h in hz])
    w = tvf.rediscount_weights(0.99997, 0.999, [int(h) for h in hz]).astype(np.float64)
    v = np.random.RandomState(0).normal(size=[50, len(hz)])
    ref = ((v[:, 1:] - v[:, :-1]) * rho[1:]).sum(axis=1)
    assert np.abs(v @ w - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("gamma,path", [(0.999, rt.PATH_TSLAB), (0.99997, rt.PATH_TSLAB), (0.999, rt.PATH_RING),
                                        (0.99997, rt.PATH_GENERIC)])
def test_value_estimate_fused_into_the_target_launch(gamma, path):
    """ppo_tvf_targets_ex: targets unchanged, value estimate = rediscounted curve (gamma != tvf gamma) or the last
    horizon (same gamma); on the transposed slab kernel it comes out of the same launch (flag 0x200)."""
    steps, envs, tvf_gamma = 96, 210, 0.99997
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=8)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=9)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=2)
    out = torch.empty((steps + 1, envs), dtype=torch.float32, device="cuda")
    w = None if gamma == tvf_gamma else tvf.rediscount_weights(tvf_gamma, gamma, [int(h) for h in hz])
    got = rt._launch(tvf_gamma, r, d, hz, hz, vs, table, False, path=path, value_out=out, value_weights=w)
    last = _lib.load().ppo_tvf_last_path()
    assert (last & 0xff) == path and bool(last & 0x200) == (path == rt.PATH_TSLAB)
    assert rel_err(got, oracle.sampled_return_multi(tvf_gamma, r, d, hz, hz, vs, n_step_samples=table)) <= TOL
    flat = vs.reshape(-1, len(hz))
    ref = oracle.get_rediscounted_value_estimate(flat, tvf_gamma, gamma, hz).reshape(steps + 1, envs)
    if gamma == tvf_gamma:
        assert np.array_equal(out.cpu().numpy(), vs[:, :, -1])
    else:
        assert rel_err(out.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("gamma", [0.999, 0.99997])
def test_return_pass_with_values_from_the_curves(gamma):
    """rl/rollout.py:1196 data flow: GAE reads the (rediscounted) estimate taken off the value curves."""
    steps, envs, tvf_gamma = 128, 300, 0.99997
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=3)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=4)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    rp = rollout.ReturnPass(steps, envs, hz, gamma, tvf_gamma, 0.95, 0.9)
    rp.fuse_gae = True  # (opt-in: GAE, TD(lambda) and their statistics inside the target launch)
    rp.set_samples(table)
    rp.run(torch.from_numpy(r).cuda(), None, torch.from_numpy(d).cuda().view(torch.uint8), torch.from_numpy(vs).cuda())
    torch.cuda.synchronize()
    # targets (+ value estimate, GAE, TD(lambda) and both statistics records as by-products), normalisation
    assert rp.tvf_kernel == "tslab" and rp.kernel_launches_per_pass == 2 and rp.gae_fused
    values = oracle.get_rediscounted_value_estimate(vs.reshape(-1, len(hz)), tvf_gamma, gamma, hz).reshape(steps + 1, envs)
    assert rel_err(rp.value_estimates.cpu().numpy(), values) <= TOL
    got_v = rp.value_estimates.cpu().numpy()  # GAE parity on the estimate the device actually used
    assert rel_err(rp.advantage.cpu().numpy(), oracle.gae(r, got_v[:-1], got_v[-1], d, gamma, 0.95)) <= TOL
    assert rel_err(rp.ext_returns.cpu().numpy(), oracle.td_lambda(r, got_v[:-1], got_v[-1], d, gamma, 0.9)) <= TOL
    ref = oracle.sampled_return_multi(tvf_gamma, r, d, hz, hz, vs, n_step_samples=table)
    assert rel_err(rp.tvf_returns.cpu().numpy(), ref) <= TOL
    # the fused GAE stage against the stand-alone kernels on the same estimate: outputs within one float32 ulp (the
    # float64 carries are composed in a different order), statistics of exactly these outputs
    adv, ret = rp.advantage.cpu().numpy(), rp.ext_returns.cpu().numpy()
    ref_adv = returns.gae(r, got_v[:-1], got_v[-1], d, gamma, 0.95)
    ref_ret = returns.td_lambda(r, got_v[:-1], got_v[-1], d, gamma, 0.9)
    assert np.abs(adv - ref_adv).max() <= np.spacing(np.abs(ref_adv).max()) and np.abs(ret - ref_ret).max() <= np.spacing(np.abs(ref_ret).max())
    st = rp.stats.cpu().numpy()
    a64, r64, v64 = adv.astype(np.float64), ret.astype(np.float64), got_v.astype(np.float64)
    assert st[0] == adv.size and st[3] == adv.min() and st[4] == adv.max() and st[7] == ret.min() and st[8] == ret.max()
    assert st[11] == got_v.size and st[12] == np.count_nonzero(adv == 0) and st[13] == np.count_nonzero(ret == 0)
    np.testing.assert_allclose([st[1], st[2], st[5], st[6], st[9], st[10]],
                               [a64.sum(), (a64 ** 2).sum(), r64.sum(), (r64 ** 2).sum(), v64.sum(), (v64 ** 2).sum()], rtol=1e-11)
    assert rel_err(rp.advantages.cpu().numpy(), oracle.normalize_advantages(adv)) <= TOL
    # ... and with the fusion switched off (separate GAE launch) the pass gives the same results within that ulp
    rp2 = rollout.ReturnPass(steps, envs, hz, gamma, tvf_gamma, 0.95, 0.9)
    rp2.fuse_gae = False
    rp2.set_samples(table)
    rp2.run(torch.from_numpy(r).cuda(), None, torch.from_numpy(d).cuda().view(torch.uint8), torch.from_numpy(vs).cuda())
    torch.cuda.synchronize()
    assert not rp2.gae_fused and rp2.kernel_launches_per_pass == 2 and torch.equal(rp2.tvf_returns, rp.tvf_returns)
    assert np.abs(rp2.advantage.cpu().numpy() - adv).max() <= np.spacing(np.abs(adv).max())
    # the same through calculate_returns, NumPy buffers and pinned host buffers (env-blocked pipeline)
    kw = dict(value_samples=vs, horizons=hz, gamma=tvf_gamma, mode="advanced", distribution="exponential", n_step=20,
              max_samples=8, seed=1)
    res = rollout.calculate_returns(r, None, d, gamma, 0.95, 0.9, tvf=kw)
    assert np.abs(res["advantage"] - adv).max() <= np.spacing(np.abs(adv).max())
    assert np.array_equal(res["tvf_returns"], rp.tvf_returns.cpu().numpy())
    pinned = dict(kw, value_samples=torch.from_numpy(vs).pin_memory())
    out = {"tvf_returns": torch.empty((steps, envs, len(hz))).pin_memory(), "advantage": torch.empty((steps, envs)).pin_memory()}
    res2 = rollout.calculate_returns(torch.from_numpy(r).pin_memory(), None, torch.from_numpy(d).pin_memory(), gamma,
                                     0.95, 0.9, tvf=pinned, out=out)
    assert torch.equal(out["advantage"], torch.from_numpy(res["advantage"])) and torch.equal(out["tvf_returns"], rp.tvf_returns.cpu())
    # (the pass without the fusion and calculate_returns take the same kernel -- bit-identical estimates; the fused pass
    # always runs on the transposed slab kernel, and the quad slab kernel (PPO_TVF_QSLAB=1) adds the 128 products of a
    # row in another order)
    assert torch.equal(res2["value_estimates"].cpu(), rp2.value_estimates.cpu())
    assert rel_err(res2["value_estimates"].cpu().numpy(), rp.value_estimates.cpu().numpy()) <= 1e-6


def test_runner_module_value_estimate_and_returns():
    """TVFRunnerModule with a stand-in runner: the rediscounted branch of get_tvf_ext_value_estimate
    (rl/tvf.py:342-360) and calculate_tvf_returns (rl/tvf.py:210-271)."""
    steps, envs = 24, 12
    hz = oracle.get_value_head_horizons(32, 1000)
    r, d, _ = synthetic.rollout(steps, envs, seed=12)
    log = types.SimpleNamespace(watch_mean=lambda *a, **k: None)
    runner = types.SimpleNamespace(ext_rewards=r, terminals=d, all_obs=np.zeros([steps + 1, envs, 2], np.float32),
                                   tvf_horizons=list(hz), value_heads=["ext", "int"], N=steps, log=log)
    args = types.SimpleNamespace(
        gamma=0.999, tvf_return_n_step=8,
        tvf=types.SimpleNamespace(enabled=True, gamma=0.9999, trimming_mode="off", trim_advantages="untrimmed",
                                  return_mode="advanced", return_distribution="exponential", return_samples=4,
                                  return_use_log_interpolation=False))
    mod = tvf.TVFRunnerModule(runner, args)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=13)
    mod.tvf_value[..., 0] = vs
    mod.tvf_untrimmed_value[..., 0] = vs
    same = mod.get_tvf_ext_value_estimate(0.9999)
    assert np.array_equal(same, vs[:, :, -1])
    got = mod.get_tvf_ext_value_estimate(0.999)
    ref = oracle.get_rediscounted_value_estimate(vs.reshape(-1, len(hz)), 0.9999, 0.999, hz).reshape(steps + 1, envs)
    assert got.shape == (steps + 1, envs) and rel_err(got, ref) <= TOL
    np.random.seed(5)
    targets = mod.calculate_tvf_returns()
    np.random.seed(5)
    ref = oracle.get_return_estimate("exponential", "advanced", 0.9999, r, d, hz, hz, vs, n_step=8, max_samples=4)
    assert targets.shape == (steps, envs, len(hz)) and rel_err(targets, ref) <= TOL
    assert rel_err(mod.rediscount_horizons(vs.reshape(-1, len(hz))),
                   oracle.rediscount_horizons(vs.reshape(-1, len(hz)), 0.9999, 0.999, hz)) <= TOL


def test_horizon_interpolate_golden_and_oracle(golden):
    hz = golden["prod_in_horizons"]
    got = tvf.horizon_interpolate(hz, golden["hinterp_in_values"], golden["hinterp_in_targets"])
    assert got.dtype == np.float64 and got.shape == golden["hinterp"].shape
    assert np.abs(got - golden["hinterp"]).max() <= 1e-12 * max(np.abs(golden["hinterp"]).max(), 1.0)
    # the log-scaled axis of trim_horizons (rl/tvf.py:143-152), targets beyond both ends of the axis, exact hits
    rng = np.random.RandomState(2)
    v = rng.normal(size=[5000, len(hz)]).astype(np.float32)
    t = rng.uniform(-10, 40000, size=[5000])
    t[:len(hz)] = hz
    scale = lambda x: np.log10(10 + np.maximum(x, -9)) - 1  # noqa: E731
    ref = oracle.horizon_interpolate(scale(hz.astype(np.float64)), v, scale(t))
    got = tvf.horizon_interpolate(scale(hz.astype(np.float64)), v, scale(t))
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    with pytest.raises(AssertionError):
        tvf.horizon_interpolate(hz[1:], v[:, 1:], t)  # first horizon must be 0


def _trim_module(golden, clip=-1):
    hz = golden["trim_in_horizons"]
    log = types.SimpleNamespace(watch_stats=lambda *a, **k: None, watch_mean=lambda *a, **k: None)
    runner = types.SimpleNamespace(ext_rewards=np.zeros([4, 37], np.float32), terminals=None, all_obs=None,
                                   tvf_horizons=list(hz), value_heads=["ext", "int"], N=4, K=len(hz), log=log,
                                   episode_length_buffer=list(golden["trim_in_ep_lengths"]))
    args = types.SimpleNamespace(env=types.SimpleNamespace(timeout=900),
                                 tvf=types.SimpleNamespace(eta_percentile=90, eta_buffer=32, eta_minh=128, trim_clip=clip))
    return tvf.TVFRunnerModule(runner, args)


def test_trim_horizons_golden(golden):
    """TVFRunnerModule.trim_horizons (rl/tvf.py:91-208) against outputs of the real reference, every method / mode."""
    g = golden
    tv, tt = g["trim_in_values"], g["trim_in_time"]
    mod = _trim_module(g)
    for method in ("timelimit", "est_term"):
        for mode in ("interpolate", "average", "substitute"):
            new, fin, ttt = mod.trim_horizons(tv, tt, method=method, mode=mode)
            ref = g[f"trim_{method}_{mode}_values"]
            assert new.dtype == np.float32 and new.shape == ref.shape
            assert rel_err(new, ref) <= TOL and rel_err(fin, g[f"trim_{method}_{mode}_final"]) <= TOL
            assert np.array_equal(np.asarray(ttt), g[f"trim_{method}_{mode}_ttt"])  # integer work: exact
    new, fin, _ = _trim_module(g, clip=0.25).trim_horizons(tv, tt, method="timelimit", mode="average")
    assert rel_err(new, g["trim_clip_values"]) <= TOL and rel_err(fin, g["trim_clip_final"]) <= TOL
    np.random.seed(123)
    new, fin, _ = mod.trim_horizons(tv, tt, method="timelimit", mode="random")
    assert np.array_equal(new, g["trim_random_values"])  # pure gathers with the reference's random picks: exact
    new, fin, ttt = mod.trim_horizons(tv, tt, method="off", mode="average")
    assert np.array_equal(new, g["trim_off_values"]) and fin == 0 and ttt is None
    with pytest.raises(ValueError):
        mod.trim_horizons(tv, tt, method="nope")
    with pytest.raises(ValueError):
        mod.trim_horizons(tv, tt, method="timelimit", mode="nope")


def test_trim_horizons_production_shape_vs_oracle():
    hz = oracle.get_value_head_horizons(128, 30000)
    rng = np.random.RandomState(5)
    n_envs = 512
    tv = np.cumsum(np.abs(rng.normal(0, 1, [n_envs, len(hz), 1])), axis=1).astype(np.float32)
    tt = rng.randint(0, 30000, size=[n_envs])
    log = types.SimpleNamespace(watch_stats=lambda *a, **k: None, watch_mean=lambda *a, **k: None)
    runner = types.SimpleNamespace(ext_rewards=np.zeros([2, n_envs], np.float32), tvf_horizons=list(hz),
                                   value_heads=["ext"], N=2, K=len(hz), log=log, episode_length_buffer=[5000, 9000])
    args = types.SimpleNamespace(env=types.SimpleNamespace(timeout=27000),
                                 tvf=types.SimpleNamespace(eta_percentile=90, eta_buffer=32, eta_minh=128, trim_clip=-1))
    mod = tvf.TVFRunnerModule(runner, args)
    for mode in ("interpolate", "average", "substitute"):
        new, fin, ttt = mod.trim_horizons(tv, tt, method="timelimit", mode=mode)
        ref_new, ref_fin, ref_ttt = oracle.trim_horizons(hz, tv, tt, "timelimit", mode, timeout=27000)
        assert rel_err(new, ref_new) <= TOL and rel_err(fin, ref_fin) <= TOL and np.array_equal(ttt, ref_ttt)
