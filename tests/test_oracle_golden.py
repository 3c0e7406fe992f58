"""CPU: the oracle restatement must reproduce the real reference bit-for-bit on the golden vectors.

The vectors were produced by tests/golden/make_golden.py, which imports the unmodified reference.
Unless a tolerance is written in a test, comparisons are exact (np.array_equal).
"""
import numpy as np
import pytest

import oracle


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


def test_gae_family(golden):
    g = golden
    r, d, v = g["gae_in_rewards"], g["gae_in_dones"], g["gae_in_values"]
    assert same(oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95), g["gae_bool"])
    assert same(oracle.gae(r, v[:-1], v[-1], d.astype(np.float32), 0.999, 0.95), g["gae_f32done"])
    assert same(oracle.gae(r, v[:-1], v[-1], None, 0.99, 0.9), g["gae_noterm"])
    assert same(oracle.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.8), g["tdl_bool"])
    assert same(oracle.calculate_bootstrapped_returns(r, d, v[-1], 0.999), g["boot_scalar"])
    assert same(oracle.calculate_bootstrapped_returns(r, d, v[-1], g["boot_in_gamma_arr"]), g["boot_array"])


def test_gae_known_answer():
    # rl/unit_tests.py:203-210
    r = np.asarray([1, 0, 2, 4, 6], dtype=np.float32)[:, None]
    d = np.asarray([0, 0, 1, 0, 0], dtype=np.float32)[:, None]
    v = np.asarray([0, 0.5, 0.5, 3, 4], dtype=np.float32)[:, None]
    res = oracle.gae(r, v, np.asarray(5, dtype=np.float32), d, gamma=0.5, lamb=1.0)
    assert np.max(np.abs(res - np.asarray([1.5, 0.5, 1.5, 5.25, 4.5])[:, None])) < 1e-6


def test_advantage_normalisation(golden):
    assert same(oracle.normalize_advantages(golden["gae_bool"], 1e-8), golden["advnorm"])
    assert same(oracle.normalize_advantages(golden["gae_bool"], 1e-8, 1.5), golden["advnorm_clip"])


def _tvf_args(g):
    keys = ['gamma', 'rewards', 'dones', 'required_horizons', 'value_sample_horizons', 'value_samples']
    a = {k: g["tvf_in_" + k] for k in keys}
    a['gamma'] = float(a['gamma'])
    return a


@pytest.mark.parametrize("n", [1, 8, 48])
def test_tvf_fixed_lists(golden, n):
    a = _tvf_args(golden)
    assert same(oracle.sampled_return_multi(n_step_list=[n], **a), golden[f"tvf_fast_n{n}"])
    assert same(oracle.sampled_return_slow(n_step_list=[n], **a), golden[f"tvf_slow_n{n}"])


def test_tvf_sample_tables(golden):
    a = _tvf_args(golden)
    s = golden["tvf_in_samples"]
    assert same(oracle.sampled_return_multi(n_step_samples=s, **a), golden["tvf_fast_samples"])
    assert same(oracle.sampled_return_slow(n_step_samples=s, **a), golden["tvf_slow_samples"])
    assert same(oracle.sampled_return_multi(n_step_samples=s, use_log_interpolation=True, **a),
                golden["tvf_fast_samples_log"])
    # the reference's own acceptance criterion, tests/test_tvf.py:46
    ref = golden["tvf_slow_samples"]
    assert np.abs(ref - golden["tvf_fast_samples"]).max() <= 1e-5 * ref.max()


@pytest.mark.parametrize("dist,mode", [("fixed", "standard"), ("exponential", "standard"),
                                       ("exponential", "advanced"), ("uniform", "advanced"),
                                       ("hyperbolic", "clipped"), ("quadratic", "adaptive"),
                                       ("exponential", "mcx"), ("exponential", "full")])
def test_tvf_dispatcher(golden, dist, mode):
    a = _tvf_args(golden)
    got = oracle.get_return_estimate(dist, mode, n_step=6, max_samples=5, seed=17, **a)
    assert same(got, golden[f"tvf_gre_{dist}_{mode}"])


def test_tvf_dispatcher_rejects_unknown(golden):
    a = _tvf_args(golden)
    with pytest.raises(ValueError):
        oracle.get_return_estimate("nope", "standard", **a)
    with pytest.raises(ValueError):
        oracle.get_return_estimate("uniform", "nope", **a)


def test_tvf_production_horizons(golden):
    g = golden
    hz, r, d, vs = g["prod_in_horizons"], g["prod_in_rewards"], g["prod_in_dones"], g["prod_in_vs"]
    assert same(oracle.get_value_head_horizons(32, 1000), hz)
    got = oracle.get_return_estimate("exponential", "advanced", 0.99997, r, d, hz, hz, vs,
                                     n_step=20, max_samples=8, seed=1)
    assert same(got, g["prod_adv"])
    got = oracle.get_return_estimate("fixed", "advanced", 0.99997, r, d, hz, hz, vs,
                                     n_step=40, max_samples=8, seed=1)
    assert same(got, g["prod_fixed"])
    got = oracle.get_return_estimate("exponential", "advanced", 0.99997, r, d, hz, hz, vs,
                                     n_step=20, max_samples=8, use_log_interpolation=True, seed=1)
    assert same(got, g["prod_adv_log"])


def test_head_horizons(golden):
    assert same(oracle.get_value_head_horizons(128, 30000), golden["heads_128_30000"])
    h, w = oracle.get_value_head_horizons(16, 100, include_weight=True)
    assert same(h, golden["heads_16_100"]) and same(w, golden["heads_16_100_w"])
    assert same(oracle.get_value_head_horizons(9, 64, spacing="linear"), golden["heads_lin_9_64"])


def test_horizon_interpolate(golden):
    # tests/test_tvf.py:121-129 known answer
    kh = np.asarray([0, 1, 2, 10, 100])
    kv = np.asarray([0, 5, 10, -1, 2])[None, :].repeat(11, axis=0)
    kt = np.asarray([-100, -1, 0, 1, 2, 3, 4, 99, 100, 101, 200])
    res = oracle.horizon_interpolate(kh, kv, kt)
    expected = [0, 0, 0, 5, 10, (7 / 8) * 10 + (1 / 8) * -1, (6 / 8) * 10 + (2 / 8) * -1, 1.96666667, 2, 2, 2]
    assert np.max(np.abs(np.asarray(expected) - res)) < 1e-6
    assert same(res, golden["hinterp_kat"])
    hz = golden["prod_in_horizons"]
    assert same(oracle.horizon_interpolate(hz, golden["hinterp_in_values"], golden["hinterp_in_targets"]),
                golden["hinterp"])


def test_rediscount(golden):
    hz, rv = golden["prod_in_horizons"], golden["redis_in_values"]
    assert same(oracle.get_rediscounted_value_estimate(rv, 0.9999, 0.999, hz), golden["redis_down"])
    assert same(oracle.get_rediscounted_value_estimate(rv, 0.99, 0.9999, hz), golden["redis_up"])
    assert same(oracle.get_rediscounted_value_estimate(rv, 0.99, 0.99, hz), golden["redis_same"])


def test_trim_horizons(golden):
    """rl/tvf.py:91-208 through a stand-in runner (tests/golden/make_golden.py), every method / mode, bit for bit."""
    g = golden
    hz, tv, tt, ep = g["trim_in_horizons"], g["trim_in_values"], g["trim_in_time"], list(g["trim_in_ep_lengths"])
    for method in ("timelimit", "est_term"):
        for mode in ("interpolate", "average", "substitute"):
            new, fin, ttt = oracle.trim_horizons(hz, tv, tt, method, mode, timeout=900, episode_length_buffer=ep)
            assert same(new, g[f"trim_{method}_{mode}_values"]) and same(fin, g[f"trim_{method}_{mode}_final"])
            assert same(np.asarray(ttt), g[f"trim_{method}_{mode}_ttt"])
    new, fin, _ = oracle.trim_horizons(hz, tv, tt, "timelimit", "average", timeout=900, trim_clip=0.25)
    assert same(new, g["trim_clip_values"]) and same(fin, g["trim_clip_final"])
    np.random.seed(123)
    new, fin, _ = oracle.trim_horizons(hz, tv, tt, "timelimit", "random", timeout=900)
    assert same(new, g["trim_random_values"]) and same(fin, g["trim_random_final"])
    new, fin, ttt = oracle.trim_horizons(hz, tv, tt, "off", "average")
    assert same(new, g["trim_off_values"]) and fin == 0 and ttt is None


def test_vtrace(golden):
    g = golden
    beh, tgt, act = g["vt_in_beh"], g["vt_in_tgt"], g["vt_in_act"]
    r, d, v = g["vt_in_rewards"], g["vt_in_dones"], g["vt_in_values"]
    vs, adv, cs = oracle.importance_sampling_v_trace(beh, tgt, act, r, d, v[:-1], v[-1], 0.99, 0.9)
    assert same(vs, g["vt_vs"]) and same(adv, g["vt_adv"]) and same(cs, g["vt_cs"])
    vs, adv, _ = oracle.importance_sampling_v_trace(beh, beh, act, r, d, v[:-1], v[-1], 0.99, 1.0)
    assert same(vs, g["vt_onpolicy_vs"]) and same(adv, g["vt_onpolicy_adv"])
    assert same(oracle.v_trace_trust_region(beh, tgt), g["vt_trust"])
    # on-policy identities of rl/unit_tests.py:61-87 (float32 rounding tolerance 1e-5 relative)
    ret = oracle.calculate_bootstrapped_returns(r, d, v[-1], 0.99)
    gae = oracle.gae(r, v[:-1], v[-1], d, 0.99, 1.0)
    assert np.abs(vs - ret).max() <= 1e-5 * np.abs(ret).max()
    assert np.abs(adv - gae).max() <= 1e-5 * np.abs(gae).max()


def test_running_mean_std(golden_norm):
    """rl/utils.py:379-443, bit for bit."""
    g = golden_norm
    rms, at, states = oracle.RunningMeanStd(), 0, []
    for n in g["rms_in_sizes"]:
        rms.update(g["rms_in_batches"][at:at + n])
        at += n
        states.append(rms.state())
    assert same(np.asarray(states), g["rms_states"])


def test_intrinsic_returns(golden_norm):
    """Runner.calculate_intrinsic_returns (rl/rollout.py:1140-1180) over two consecutive rollouts, bit for bit."""
    g = golden_norm
    for tag, normalize, center, propagation in [("norm", True, False, False), ("center", True, True, False),
                                                ("prop", True, False, True), ("raw", False, True, False)]:
        ems, rms = np.zeros(g["ir_in_rewards_0"].shape[1]), oracle.RunningMeanStd()
        for roll in range(2):
            adv, ret, rew, scale = oracle.intrinsic_returns(
                g[f"ir_in_rewards_{roll}"], g[f"ir_in_value_{roll}"], g[f"ir_in_terminals_{roll}"], 0.99, 0.95, ems,
                rms, normalize=normalize, center=center, propagation=propagation)
            assert same(adv, g[f"ir_{tag}_adv_{roll}"]) and same(rew, g[f"ir_{tag}_rewards_{roll}"]), tag
            assert same(ret.astype(np.float32), g[f"ir_{tag}_returns_{roll}"]), tag
            assert same(np.asarray(scale, dtype=np.float64), g[f"ir_{tag}_scale_{roll}"])
            assert same(ems, g[f"ir_{tag}_ems_{roll}"]) and same(rms.state(), g[f"ir_{tag}_rms_{roll}"])


def test_vec_normalize_rewards(golden_norm):
    """VecNormalizeRewardWrapper.step (rl/wrappers.py:795-895) over a rollout, bit for bit."""
    g = golden_norm
    for tag, kwargs in [("default", dict(gamma=0.99)), ("noclip", dict(gamma=0.999, clip=-1, scale=0.5)),
                        ("tight", dict(gamma=0.9, clip=1.0))]:
        cur, rms = np.zeros(g["vn_in_rewards"].shape[1], dtype=np.float32), oracle.RunningMeanStd()
        cur = cur.astype(np.float64)  # the reference's float32 carry becomes float64 after the first step (:847)
        scaled, clips = oracle.vec_normalize_rewards(g["vn_in_rewards"], g["vn_in_dones"], kwargs["gamma"], cur, rms,
                                                     clip=kwargs.get("clip", 10.0), scale=kwargs.get("scale", 1.0))
        assert same(scaled, g[f"vn_{tag}_scaled"]) and clips == int(g[f"vn_{tag}_clips"]), tag
        assert same(cur, g[f"vn_{tag}_returns"]) and same(rms.state(), g[f"vn_{tag}_rms"])
    # modes "ema" and "custom" (rl/wrappers.py:878-881, :896-897)
    for tag, kwargs in [("ema", dict(mode="ema", ema_horizon=200.0)), ("custom", dict(mode="custom", ret_var=6.25, clip=3.0))]:
        cur, rms = np.zeros(g["vn_in_rewards"].shape[1], dtype=np.float64), oracle.RunningMeanStd()
        scaled, clips, ret_var = oracle.vec_normalize_rewards(g["vn_in_rewards"], g["vn_in_dones"], 0.99, cur, rms, **kwargs)
        assert same(scaled, g[f"vn_{tag}_scaled"]) and clips == int(g[f"vn_{tag}_clips"]), tag
        assert same(cur, g[f"vn_{tag}_returns"]) and same(rms.state(), g[f"vn_{tag}_rms"])
        # (the reference takes np.var of its float32-turned-float64 tracker: same bits here)
        assert same(np.asarray(ret_var, dtype=np.float64), g[f"vn_{tag}_ret_var"]), tag


def test_vec_normalize_rewards_episodic_discounting(golden_norm):
    """The wrapper's ed_type / returns_transform (rl/wrappers.py:870-876, EpisodicDiscounting.get_normalization_constant
    :52-72), bit for bit."""
    g = golden_norm
    for tag, gamma, kwargs in [("edquad", 0.99, dict(ed_type="quadratic")),
                               ("edharm", 0.99, dict(ed_type="harmonic", ed_bias=2.0, mode="ema", ema_horizon=300.0)),
                               ("edpower", 0.995, dict(ed_type="power", returns_transform=lambda x: x * 0.5))]:
        cur, rms = np.zeros(g["vn_in_rewards"].shape[1], dtype=np.float64), oracle.RunningMeanStd()
        res = oracle.vec_normalize_rewards(g["vn_in_rewards"], g["vn_in_dones"], gamma, cur, rms, times=g["vn_in_times"],
                                           **kwargs)
        assert same(res[0], g[f"vn_{tag}_scaled"]) and res[1] == int(g[f"vn_{tag}_clips"]), tag
        assert same(cur, g[f"vn_{tag}_returns"]) and same(rms.state(), g[f"vn_{tag}_rms"]), tag
        if len(res) == 3:
            assert same(np.asarray(res[2], dtype=np.float64), g[f"vn_{tag}_ret_var"]), tag
    with pytest.raises(ZeroDivisionError):  # the wrapper's gamma = 1.0 (:872): 1 / (1 - gamma) in plain Python (:55)
        oracle.episodic_normalization_constant(np.arange(3), "finite")
    with np.errstate(divide="ignore"):      # ... and an array divided by zero (:60): infinite normalisers, no exception
        assert np.all(np.isinf(oracle.episodic_normalization_constant(np.arange(3), "geometric")))
    with pytest.raises(ValueError):
        oracle.episodic_normalization_constant(np.arange(3), "linear")


def test_return_tracker_masked(golden_norm):
    """ReturnTracker.update (rl/wrappers.py:737-755) with per-step env masks, bit for bit (float32 state)."""
    g = golden_norm
    cur, rms = np.zeros(g["rt_in_rewards"].shape[1], dtype=np.float32), oracle.RunningMeanStd()
    for t in range(g["rt_in_rewards"].shape[0]):
        oracle.return_tracker_update(cur, rms, g["rt_in_rewards"][t], g["rt_in_dones"][t], g["rt_in_mask"][t], 0.99)
        assert cur.dtype == np.float32 and np.array_equal(cur, g["rt_returns"][t]), t
        assert same(rms.state(), g["rt_states"][t]), t
    assert np.array_equal(g["rt_states"][7], g["rt_states"][6])  # the step without a masked env changed nothing


LOSS_CASES = [("all", "off", None, None), ("hweighted", "h_weighted", None, None), ("single", "h_weighted", [5], 5),
              ("prefix", "off", list(range(8)), -7)]


def test_tvf_value_loss(golden_norm):
    """TVFRunnerModule.on_train_value_minibatch (rl/tvf.py:32-77) and its autograd gradient.  Tolerance: 2 float32
    ulp on the loss (torch and NumPy sum the horizon axis in different orders) and on the gradient (autograd
    chains float32 products, the oracle rounds once)."""
    from util import ulp_diff
    g = golden_norm
    for tag, weighting, required, single in LOSS_CASES:
        sel = slice(None) if required is None else required
        loss, grad = oracle.tvf_value_loss(g["loss_in_targets"], g["loss_in_predictions"][:, sel, 0], g["loss_in_weights"],
                                           g["loss_in_horizons"], 0.7, required, single, weighting, 1000, 20,
                                           loss_scale=0.25)
        assert ulp_diff(loss, g[f"loss_{tag}"]) <= 2, tag
        assert ulp_diff(grad, g[f"loss_{tag}_grad"][:, :, 0]) <= 2, tag
        assert not g[f"loss_{tag}_grad"][:, :, 1].any()  # only the "ext" value head is trained (:43)
