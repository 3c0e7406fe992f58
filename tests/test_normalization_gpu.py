"""GPU parity: ppo_b200.normalization (return trackers + running moments, CUDA through the C ABI) against the golden
vectors produced by the reference's own source and against the oracle on seeded rollouts."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import _device as dev, _lib, normalization as nz
from util import TOL, rel_err

pytestmark = pytest.mark.gpu

STATE_TOL = 1e-12  # float64 running moments: only the summation order inside a row differs


def close64(a, b, tol=STATE_TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300)


def test_running_mean_std_golden(golden_norm):
    g = golden_norm
    rms, at = nz.RunningMeanStd(), 0
    for i, n in enumerate(g["rms_in_sizes"]):
        rms.update(g["rms_in_batches"][at:at + n])
        at += n
        assert close64(rms.save_state(), g["rms_states"][i])
    other = nz.RunningMeanStd()
    other.restore_state(rms.save_state())
    other.update(3.5)  # rl/utils.py:427-430: a python float is a batch of one
    ref = oracle.RunningMeanStd()
    ref.mean, ref.var, ref.count = (np.float64(v) for v in g["rms_states"][-1])
    ref.update_from_moments(3.5, 0, 1)
    assert close64(other.save_state(), ref.state())


@pytest.mark.parametrize("tag,normalize,center,propagation", [("norm", True, False, False), ("center", True, True, False),
                                                              ("prop", True, False, True), ("raw", False, True, False)])
def test_intrinsic_returns_golden(golden_norm, tag, normalize, center, propagation):
    """Runner.calculate_intrinsic_returns (rl/rollout.py:1140-1180), two consecutive rollouts: the carry and the
    running moments persist on the device."""
    g = golden_norm
    ir = nz.IntrinsicReturns(g["ir_in_rewards_0"].shape[1])
    for roll in range(2):
        res = ir.calculate_intrinsic_returns(g[f"ir_in_rewards_{roll}"], g[f"ir_in_value_{roll}"],
                                             g[f"ir_in_terminals_{roll}"], 0.99, 0.95, normalize=normalize,
                                             center=center, propagation=propagation)
        assert isinstance(res["int_advantage"], np.ndarray) and res["int_advantage"].dtype == np.float32
        assert rel_err(res["int_advantage"], g[f"ir_{tag}_adv_{roll}"]) <= TOL
        assert rel_err(res["int_returns"], g[f"ir_{tag}_returns_{roll}"]) <= TOL
        assert rel_err(res["int_rewards"], g[f"ir_{tag}_rewards_{roll}"]) <= TOL
        assert close64(res["scale"].item(), g[f"ir_{tag}_scale_{roll}"])
        assert close64(ir.ems_norm.cpu().numpy(), g[f"ir_{tag}_ems_{roll}"])
        assert close64(ir.intrinsic_returns_rms.save_state(), g[f"ir_{tag}_rms_{roll}"])


@pytest.mark.parametrize("tag,kwargs", [("default", dict(gamma=0.99)), ("noclip", dict(gamma=0.999, clip=-1, scale=0.5)),
                                        ("tight", dict(gamma=0.9, clip=1.0)),
                                        ("ema", dict(gamma=0.99, mode="ema", ema_horizon=200.0)),
                                        ("custom", dict(gamma=0.99, mode="custom", clip=3.0))])
def test_vec_normalize_rewards_golden(golden_norm, tag, kwargs):
    g = golden_norm
    w = nz.VecNormalizeRewards(g["vn_in_rewards"].shape[1], **kwargs)
    if tag == "custom":
        w.ret_var = 6.25
    scaled, clips = w.normalize_rollout(g["vn_in_rewards"], g["vn_in_dones"])
    assert abs(w.ret_var - float(g[f"vn_{tag}_ret_var"])) <= 1e-9 * max(1.0, float(g[f"vn_{tag}_ret_var"]))
    assert rel_err(scaled, g[f"vn_{tag}_scaled"]) <= 1e-6
    assert int(clips.item()) == int(g[f"vn_{tag}_clips"])
    assert close64(w.current_returns.cpu().numpy(), g[f"vn_{tag}_returns"])
    assert close64(w.ret_rms.save_state(), g[f"vn_{tag}_rms"])
    buf = {}
    w.save_state(buf)
    w2 = nz.VecNormalizeRewards(g["vn_in_rewards"].shape[1], **kwargs)
    w2.restore_state(buf)
    assert close64(w2.ret_rms.save_state(), w.ret_rms.save_state()) and abs(w2.std - w.std) <= 1e-12 * w.std


@pytest.mark.parametrize("tag,kwargs", [
    ("edquad", dict(gamma=0.99, ed_type="quadratic")),
    ("edharm", dict(gamma=0.99, ed_type="harmonic", ed_bias=2.0, mode="ema", ema_horizon=300.0)),
    ("edpower", dict(gamma=0.995, ed_type="power", returns_transform=lambda x: x * 0.5))])
def test_vec_normalize_rewards_episodic_discounting_golden(golden_norm, tag, kwargs):
    """ed_type / ed_bias / returns_transform (rl/wrappers.py:870-876) against the reference's own wrapper."""
    g = golden_norm
    w = nz.VecNormalizeRewards(g["vn_in_rewards"].shape[1], **kwargs)
    scaled, clips = w.normalize_rollout(g["vn_in_rewards"], g["vn_in_dones"], times=g["vn_in_times"])
    assert rel_err(scaled, g[f"vn_{tag}_scaled"]) <= 1e-6
    assert int(clips.item()) == int(g[f"vn_{tag}_clips"])
    assert close64(w.current_returns.cpu().numpy(), g[f"vn_{tag}_returns"])
    assert close64(w.ret_rms.save_state(), g[f"vn_{tag}_rms"])
    assert abs(w.ret_var - float(g[f"vn_{tag}_ret_var"])) <= 1e-9 * max(1.0, float(g[f"vn_{tag}_ret_var"]))
    with pytest.raises(ValueError, match="times"):
        nz.VecNormalizeRewards(4, ed_type="quadratic").normalize_rollout(np.zeros((2, 4), np.float32), np.zeros((2, 4), bool))
    with pytest.raises(ZeroDivisionError):
        nz.episodic_normalization_constant(torch.arange(3, device="cuda"), "finite")
    assert bool(torch.isinf(nz.episodic_normalization_constant(torch.arange(3, device="cuda"), "geometric")).all())
    with pytest.raises(ValueError, match="Invalid discount_type"):
        nz.episodic_normalization_constant(torch.arange(3, device="cuda"), "linear")


def test_return_tracker_masked_golden(golden_norm):
    """ReturnTracker (rl/wrappers.py:737-755): the whole rollout in one call, then step by step like the reference's
    caller; float32 state bit for bit, running moments to the float32 rounding of NumPy's batch mean / variance."""
    g = golden_norm
    steps, envs = g["rt_in_rewards"].shape
    tr = nz.ReturnTracker(envs, 0.99)
    row_var = tr.update_rollout(g["rt_in_rewards"], g["rt_in_dones"], g["rt_in_mask"])
    assert np.array_equal(tr.current_returns.cpu().numpy(), g["rt_returns"][-1])
    assert close64(tr.ret_rms.save_state(), g["rt_states"][-1], tol=1e-6)
    assert close64(row_var.cpu().numpy(), g["rt_states"][:, 1], tol=1e-6)
    step = nz.ReturnTracker(envs, 0.99)
    for t in range(steps):
        step.update(g["rt_in_rewards"][t], g["rt_in_dones"][t], g["rt_in_mask"][t])
        assert np.array_equal(step.current_returns.cpu().numpy(), g["rt_returns"][t]), t
        assert close64(step.ret_rms.save_state(), g["rt_states"][t], tol=1e-6), t
    assert torch.equal(step.current_returns, tr.current_returns) and torch.equal(step.ret_rms.state, tr.ret_rms.state)
    tr.reset()
    assert float(tr.current_returns.abs().sum()) == 0.0


@pytest.mark.parametrize("steps,envs,kind", [(1, 1, "bool"), (64, 4096, "bool"), (300, 70, "f32"), (17, 1000, "none")])
def test_return_tracker_masked_vs_oracle(steps, envs, kind):
    rng = np.random.RandomState(steps + envs)
    r = (rng.randint(-1, 3, size=[steps, envs]) * rng.choice([1.0, 7.0], size=[steps, envs])).astype(np.float32)
    d = rng.random_sample([steps, envs]) < 0.05
    dones = d if kind == "bool" else d.astype(np.float32) if kind == "f32" else None
    mask = rng.random_sample([steps, envs]) < 0.5
    mask[steps // 2] = False
    cur, rms = np.zeros(envs, dtype=np.float32), oracle.RunningMeanStd()
    var = []
    for t in range(steps):
        oracle.return_tracker_update(cur, rms, r[t], dones[t] if dones is not None else np.zeros(envs, bool), mask[t], 0.97)
        var.append(rms.var)
    tr = nz.ReturnTracker(envs, 0.97)
    row_var = tr.update_rollout(r, dones, mask)
    got = tr.current_returns.cpu().numpy()
    assert np.max(np.abs(got - cur)) <= 1e-6 * max(np.max(np.abs(cur)), 1.0)  # (float32 dones: one float32 product more)
    if kind != "f32":
        assert np.array_equal(got, cur)
    assert close64(tr.ret_rms.save_state(), rms.state(), tol=1e-5)
    assert close64(row_var.cpu().numpy(), var, tol=1e-5)


@pytest.mark.parametrize("steps,envs", [(1, 1), (5, 33), (256, 4096), (300, 70), (700, 40)])
def test_trackers_vs_oracle(steps, envs):
    """Ragged shapes, more steps than one register pass (256), device tensors in -> device tensors out."""
    rng = np.random.RandomState(steps * 13 + envs)
    r = np.abs(rng.normal(0.2, 0.3, [steps, envs])).astype(np.float32)
    d = rng.random_sample([steps, envs]) < 0.03
    v = rng.normal(0.5, 0.4, [steps + 1, envs]).astype(np.float32)
    ems, rms = rng.normal(1.0, 0.5, envs), oracle.RunningMeanStd()
    ir = nz.IntrinsicReturns(envs)
    ir.ems_norm.copy_(torch.from_numpy(ems))
    adv, ret, rew, scale = oracle.intrinsic_returns(r, v, d, 0.99, 0.95, ems, rms, normalize=True, center=True)
    res = ir.calculate_intrinsic_returns(torch.from_numpy(r).cuda(), torch.from_numpy(v).cuda(),
                                         torch.from_numpy(d).cuda(), 0.99, 0.95, normalize=True, center=True)
    assert res["int_advantage"].is_cuda
    assert rel_err(res["int_advantage"].cpu().numpy(), adv) <= TOL and rel_err(res["int_returns"].cpu().numpy(), ret) <= TOL
    assert rel_err(res["int_rewards"].cpu().numpy(), rew) <= TOL
    assert close64(ir.ems_norm.cpu().numpy(), ems) and close64(ir.intrinsic_returns_rms.save_state(), rms.state(), 1e-10)
    # reward wrapper on the same rollout
    cur, rms2 = np.zeros(envs), oracle.RunningMeanStd()
    scaled, clips = oracle.vec_normalize_rewards(r * 40, d, 0.99, cur, rms2, clip=2.0, scale=0.7)
    w = nz.VecNormalizeRewards(envs, gamma=0.99, clip=2.0, scale=0.7)
    got, nclip = w.normalize_rollout(r * 40, d)
    assert rel_err(got, scaled) <= 1e-6 and abs(int(nclip.item()) - clips) <= max(2, clips // 100000)
    assert close64(w.current_returns.cpu().numpy(), cur) and close64(w.ret_rms.save_state(), rms2.state(), 1e-10)


def test_env_shards_add_up():
    """Env-sharded runs exchange only the two [N] row vectors: summing them over two shards by hand gives the moments
    of the unsharded rollout (what the all-reduce in normalization._fold_rows does over NCCL)."""
    lib = _lib.load()
    steps, envs = 64, 96
    rng = np.random.RandomState(2)
    x = torch.from_numpy(rng.normal(2.0, 1.5, [steps, envs])).cuda()
    s = dev.stream_ptr()

    def rows(t, center, total):
        out = torch.empty(steps, dtype=torch.float64, device="cuda")
        _lib.check(lib.ppo_row_sums(t.data_ptr(), steps, t.shape[1], dev.ptr(center), float(total), out.data_ptr(), s),
                   "ppo_row_sums")
        return out

    halves = [x[:, :40].contiguous(), x[:, 40:].contiguous()]
    sums = rows(halves[0], None, envs) + rows(halves[1], None, envs)
    sq = rows(halves[0], sums, envs) + rows(halves[1], sums, envs)
    state = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device="cuda")
    _lib.check(lib.ppo_rms_fold(sums.data_ptr(), sq.data_ptr(), steps, float(envs), state.data_ptr(), 0, s), "ppo_rms_fold")
    ref = oracle.RunningMeanStd()
    for t in range(steps):
        ref.update(x[t].cpu().numpy())
    assert close64(state.cpu().numpy(), ref.state(), 1e-10)


def test_error_behaviour():
    with pytest.raises(ValueError):
        nz.RunningMeanStd(shape=(3,))
    with pytest.raises(ValueError):
        nz.VecNormalizeRewards(4, mode="median")
    ir = nz.IntrinsicReturns(4)
    with pytest.raises(ValueError):
        ir.calculate_intrinsic_returns(np.zeros([3, 4], np.float32), np.zeros([3, 4], np.float32),
                                       np.zeros([3, 4], bool), 0.99)
    with pytest.raises(ValueError):
        nz.track_returns(np.zeros([3, 5], np.float32), None, 0.9, ir.ems_norm, ir.intrinsic_returns_rms)
