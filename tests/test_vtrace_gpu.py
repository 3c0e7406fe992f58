"""GPU parity: ppo_b200.vtrace (CUDA, through the C ABI) against the oracle, the golden vectors and the
reference's on-policy identities (rl/unit_tests.py:61-87)."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import returns, synthetic, vtrace
from util import TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("steps,envs,n_act,lamb", [(256, 512, 18, 1.0), (128, 128, 18, 0.9), (7, 33, 4, 1.0),
                                                   (1, 1, 2, 1.0), (300, 40, 6, 0.95)])
def test_vtrace_vs_oracle(steps, envs, n_act, lamb):
    r, d, v = synthetic.rollout(steps, envs, seed=steps + 3 * envs)
    beh, tgt, act = synthetic.policies(steps, envs, n_act, seed=steps)
    ref = oracle.importance_sampling_v_trace(beh, tgt, act, r, d, v[:-1], v[-1], 0.999, lamb)
    got = vtrace.importance_sampling_v_trace(beh, tgt, act, r, d, v[:-1], v[-1], 0.999, lamb)
    for g, f, name in zip(got, ref, ("vs", "adv", "cs")):
        assert isinstance(g, np.ndarray) and g.dtype == np.float32 and g.shape == f.shape
        assert rel_err(g, f) <= TOL, name
    # cs = lamb * min(1, exp(dlogp)) is elementwise: bit-exact gather indexing shows up as tiny ulp distance
    assert np.abs(got[2] - ref[2]).max() <= 2e-7


@pytest.mark.parametrize("ew", ["8", "16", "32"])
@pytest.mark.parametrize("steps,envs", [(256, 512), (37, 45), (600, 24), (3, 9)])
def test_vtrace_env_tile_widths(monkeypatch, ew, steps, envs):
    """Every env-tile width of the kernel (the launcher picks one by occupancy) gives the oracle's answer, for chunk
    lengths that are register-cached (<= 16 steps per chunk) and that are not (600 steps over 16 chunks)."""
    monkeypatch.setenv("PPO_VTRACE_EW", ew)
    r, d, v = synthetic.rollout(steps, envs, seed=steps + envs)
    beh, tgt, act = synthetic.policies(steps, envs, 18, seed=envs)
    ref = oracle.importance_sampling_v_trace(beh, tgt, act, r, d, v[:-1], v[-1], 0.999, 0.95)
    got = vtrace.importance_sampling_v_trace(beh, tgt, act, r, d, v[:-1], v[-1], 0.999, 0.95)
    for g, f, name in zip(got, ref, ("vs", "adv", "cs")):
        assert rel_err(g, f) <= TOL, name


def test_vtrace_device_actions_out_of_range_raise():
    """rl/vtrace.py:90 indexes with the actions, so an out-of-range action is an IndexError; for device tensors the
    kernel itself checks (clamped read + flag word)."""
    steps, envs, n_act = 32, 40, 6
    r, d, v = synthetic.rollout(steps, envs, seed=1)
    beh, tgt, act = synthetic.policies(steps, envs, n_act, seed=1)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    args = [cu(beh), cu(tgt), None, cu(r), cu(d), cu(v[:-1].copy()), cu(v[-1].copy()), 0.99, 0.9]
    good = vtrace.importance_sampling_v_trace(*(args[:2] + [cu(act)] + args[3:]))
    for where, value in (((0, 0), n_act), ((steps - 1, envs - 1), -n_act - 1), ((5, 7), 2 ** 40)):
        bad = act.astype(np.int64).copy()
        bad[where] = value
        with pytest.raises(IndexError):
            vtrace.importance_sampling_v_trace(*(args[:2] + [cu(bad)] + args[3:]))
        # check=False: no read-back, the index is clamped (no fault, the other envs are untouched)
        out = vtrace.importance_sampling_v_trace(*(args[:2] + [cu(bad)] + args[3:]), check=False)
        torch.cuda.synchronize()
        keep = np.ones(envs, dtype=bool)
        keep[where[1]] = False
        assert np.array_equal(out[0].cpu().numpy()[:, keep], good[0].cpu().numpy()[:, keep])
    # the most negative valid index still wraps like NumPy
    ok = act.astype(np.int64).copy()
    ok[0, 0] = -n_act
    vtrace.importance_sampling_v_trace(*(args[:2] + [cu(ok)] + args[3:]))


def test_vtrace_on_policy_identities():
    # rl/unit_tests.py:61-87: with pi == mu and lambda = 1, vs == bootstrapped returns and adv == GAE(lambda=1)
    steps, envs = 64, 48
    r, d, v = synthetic.rollout(steps, envs, seed=9)
    beh, _, act = synthetic.policies(steps, envs, 18, seed=9)
    vs, adv, cs = vtrace.importance_sampling_v_trace(beh, beh, act, r, d, v[:-1], v[-1], 0.99, 1.0)
    boot = returns.calculate_bootstrapped_returns(r, d, v[-1], 0.99)
    g = returns.gae(r, v[:-1], v[-1], d, 0.99, 1.0)
    assert rel_err(vs, boot) <= 1e-5 and rel_err(adv, g) <= 1e-5
    assert np.array_equal(cs, np.ones_like(cs))


def test_vtrace_golden_and_dtypes(golden):
    g = golden
    beh, tgt, act = g["vt_in_beh"], g["vt_in_tgt"], g["vt_in_act"]
    r, d, v = g["vt_in_rewards"], g["vt_in_dones"], g["vt_in_values"]
    got = vtrace.importance_sampling_v_trace(beh, tgt, act, r, d, v[:-1], v[-1], 0.99, 0.9)
    for x, name in zip(got, ("vt_vs", "vt_adv", "vt_cs")):
        assert rel_err(x, g[name]) <= TOL, name
    vs, adv, _ = vtrace.importance_sampling_v_trace(beh, beh, act, r, d, v[:-1], v[-1], 0.99, 1.0)
    assert rel_err(vs, g["vt_onpolicy_vs"]) <= TOL and rel_err(adv, g["vt_onpolicy_adv"]) <= TOL
    assert rel_err(vtrace.v_trace_trust_region(beh, tgt), g["vt_trust"]) <= TOL
    # int32 and negative (NumPy-style) action indices, device tensors in -> device tensors out
    n_act = beh.shape[2]
    neg = (act.astype(np.int64) - n_act).astype(np.int32)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    out = vtrace.importance_sampling_v_trace(cu(beh), cu(tgt), cu(neg), cu(r), cu(d), cu(v[:-1].copy()), cu(v[-1].copy()),
                                             0.99, 0.9)
    assert all(isinstance(o, torch.Tensor) and o.is_cuda for o in out)
    assert rel_err(out[0].cpu().numpy(), g["vt_vs"]) <= TOL
    with pytest.raises(IndexError):
        bad = act.copy()
        bad[0, 0] = n_act
        vtrace.importance_sampling_v_trace(beh, tgt, bad, r, d, v[:-1], v[-1], 0.99, 1.0)


@pytest.mark.parametrize("steps,envs,n_act", [(64, 96, 18), (5, 7, 3), (256, 512, 18)])
def test_trust_region_vs_oracle(steps, envs, n_act):
    beh, tgt, _ = synthetic.policies(steps, envs, n_act, seed=envs)
    ref = oracle.v_trace_trust_region(beh, tgt)
    got = vtrace.v_trace_trust_region(beh, tgt)
    assert got.shape == ref.shape and rel_err(got, ref) <= TOL
