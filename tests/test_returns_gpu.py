"""GPU parity: ppo_b200.returns (CUDA, through the C ABI) against the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import returns, synthetic
from util import TOL, rel_err, ulp_diff

pytestmark = pytest.mark.gpu


def test_gae_golden(golden):
    g = golden
    r, d, v = g["gae_in_rewards"], g["gae_in_dones"], g["gae_in_values"]
    assert rel_err(returns.gae(r, v[:-1], v[-1], d, 0.999, 0.95), g["gae_bool"]) <= TOL
    assert rel_err(returns.gae(r, v[:-1], v[-1], d.astype(np.float32), 0.999, 0.95), g["gae_f32done"]) <= TOL
    assert rel_err(returns.gae(r, v[:-1], v[-1], None, 0.99, 0.9), g["gae_noterm"]) <= TOL
    assert rel_err(returns.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.8), g["tdl_bool"]) <= TOL
    assert rel_err(returns.calculate_bootstrapped_returns(r, d, v[-1], 0.999), g["boot_scalar"]) <= TOL
    assert rel_err(returns.calculate_bootstrapped_returns(r, d, v[-1], g["boot_in_gamma_arr"]), g["boot_array"]) <= TOL
    # bool terminals: float64 carry like the reference -> at most 1 float32 ulp from re-association
    assert ulp_diff(returns.gae(r, v[:-1], v[-1], d, 0.999, 0.95), g["gae_bool"]) <= 1


def test_gae_known_answer():
    # rl/unit_tests.py:203-210
    r = np.asarray([1, 0, 2, 4, 6], dtype=np.float32)[:, None]
    d = np.asarray([0, 0, 1, 0, 0], dtype=np.float32)[:, None]
    v = np.asarray([0, 0.5, 0.5, 3, 4], dtype=np.float32)[:, None]
    res = returns.gae(r, v, np.asarray(5, dtype=np.float32), d, gamma=0.5, lamb=1.0)
    assert isinstance(res, np.ndarray) and res.dtype == np.float32
    assert np.max(np.abs(res - np.asarray([1.5, 0.5, 1.5, 5.25, 4.5])[:, None])) < 1e-6


@pytest.mark.parametrize("steps,envs", [(128, 128), (1, 1), (5, 33), (17, 4), (256, 4096), (300, 70), (1000, 64)])
def test_gae_vs_oracle(steps, envs):
    r, d, v = synthetic.rollout(steps, envs, seed=steps * 7 + envs)
    ref_adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    ref_ret = oracle.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.9)
    adv, ret = returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.9)
    assert rel_err(adv, ref_adv) <= TOL and rel_err(ret, ref_ret) <= TOL
    assert ulp_diff(adv, ref_adv) <= 1
    assert rel_err(returns.gae(r, v[:-1], v[-1], d, 0.999, 0.95), ref_adv) <= TOL
    assert rel_err(returns.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.9), ref_ret) <= TOL
    boot = returns.calculate_bootstrapped_returns(r, d, v[-1], 0.999)
    assert rel_err(boot, oracle.calculate_bootstrapped_returns(r, d, v[-1], 0.999)) <= TOL


def test_gae_strided_and_device_inputs():
    # value estimates arrive as a strided slice of the [N+1, A, K] TVF buffer (rl/tvf.py:325-340)
    steps, envs, k = 64, 96, 5
    r, d, _ = synthetic.rollout(steps, envs, seed=3)
    vs = synthetic.value_samples(steps, envs, k, seed=3)
    v = vs[:, :, -1]
    ref = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    assert rel_err(returns.gae(r, v[:-1], v[-1], d, 0.999, 0.95), ref) <= TOL
    dvs = torch.from_numpy(vs).cuda()
    dv = dvs[:, :, -1]
    out = returns.gae(torch.from_numpy(r).cuda(), dv[:-1], dv[-1], torch.from_numpy(d).cuda(), 0.999, 0.95)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert rel_err(out.cpu().numpy(), ref) <= TOL


def test_gae_stats_and_empty():
    r, d, v = synthetic.rollout(96, 200, seed=5)
    adv, ret, stats = returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.9, with_stats=True)
    s = stats.cpu().numpy()
    assert s[0] == adv.size
    np.testing.assert_allclose(s[1], adv.astype(np.float64).sum(), rtol=1e-12)
    np.testing.assert_allclose(s[2], (adv.astype(np.float64) ** 2).sum(), rtol=1e-12)
    assert s[3] == adv.min() and s[4] == adv.max()
    np.testing.assert_allclose(s[5], ret.astype(np.float64).sum(), rtol=1e-12)
    assert s[7] == ret.min() and s[8] == ret.max()
    # run twice: the ticket must have been reset
    _, _, stats2 = returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.9, with_stats=True)
    assert np.array_equal(stats2.cpu().numpy(), s)
    e = returns.gae(np.zeros([0, 4], np.float32), np.zeros([0, 4], np.float32), np.zeros([4], np.float32),
                    np.zeros([0, 4], bool), 0.99)
    assert e.shape == (0, 4)
