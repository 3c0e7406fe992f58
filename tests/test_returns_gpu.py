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
    r[:, 7:40:3], v[:, 7:40:3] = 0.0, 0.0  # silent envs: exact zeros for the sparsity the logger reports (rl/logger.py:201-229)
    r[50:, 100] = 0.0
    v[50:, 100] = 0.0
    adv, ret, stats = returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.9, with_stats=True)
    s = stats.cpu().numpy()
    assert s[0] == adv.size
    assert s[12] == np.count_nonzero(adv == 0) > 0 and s[13] == np.count_nonzero(ret == 0) > 0
    np.testing.assert_allclose(s[1], adv.astype(np.float64).sum(), rtol=1e-12)
    np.testing.assert_allclose(s[2], (adv.astype(np.float64) ** 2).sum(), rtol=1e-12)
    assert s[3] == adv.min() and s[4] == adv.max()
    np.testing.assert_allclose(s[5], ret.astype(np.float64).sum(), rtol=1e-12)
    assert s[7] == ret.min() and s[8] == ret.max()
    # moments of the value estimates the scan read, all N + 1 rows (rl/rollout.py:1205)
    assert s[11] == v.size
    np.testing.assert_allclose(s[9], v.astype(np.float64).sum(), rtol=1e-12)
    np.testing.assert_allclose(s[10], (v.astype(np.float64) ** 2).sum(), rtol=1e-12)
    # run twice: the ticket must have been reset
    _, _, stats2 = returns.gae_td_lambda(r, v[:-1], v[-1], d, 0.999, 0.95, 0.9, with_stats=True)
    assert np.array_equal(stats2.cpu().numpy(), s)
    e = returns.gae(np.zeros([0, 4], np.float32), np.zeros([0, 4], np.float32), np.zeros([4], np.float32),
                    np.zeros([0, 4], bool), 0.99)
    assert e.shape == (0, 4)


@pytest.mark.parametrize("steps,envs,clip,fused", [(256, 4096, None, True), (96, 200, 1.5, True), (33, 5, None, True),
                                                    (256, 32768, None, False)])
def test_gae_with_fused_normalisation(steps, envs, clip, fused):
    """ppo_gae_normalized: one cooperative launch when the grid is resident (flagged by the return value), else the two
    launches; either way the bits of ppo_gae followed by ppo_adv_normalize."""
    from ppo_b200 import _lib, _device as dev
    lib = _lib.load()
    r, d, v = synthetic.rollout(steps, envs, seed=3)
    r, d, v = (torch.from_numpy(x).cuda() for x in (r, d, v))
    d = d.view(torch.uint8)
    outs = [torch.empty((steps, envs), device="cuda") for _ in range(6)]
    st = [torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device="cuda") for _ in range(2)]
    ws = torch.zeros(int(lib.ppo_gae_workspace_bytes(envs)), dtype=torch.uint8, device="cuda")
    s = dev.stream_ptr()
    args = (r.data_ptr(), envs, v.data_ptr(), envs, 1, v[steps].data_ptr(), 1, d.data_ptr(), _lib.DONE_U8, envs,
            steps, envs, 0.999, 0.95, 0.9)
    # the shape query that env-sharded ranks agree on before relying on the in-launch exchange says the same
    assert lib.ppo_gae_normalized_fits(steps, envs, _lib.DONE_U8, 1) == (1 if fused else 0)
    for _ in range(2):  # twice: the flags of the cooperative launch must be left clear
        rc = lib.ppo_gae_normalized(*args, 1e-8, clip or 0.0, outs[0].data_ptr(), outs[1].data_ptr(),
                                    outs[2].data_ptr(), st[0].data_ptr(), ws.data_ptr(), ws.numel(), s)
        assert rc == (1 if fused else 0), lib.ppo_last_error()
    _lib.check(lib.ppo_gae(*args, outs[3].data_ptr(), outs[4].data_ptr(), st[1].data_ptr(), ws.data_ptr(), ws.numel(), s),
               "ppo_gae")
    _lib.check(lib.ppo_adv_normalize(outs[3].data_ptr(), steps * envs, st[1].data_ptr(), 1e-8, clip or 0.0,
                                     outs[5].data_ptr(), s), "ppo_adv_normalize")
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[3]) and torch.equal(outs[1], outs[4]) and torch.equal(st[0], st[1])
    assert torch.equal(outs[2], outs[5])
    assert not ws[:256].any()
