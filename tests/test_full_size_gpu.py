"""GPU parity at BASELINE.json's full per-GPU sizes (256 steps x 4096 envs x 128 horizons) through size-independent
properties, plus spot checks of a few envs against the oracle (the oracle cannot run the full size in seconds)."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import _lib, rollout
from ppo_b200 import returns_truncated as rt
from util import TOL, rel_err

pytestmark = pytest.mark.gpu

STEPS, ENVS, HEADS = 256, 4096, 128
GAMMA, TVF_GAMMA = 0.999, 0.99997


@pytest.fixture(scope="module")
def full():
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    hz = oracle.get_value_head_horizons(HEADS, 30000)
    r = torch.randint(-1, 3, (STEPS, ENVS), device=dev, generator=g).float()
    d = (torch.rand((STEPS, ENVS), device=dev, generator=g) < 3 / 101).view(torch.uint8)
    vs = torch.randn((STEPS + 1, ENVS, HEADS), device=dev, generator=g) * 0.4 + 0.1
    vs[:, :, 0] = 0
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, STEPS, 20, 8, seed=1)
    rp = rollout.ReturnPass(STEPS, ENVS, hz, GAMMA, TVF_GAMMA, 0.95, 0.95)
    rp.set_samples(table)
    rp.run(r, None, d, vs)
    torch.cuda.synchronize()
    assert rp.tvf_kernel in ("tslab", "qslab")  # (PATH_AUTO; the quad slab kernel with PPO_TVF_QSLAB=1)
    return dict(hz=hz, r=r, d=d, vs=vs, table=table, rp=rp)


def test_spot_envs_match_the_oracle(full):
    f, rp = full, full["rp"]
    envs = [0, 1, 147, 148, 2047, 4095]  # first / last env of a CTA, of the grid
    r = f["r"][:, envs].cpu().numpy()
    d = f["d"][:, envs].cpu().numpy().astype(bool)
    vs = f["vs"][:, envs].cpu().numpy()
    ref = oracle.sampled_return_multi(TVF_GAMMA, r, d, f["hz"], f["hz"], vs, n_step_samples=f["table"])
    assert rel_err(rp.tvf_returns[:, envs].cpu().numpy(), ref) <= TOL
    v = oracle.get_rediscounted_value_estimate(vs.reshape(-1, HEADS), TVF_GAMMA, GAMMA, f["hz"]).reshape(STEPS + 1, -1)
    assert rel_err(rp.value_estimates[:, envs].cpu().numpy(), v) <= TOL
    got_v = rp.value_estimates[:, envs].cpu().numpy()
    assert rel_err(rp.advantage[:, envs].cpu().numpy(), oracle.gae(r, got_v[:-1], got_v[-1], d, GAMMA, 0.95)) <= TOL
    assert rel_err(rp.ext_returns[:, envs].cpu().numpy(),
                   oracle.td_lambda(r, got_v[:-1], got_v[-1], d, GAMMA, 0.95)) <= TOL


def test_statistics_and_normalisation_are_consistent(full):
    rp = full["rp"]
    adv = rp.advantage.double()
    s = rp.stats.cpu().numpy()
    assert s[_lib.STAT_COUNT] == STEPS * ENVS
    np.testing.assert_allclose(s[_lib.STAT_ADV_SUM], adv.sum().item(), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(s[_lib.STAT_ADV_SUMSQ], (adv * adv).sum().item(), rtol=1e-9)
    assert s[_lib.STAT_ADV_ZEROS] == (rp.advantage == 0).sum().item()
    assert s[_lib.STAT_RET_ZEROS] == (rp.ext_returns == 0).sum().item()
    norm = rp.advantages.double()
    assert abs(norm.mean().item()) <= 1e-5 and abs(norm.std(unbiased=False).item() - 1.0) <= 1e-5
    last = rp.tvf_returns[:, :, -1].double()
    t = rp.tvf_stats.cpu().numpy()
    assert t[_lib.STAT_COUNT] == STEPS * ENVS
    np.testing.assert_allclose(t[_lib.STAT_ADV_SUM], last.sum().item(), rtol=1e-9)
    assert t[_lib.STAT_ADV_MIN] == last.min().item() and t[_lib.STAT_ADV_MAX] == last.max().item()
    assert t[_lib.STAT_ADV_ZEROS] == (last == 0).sum().item()


def test_linearity_is_exact_for_powers_of_two(full):
    """Targets are linear in (rewards, values): doubling both doubles every target bit for bit."""
    f = full
    got = rt._calculate_sampled_return_multi_fast(TVF_GAMMA, f["r"] * 2, f["d"], f["hz"], f["hz"], f["vs"] * 2,
                                                  n_step_samples=f["table"])
    assert torch.equal(got, f["rp"].tvf_returns * 2)


def test_env_blocks_and_zero_horizon(full):
    """Envs are independent (any env block alone gives the same bits) and horizon 0 is identically zero."""
    f, rp = full, full["rp"]
    lo, hi = 1000, 1000 + 300
    part = rt._calculate_sampled_return_multi_fast(TVF_GAMMA, f["r"][:, lo:hi], f["d"][:, lo:hi], f["hz"], f["hz"],
                                                   f["vs"][:, lo:hi], n_step_samples=f["table"])
    assert torch.equal(part, rp.tvf_returns[:, lo:hi])
    assert f["hz"][0] == 0 and not rp.tvf_returns[:, :, 0].any()
    # the h = 1 target is the one-step reward r_t (no bootstrap, rl/returns_truncated.py:580-620).  The transposed slab
    # kernel forms it as R(t) - gamma R(t+1) from float32 reward-to-go values, so it is exact to a few ulp of max|R|,
    # not bit for bit (error model); the ring kernel's prefix sums give it exactly.
    k1 = int(np.where(f["hz"] == 1)[0][0])
    err = (rp.tvf_returns[:, :, k1] - f["r"]).abs().max().item()
    assert err <= TOL * rp.tvf_returns.abs().max().item(), err
    ring = rt._calculate_sampled_return_multi_fast(TVF_GAMMA, f["r"][:, :64], f["d"][:, :64], f["hz"], f["hz"],
                                                   f["vs"][:, :64], n_step_samples=f["table"], path=rt.PATH_RING)
    assert torch.equal(ring[:, :, k1], f["r"][:, :64])


def test_kernels_agree_at_full_size(full):
    """The transposed slab kernel and the ring kernel (independent formulations) agree within the tolerance."""
    f = full
    ring = rt._calculate_sampled_return_multi_fast(TVF_GAMMA, f["r"], f["d"], f["hz"], f["hz"], f["vs"],
                                                   n_step_samples=f["table"], path=rt.PATH_RING)
    diff = (ring - f["rp"].tvf_returns).abs().max().item()
    assert diff <= TOL * ring.abs().max().item()
