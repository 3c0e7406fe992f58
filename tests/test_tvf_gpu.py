"""GPU parity: ppo_b200.returns_truncated (CUDA, through the C ABI) against the oracle and golden vectors.

Tolerance: the reference's own criterion max|d| <= 1e-5 * max|ref| (tests/test_tvf.py:46); horizon
indexing is bit-exact by construction (host tables) and is checked separately on the CPU.
"""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import returns_truncated as rt
from ppo_b200 import synthetic
from util import TOL, rel_err

pytestmark = pytest.mark.gpu

PATHS = [rt.PATH_GENERIC, rt.PATH_AUTO]


def _tvf_args(g):
    keys = ['gamma', 'rewards', 'dones', 'required_horizons', 'value_sample_horizons', 'value_samples']
    a = {k: g["tvf_in_" + k] for k in keys}
    a['gamma'] = float(a['gamma'])
    return a


@pytest.mark.parametrize("path", PATHS)
def test_reference_estimator_cases(golden, path):
    """The four cases of the reference's tests/test_tvf.py:116-119 on the golden inputs."""
    a = _tvf_args(golden)
    for n in (1, 8, 48):
        got = rt._calculate_sampled_return_multi_fast(n_step_list=[n], path=path, **a)
        assert got.dtype == np.float32 and got.shape == golden[f"tvf_fast_n{n}"].shape
        assert rel_err(got, golden[f"tvf_fast_n{n}"]) <= TOL
        # against the reference's slow triple-loop estimator with the reference's criterion (:46)
        slow = golden[f"tvf_slow_n{n}"]
        assert np.abs(slow - got).max() <= 1e-5 * slow.max()
    s = golden["tvf_in_samples"]
    got = rt._calculate_sampled_return_multi_fast(n_step_samples=s, path=path, **a)
    assert rel_err(got, golden["tvf_fast_samples"]) <= TOL
    got = rt._calculate_sampled_return_multi_fast(n_step_samples=s, use_log_interpolation=True, path=path, **a)
    assert rel_err(got, golden["tvf_fast_samples_log"]) <= TOL


@pytest.mark.parametrize("dist,mode", [("fixed", "standard"), ("exponential", "standard"),
                                       ("exponential", "advanced"), ("uniform", "advanced"),
                                       ("hyperbolic", "clipped"), ("quadratic", "adaptive"),
                                       ("exponential", "mcx"), ("exponential", "full")])
def test_dispatcher_golden(golden, dist, mode):
    a = _tvf_args(golden)
    got = rt.get_return_estimate(dist, mode, n_step=6, max_samples=5, seed=17, **a)
    assert rel_err(got, golden[f"tvf_gre_{dist}_{mode}"]) <= TOL


def test_production_horizons_golden(golden):
    g = golden
    hz, r, d, vs = g["prod_in_horizons"], g["prod_in_rewards"], g["prod_in_dones"], g["prod_in_vs"]
    got = rt.get_return_estimate("exponential", "advanced", 0.99997, r, d, hz, hz, vs, n_step=20, max_samples=8, seed=1)
    assert rel_err(got, g["prod_adv"]) <= TOL
    got = rt.get_return_estimate("fixed", "advanced", 0.99997, r, d, hz, hz, vs, n_step=40, max_samples=8, seed=1)
    assert rel_err(got, g["prod_fixed"]) <= TOL
    got = rt.get_return_estimate("exponential", "advanced", 0.99997, r, d, hz, hz, vs, n_step=20, max_samples=8,
                                 use_log_interpolation=True, seed=1)
    assert rel_err(got, g["prod_adv_log"]) <= TOL


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("steps,envs,heads,max_h,n_step,samples,done_p", [
    (128, 64, 128, 30000, 20, 8, 0.03),    # c3 shape, fewer envs
    (256, 24, 128, 30000, 20, 8, 0.03),    # c5 shape, fewer envs
    (256, 8, 128, 30000, 20, 8, 0.0),      # no terminals at all
    (64, 16, 128, 30000, 64, 1, 0.3),      # n-step == N, many terminals
    (33, 5, 16, 200, 7, 3, 0.1),           # ragged sizes
    (100, 7, 128, 30000, 20, 8, 0.05),     # slab kernel with a partial last TMA box and a partial last block
    (45, 3, 32, 1000, 30, 4, 0.1),         # slab kernel, V = 32
    (70, 301, 64, 5000, 9, 5, 0.02),       # slab kernel, V = 64, more envs than SMs
    (1, 3, 8, 50, 2, 2, 0.0),              # a single step
])
def test_vs_oracle(path, steps, envs, heads, max_h, n_step, samples, done_p):
    hz = oracle.get_value_head_horizons(heads, max_h)
    r, d, _ = synthetic.rollout(steps, envs, seed=steps + envs, done_prob=done_p)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=steps)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, n_step, samples, seed=5)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    got = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=path)
    assert rel_err(got, ref) <= TOL
    fixed = np.zeros([len(hz), 1], dtype=np.int32) + min(n_step, steps)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=fixed)
    got = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=fixed, path=path)
    assert rel_err(got, ref) <= TOL


@pytest.mark.parametrize("path", [rt.PATH_TSLAB, rt.PATH_RING])
@pytest.mark.parametrize("steps,envs,heads,max_h,n_step,samples,done_p,use_log", [
    (128, 64, 128, 30000, 20, 8, 0.03, False),   # c3 shape, fewer envs
    (256, 24, 128, 30000, 20, 8, 0.03, False),   # c5 shape, fewer envs
    (256, 8, 128, 30000, 20, 8, 0.0, False),     # no terminals at all: the reward-to-go spans the whole rollout
    (256, 160, 128, 30000, 20, 8, 0.03, True),   # more envs than SMs (in-place slab replacement), log interpolation
    (64, 16, 128, 30000, 64, 1, 0.3, False),     # n-step == N, many terminals
    (100, 7, 128, 30000, 20, 8, 0.05, False),    # partial last TMA box and partial last block
    (45, 3, 32, 1000, 30, 4, 0.1, False),        # V = 32
    (70, 301, 64, 5000, 9, 5, 0.02, False),      # V = 64, more envs than SMs
    (31, 2, 32, 1000, 31, 6, 0.0, False),        # a single (partial) block, windows as long as the rollout
])
def test_slab_kernels_vs_oracle(path, steps, envs, heads, max_h, n_step, samples, done_p, use_log):
    """The two shared-memory slab kernels explicitly (PATH_RING / PATH_TSLAB fail instead of falling back)."""
    hz = oracle.get_value_head_horizons(heads, max_h)
    if len(hz) % 4:
        pytest.skip("horizon count is not a multiple of 4")
    r, d, _ = synthetic.rollout(steps, envs, seed=steps + envs, done_prob=done_p)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=steps)
    for mode in ("advanced", "standard"):
        table = oracle.draw_n_step_samples("exponential", mode, hz, steps, n_step, samples, seed=5)
        ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table,
                                          use_log_interpolation=use_log)
        got = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table,
                                                      use_log_interpolation=use_log, path=path)
        assert rel_err(got, ref) <= TOL


def test_tslab_fused_statistics_and_large_rewards():
    """Statistics of the longest horizon fused into the transposed slab kernel; rewards of one sign and no
    terminals make the reward-to-go as large as it gets relative to the targets."""
    steps, envs = 256, 200
    hz = oracle.get_value_head_horizons(128, 30000)
    rng = np.random.RandomState(3)
    r = (rng.rand(steps, envs) * 3 + 1).astype(np.float32)
    d = np.zeros((steps, envs), dtype=bool)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=4)
    r[:, 5:11], vs[:, 5:11] = 0.0, 0.0  # silent envs: exact zeros, counted for the logger's sparsity (rl/logger.py:201-229)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=2)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    res, stats = rt._launch(0.99997, r, d, hz, hz, vs, table, False, path=rt.PATH_TSLAB, with_stats=True)
    from ppo_b200 import _lib
    assert _lib.load().ppo_tvf_last_path() == (rt.PATH_TSLAB | 0x100)  # one launch, statistics fused
    assert rel_err(res, ref) <= TOL
    # PATH_AUTO picks a slab kernel for eight n-steps per horizon (the transposed slab kernel; the quad slab kernel with
    # PPO_TVF_QSLAB=1; statistics fused either way), and the ring kernel for a fixed n-step
    rt._launch(0.99997, r, d, hz, hz, vs, table, False, with_stats=True)
    assert _lib.load().ppo_tvf_last_path() in (rt.PATH_TSLAB | 0x100, rt.PATH_QSLAB | 0x100)
    rt._launch(0.99997, r, d, hz, hz, vs, np.zeros([len(hz), 1], dtype=np.int32) + 20, False)
    assert _lib.load().ppo_tvf_last_path() == rt.PATH_RING
    s = stats.cpu().numpy()
    last = res[:, :, -1].astype(np.float64)
    assert s[0] == last.size and s[3] == last.min() and s[4] == last.max()
    assert s[12] == np.count_nonzero(last == 0) == 6 * steps
    np.testing.assert_allclose(s[1], last.sum(), rtol=1e-12)
    np.testing.assert_allclose(s[2], (last ** 2).sum(), rtol=1e-12)


def test_numpy_and_native_tables_agree_through_the_kernels(golden):
    a = _tvf_args(golden)
    s = golden["tvf_in_samples"]
    steps = a['rewards'].shape[0]
    x = rt._launch(a['gamma'], a['rewards'], a['dones'], a['required_horizons'], a['value_sample_horizons'],
                   a['value_samples'], s, False, path=rt.PATH_GENERIC, numpy_tables=True)
    y = rt._launch(a['gamma'], a['rewards'], a['dones'], a['required_horizons'], a['value_sample_horizons'],
                   a['value_samples'], s, False, path=rt.PATH_GENERIC, numpy_tables=False)
    assert np.array_equal(x, y) and steps > 0


def test_device_tensors_strided_values_and_stats():
    steps, envs = 48, 40
    hz = oracle.get_value_head_horizons(32, 1000)
    r, d, _ = synthetic.rollout(steps, envs, seed=2)
    full = np.stack([synthetic.value_samples(steps, envs, len(hz), seed=2),
                     synthetic.value_samples(steps, envs, len(hz), seed=3)], axis=-1)  # [N+1, A, K, VH]
    vs = full[..., 1]  # rl/tvf.py:248 strided view
    r[:, 3:9], full[:, 3:9] = 0.0, 0.0
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 10, 4, seed=1)
    ref = oracle.sampled_return_multi(0.999, r, d, hz, hz, vs, n_step_samples=table)
    got = rt._calculate_sampled_return_multi_fast(
        0.999, torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda(), hz, hz,
        torch.from_numpy(full).cuda()[..., 1], n_step_samples=table)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    assert rel_err(got.cpu().numpy(), ref) <= TOL
    res, stats = rt._launch(0.999, r, d, hz, hz, vs, table, False, with_stats=True)
    s = stats.cpu().numpy()
    last = res[:, :, -1].astype(np.float64)
    assert s[0] == last.size and s[3] == last.min() and s[4] == last.max()
    assert s[12] == np.count_nonzero(last == 0) >= 6 * steps  # also zero reward on a terminal step
    np.testing.assert_allclose(s[1], last.sum(), rtol=1e-12)
    np.testing.assert_allclose(s[2], (last ** 2).sum(), rtol=1e-12)


def test_error_behaviour(golden):
    a = _tvf_args(golden)
    with pytest.raises(ValueError):
        rt.get_return_estimate("nope", "standard", **a)
    with pytest.raises(ValueError):
        rt.get_return_estimate("uniform", "nope", **a)
    with pytest.raises(AssertionError):  # rl/returns_truncated.py:597
        rt._calculate_sampled_return_multi_fast(n_step_list=[49], **a)
    b = dict(a)
    b['required_horizons'] = np.asarray([5000], dtype=np.int32)  # beyond the last value horizon: :157 IndexError
    with pytest.raises(IndexError):
        rt._calculate_sampled_return_multi_fast(n_step_list=[4], **b)


@pytest.mark.parametrize("seed", range(10))
def test_slab_kernels_fuzz_against_the_generic_kernel(seed):
    """Random shapes (steps, envs, horizon counts 32 / 64 / 96 / 128, sample counts, terminal rates, random horizon
    sets): both slab kernels against the generic kernel, which the other tests hold to the oracle."""
    rng = np.random.RandomState(500 + seed)
    steps = int(rng.randint(2, 257))
    envs = int(rng.choice([1, 3, 50, 149, 300, 700]))
    n_val = int(rng.choice([32, 64, 128]))
    n_req = int(rng.choice([32, 64, 96, 128]))
    vsh = np.unique(np.concatenate([[0], rng.randint(1, int(rng.choice([200, 5000, 30000])), size=4 * n_val)]))[:n_val]
    while len(vsh) < n_val:  # tiny ranges: pad with consecutive horizons
        vsh = np.unique(np.concatenate([vsh, vsh[-1:] + 1 + np.arange(n_val - len(vsh))]))
    req = np.sort(rng.randint(0, vsh[-1] + 1, size=n_req))
    c = int(rng.randint(1, 9))
    table = rng.randint(1, steps + 1, size=[n_req, c])
    r = rng.randint(-1, 3, size=[steps, envs]).astype(np.float32)
    d = rng.rand(steps, envs) < float(rng.choice([0.0, 0.03, 0.3]))
    vs = rng.normal(0.1, 0.4, size=[steps + 1, envs, n_val]).astype(np.float32)
    use_log = bool(seed % 2)
    ref = rt._calculate_sampled_return_multi_fast(0.9997, r, d, req, vsh, vs, n_step_samples=table,
                                                  use_log_interpolation=use_log, path=rt.PATH_GENERIC)
    for path in (rt.PATH_TSLAB, rt.PATH_RING):
        got = rt._calculate_sampled_return_multi_fast(0.9997, r, d, req, vsh, vs, n_step_samples=table,
                                                      use_log_interpolation=use_log, path=path)
        assert rel_err(got, ref) <= TOL, (path, steps, envs, n_val, n_req, c)
