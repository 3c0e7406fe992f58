"""GPU parity at the edges the reference's own tests poke at (tests/test_tvf.py, rl/unit_tests.py) and at the limits
of the kernels' fast paths: empty and single-element rollouts, every step terminal, rollouts longer than a slab,
horizon counts that no slab kernel takes, repeated horizons."""
import numpy as np
import pytest

import oracle
from ppo_b200 import _lib, normalization as nz, returns, returns_truncated as rt, rollout, synthetic, vtrace
from util import TOL, rel_err

pytestmark = pytest.mark.gpu


def test_empty_rollouts():
    hz = oracle.get_value_head_horizons(32, 1000)
    for steps, envs in ((0, 5), (4, 0)):
        r = np.zeros([steps, envs], np.float32)
        d = np.zeros([steps, envs], bool)
        vs = np.zeros([steps + 1, envs, len(hz)], np.float32)
        if steps:  # the reference asserts 1 <= n-step <= N, so N = 0 has no valid sample table
            out = rt._calculate_sampled_return_multi_fast(0.99, r, d, hz, hz, vs, n_step_list=[2])
            assert out.shape == (steps, envs, len(hz)) and out.dtype == np.float32
        assert returns.gae(r, vs[:-1, :, -1], vs[-1, :, -1], d, 0.99).shape == (steps, envs)
        assert returns.calculate_bootstrapped_returns(r, d, vs[-1, :, -1], 0.99).shape == (steps, envs)
        pol = np.zeros([steps, envs, 4], np.float32)
        v, a, c = vtrace.importance_sampling_v_trace(pol, pol, np.zeros([steps, envs], np.int64), r, d,
                                                     vs[:-1, :, -1], vs[-1, :, -1], 0.99)
        assert v.shape == a.shape == c.shape == (steps, envs)
    res = nz.IntrinsicReturns(3).calculate_intrinsic_returns(np.zeros([0, 3], np.float32), np.zeros([1, 3], np.float32),
                                                             np.zeros([0, 3], bool), 0.99)
    assert res["int_advantage"].shape == (0, 3)


@pytest.mark.parametrize("path", [rt.PATH_AUTO, rt.PATH_GENERIC, rt.PATH_TSLAB, rt.PATH_RING])
def test_every_step_terminal_and_no_reward(path):
    """All-terminal rollouts: every n-step window is cut after one step, nothing bootstraps (out = r);
    all-zero rewards and values: exact zeros everywhere."""
    steps, envs = 96, 40
    hz = oracle.get_value_head_horizons(32, 1000)
    r, _, _ = synthetic.rollout(steps, envs, seed=1)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=1)
    d = np.ones([steps, envs], bool)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 10, 4, seed=2)
    got = rt._calculate_sampled_return_multi_fast(0.999, r, d, hz, hz, vs, n_step_samples=table, path=path)
    assert rel_err(got, oracle.sampled_return_multi(0.999, r, d, hz, hz, vs, n_step_samples=table)) <= TOL
    assert np.array_equal(got[:, :, 1:], np.broadcast_to(r[:, :, None], got[:, :, 1:].shape)) and not got[:, :, 0].any()
    zero = rt._calculate_sampled_return_multi_fast(0.999, r * 0, d, hz, hz, vs * 0, n_step_samples=table, path=path)
    assert not zero.any()


@pytest.mark.parametrize("steps,heads,max_h,expect", [(512, 32, 1000, "tslab"), (600, 32, 1000, "generic"),
                                                      (64, 17, 300, "generic"), (64, 48, 3000, "generic")])
def test_auto_path_limits(steps, heads, max_h, expect):
    """PATH_AUTO at the limits of the slab kernels: 512 steps still fit one slab (V = 32), 600 do not; horizon counts
    other than 32 / 64 / 128 take the generic kernel.  Same targets either way."""
    hz = oracle.get_value_head_horizons(heads, max_h)
    envs = 6
    r, d, _ = synthetic.rollout(steps, envs, seed=steps)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=steps)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    got = rt._calculate_sampled_return_multi_fast(0.9999, r, d, hz, hz, vs, n_step_samples=table)
    kernel = {1: "generic", 2: "ring", 3: "tslab"}[_lib.load().ppo_tvf_last_path() & 0xff]
    assert kernel == expect
    assert rel_err(got, oracle.sampled_return_multi(0.9999, r, d, hz, hz, vs, n_step_samples=table)) <= TOL


def test_repeated_and_unsorted_required_horizons():
    """Required horizons may repeat and need not be the value horizons (tests/test_tvf.py:20-23: geomspace cast to int
    repeats 1, the value axis starts 0, 0, 0, 0)."""
    steps, envs = 48, 9
    req = np.asarray([1, 1, 2, 7, 7, 100, 1024, 3], dtype=np.int32)
    vsh = (np.geomspace(1, 1024, num=32).astype(np.int32) - 1)
    r, d, _ = synthetic.rollout(steps, envs, seed=8)
    vs = synthetic.value_samples(steps, envs, len(vsh), seed=8)
    table = np.random.RandomState(3).randint(1, steps + 1, size=[len(req), 5])
    ref = oracle.sampled_return_multi(0.9997, r, d, req, vsh, vs, n_step_samples=table)
    got = rt._calculate_sampled_return_multi_fast(0.9997, r, d, req, vsh, vs, n_step_samples=table)
    assert rel_err(got, ref) <= TOL
    assert np.array_equal(got[:, :, 0], got[:, :, 1])  # h = 1 twice: every sample clamps to n = 1


def test_single_env_single_step_pass():
    hz = oracle.get_value_head_horizons(32, 1000)
    r, d, _ = synthetic.rollout(1, 1, seed=4)
    vs = synthetic.value_samples(1, 1, len(hz), seed=4)
    res = rollout.calculate_returns(r, None, d, 0.999, 0.95, 0.95,
                                    tvf=dict(value_samples=vs, horizons=hz, gamma=0.999, mode="advanced",
                                             distribution="fixed", n_step=1, max_samples=1))
    v = vs[:, :, -1]
    assert rel_err(res["advantage"], oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)) <= TOL
    assert np.asarray(res["advantages"]).shape == (1,) and np.isfinite(np.asarray(res["advantages"])).all()
    assert rel_err(res["tvf_returns"], oracle.get_return_estimate("fixed", "advanced", 0.999, r, d, hz, hz, vs, n_step=1,
                                                                  max_samples=1)) <= TOL
