"""GPU parity of the quad slab kernel (PPO_TVF_QSLAB, csrc/tvf_qslab.cu): the explicit path fails instead of falling
back, so every case below really runs that kernel.  It takes 128 or 256 steps and 128 required horizons on 128 value
horizons (the north-star shape and config c3); everything else stays on the transposed slab kernel."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import _lib
from ppo_b200 import returns_truncated as rt
from ppo_b200 import synthetic
from ppo_b200 import tvf
from util import TOL, rel_err

# a kernel built from hand-rolled barriers can hang instead of failing: kill the run rather than the GPU budget
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120, method="thread")]


@pytest.mark.parametrize("steps,envs,max_h,n_step,samples,done_p,use_log", [
    (256, 24, 30000, 20, 8, 0.03, False),   # the bench shape, fewer envs
    (128, 64, 30000, 20, 8, 0.03, False),   # c3 shape
    (256, 8, 30000, 20, 8, 0.0, False),     # no terminals: every window runs to the end of the rollout
    (256, 400, 30000, 20, 8, 0.03, False),  # more envs than SMs: several envs per CTA, in-place column refills
    (256, 1, 30000, 20, 8, 0.03, False),    # one env: one CTA
    (256, 40, 30000, 20, 8, 0.3, False),    # many terminals
    (256, 12, 30000, 20, 8, 0.03, True),    # log interpolation
    (128, 9, 5000, 9, 5, 0.02, False),      # short horizons: most quads read the -R column
    (256, 5, 1000, 30, 4, 0.1, False),      # long windows on short horizons
    (128, 3, 30000, 128, 6, 0.0, False),    # windows as long as the rollout
    (256, 7, 30000, 3, 15, 0.05, False),    # 15 samples per horizon
])
def test_qslab_vs_oracle(steps, envs, max_h, n_step, samples, done_p, use_log):
    hz = oracle.get_value_head_horizons(128, max_h)
    assert len(hz) == 128
    r, d, _ = synthetic.rollout(steps, envs, seed=steps + envs, done_prob=done_p)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=steps)
    for mode in ("advanced", "standard"):
        table = oracle.draw_n_step_samples("exponential", mode, hz, steps, n_step, samples, seed=5)
        if samples == 15:  # few distinct n-steps with large multiplicities (<= 8 distinct per horizon: 1024 entries)
            table = np.random.RandomState(6).randint(1, 8, size=table.shape)
        ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table, use_log_interpolation=use_log)
        got = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table,
                                                      use_log_interpolation=use_log, path=rt.PATH_QSLAB)
        assert _lib.load().ppo_tvf_last_path() & 0xff == rt.PATH_QSLAB
        assert rel_err(got, ref) <= TOL, mode


def test_qslab_fused_statistics():
    steps, envs = 256, 200
    hz = oracle.get_value_head_horizons(128, 30000)
    rng = np.random.RandomState(3)
    r = (rng.rand(steps, envs) * 3 + 1).astype(np.float32)
    d = np.zeros((steps, envs), dtype=bool)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=4)
    r[:, 5:11], vs[:, 5:11] = 0.0, 0.0
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=2)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    res, stats = rt._launch(0.99997, r, d, hz, hz, vs, table, False, path=rt.PATH_QSLAB, with_stats=True)
    last_path = _lib.load().ppo_tvf_last_path()
    assert last_path & 0xff == rt.PATH_QSLAB and last_path & 0x100
    assert rel_err(res, ref) <= TOL
    s = stats.cpu().numpy()
    last = res[:, :, -1].astype(np.float64)
    assert s[0] == last.size and s[3] == last.min() and s[4] == last.max()
    assert s[12] == np.count_nonzero(last == 0) == 6 * steps
    np.testing.assert_allclose(s[1], last.sum(), rtol=1e-12)
    np.testing.assert_allclose(s[2], (last ** 2).sum(), rtol=1e-12)


@pytest.mark.parametrize("gamma", [0.999, 0.99997])
@pytest.mark.parametrize("steps", [128, 256])
def test_qslab_value_estimate_from_the_same_launch(gamma, steps):
    """ppo_tvf_targets_ex: the rediscounted curve (gamma != tvf gamma) or the last horizon (same gamma) comes out of
    the target launch (flag 0x200), rl/tvf.py:303-360."""
    envs, tvf_gamma = 310, 0.99997
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=8)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=9)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=2)
    out = torch.empty((steps + 1, envs), dtype=torch.float32, device="cuda")
    w = None if gamma == tvf_gamma else tvf.rediscount_weights(tvf_gamma, gamma, [int(h) for h in hz])
    got = rt._launch(tvf_gamma, r, d, hz, hz, vs, table, False, path=rt.PATH_QSLAB, value_out=out, value_weights=w)
    last = _lib.load().ppo_tvf_last_path()
    assert (last & 0xff) == rt.PATH_QSLAB and (last & 0x200)
    assert rel_err(got, oracle.sampled_return_multi(tvf_gamma, r, d, hz, hz, vs, n_step_samples=table)) <= TOL
    ref = oracle.get_rediscounted_value_estimate(vs.reshape(-1, len(hz)), tvf_gamma, gamma, hz).reshape(steps + 1, envs)
    if gamma == tvf_gamma:
        assert np.array_equal(out.cpu().numpy(), vs[:, :, -1])
    else:
        assert rel_err(out.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("seed", range(8))
def test_qslab_fuzz_against_the_generic_kernel(seed):
    """Random env counts, sample counts, terminal rates and horizon sets at 128 / 256 steps."""
    rng = np.random.RandomState(1900 + seed)
    steps = int(rng.choice([128, 256]))
    envs = int(rng.choice([1, 3, 50, 149, 300]))
    n_val = n_req = 128
    vsh = np.unique(np.concatenate([[0], rng.randint(1, int(rng.choice([200, 5000, 30000])), size=4 * n_val)]))[:n_val]
    while len(vsh) < n_val:
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
    got = rt._calculate_sampled_return_multi_fast(0.9997, r, d, req, vsh, vs, n_step_samples=table,
                                                  use_log_interpolation=use_log, path=rt.PATH_QSLAB)
    assert rel_err(got, ref) <= TOL, (steps, envs, c)


@pytest.mark.parametrize("kind", ["f32", "none", "cuda"])
def test_qslab_done_kinds_and_device_inputs(kind):
    """float32 terminals (rl/returns_truncated.py accepts 0 / 1 masks of either dtype), no terminals at all, and CUDA
    tensors in -> CUDA tensor out."""
    steps, envs = 256, 77
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=31, done_prob=0.04)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=32)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=7)
    if kind == "none":
        d = np.zeros_like(d)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    if kind == "f32":
        got = rt._calculate_sampled_return_multi_fast(0.99997, r, d.astype(np.float32), hz, hz, vs, n_step_samples=table,
                                                      path=rt.PATH_QSLAB)
    elif kind == "none":
        got = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=rt.PATH_QSLAB)
    else:
        out = rt._calculate_sampled_return_multi_fast(0.99997, torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda(), hz, hz,
                                                      torch.from_numpy(vs).cuda(), n_step_samples=table, path=rt.PATH_QSLAB)
        assert isinstance(out, torch.Tensor) and out.is_cuda
        got = out.cpu().numpy()
    assert _lib.load().ppo_tvf_last_path() & 0xff == rt.PATH_QSLAB
    assert rel_err(got, ref) <= TOL


@pytest.mark.parametrize("seed", range(3))
def test_qslab_unsorted_required_horizons(seed):
    """Required horizons in any order: the column groups a unit reads are then not monotone in the order the units are
    drawn -- the case the arrival counts behind the parity waits exist for.  Against the generic kernel."""
    rng = np.random.RandomState(700 + seed)
    steps, envs = 256, int(rng.choice([5, 160, 301]))
    vsh = oracle.get_value_head_horizons(128, 30000)
    req = rng.permutation(vsh)  # every value horizon once, shuffled
    table = rng.randint(1, 60, size=[128, 6])
    r = rng.randint(-1, 3, size=[steps, envs]).astype(np.float32)
    d = rng.rand(steps, envs) < 0.03
    vs = rng.normal(0.1, 0.4, size=[steps + 1, envs, 128]).astype(np.float32)
    ref = rt._calculate_sampled_return_multi_fast(0.9997, r, d, req, vsh, vs, n_step_samples=table, path=rt.PATH_GENERIC)
    got = rt._calculate_sampled_return_multi_fast(0.9997, r, d, req, vsh, vs, n_step_samples=table, path=rt.PATH_QSLAB)
    assert _lib.load().ppo_tvf_last_path() & 0xff == rt.PATH_QSLAB
    assert rel_err(got, ref) <= TOL


def test_qslab_is_deterministic_and_matches_the_transposed_slab_kernel():
    """Same arithmetic per target as tvf_tslab.cu (same constants, same order of the entries): bit-identical targets,
    and bit-identical from run to run although quads are handed out dynamically."""
    steps, envs = 256, 333
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=21)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=22)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=3)
    a = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=rt.PATH_QSLAB)
    b = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=rt.PATH_QSLAB)
    c = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=rt.PATH_TSLAB)
    assert np.array_equal(a, b)
    assert np.array_equal(a, c)


def test_path_auto_selection(monkeypatch):
    """PATH_AUTO: the quad slab kernel at 256 steps and up to 4096 envs per launch, the transposed slab kernel beyond
    (the env-count edge) and at 128 steps; PPO_TVF_QSLAB=0 / 1 forces the choice where both apply."""
    lib = _lib.load()
    hz = oracle.get_value_head_horizons(128, 30000)
    dev = torch.device("cuda")

    def run(steps, envs):
        table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=5)
        ctab, keep = rt._upload_tables(rt.build_tables(0.99997, steps, hz, hz, table), dev)
        r = torch.zeros((steps, envs), device=dev)
        d = torch.zeros((steps, envs), dtype=torch.uint8, device=dev)
        vs = torch.zeros((steps + 1, envs, 128), device=dev)
        out = torch.empty((steps, envs, 128), device=dev)
        rc = lib.ppo_tvf_targets(r.data_ptr(), envs, d.data_ptr(), _lib.DONE_U8, envs, vs.data_ptr(), envs * 128, 128, steps,
                                 envs, 128, 128, ctypes.byref(ctab), rt.PATH_AUTO, out.data_ptr(), 0, 0, 0,
                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        return lib.ppo_tvf_last_path() & 0xff

    monkeypatch.delenv("PPO_TVF_QSLAB", raising=False)
    assert run(256, 300) == rt.PATH_QSLAB and run(256, 4096) == rt.PATH_QSLAB
    assert run(256, 4100) == rt.PATH_TSLAB and run(128, 300) == rt.PATH_TSLAB
    monkeypatch.setenv("PPO_TVF_QSLAB", "0")
    assert run(256, 300) == rt.PATH_TSLAB
    monkeypatch.setenv("PPO_TVF_QSLAB", "1")
    assert run(128, 300) == rt.PATH_QSLAB and run(256, 4100) == rt.PATH_QSLAB


def test_qslab_declines_what_it_cannot_take():
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(100, 4, seed=1)
    vs = synthetic.value_samples(100, 4, len(hz), seed=1)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, 100, 20, 8, seed=5)
    with pytest.raises(ValueError, match="128 or 256"):
        rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=rt.PATH_QSLAB)
    hz64 = oracle.get_value_head_horizons(64, 1000)
    r2, d2, _ = synthetic.rollout(128, 4, seed=1)
    with pytest.raises(ValueError, match="128 required horizons"):
        rt._calculate_sampled_return_multi_fast(0.99997, r2, d2, hz64, hz64, synthetic.value_samples(128, 4, 64, seed=1),
                                                n_step_samples=oracle.draw_n_step_samples("exponential", "advanced", hz64,
                                                                                           128, 20, 8, seed=5),
                                                path=rt.PATH_QSLAB)
    # lean tables (the transposed slab section only: what ReturnPass uploads per pass) are all this kernel needs
    r, d, _ = synthetic.rollout(128, 4, seed=1)
    vs = synthetic.value_samples(128, 4, len(hz), seed=1)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, 128, 20, 8, seed=5)
    lean = rt.build_tables(0.99997, 128, hz, hz, table, False, sections=rt.TABLES_TSLAB)
    dev = torch.device("cuda")
    ctab, keep = rt._upload_tables(lean, dev)
    out = torch.empty((128, 4, len(hz)), device=dev)
    tr, td, tvs = torch.from_numpy(r).to(dev), torch.from_numpy(d).to(dev).view(torch.uint8), torch.from_numpy(vs).to(dev)
    lib = _lib.load()
    rc = lib.ppo_tvf_targets(tr.data_ptr(), 4, td.data_ptr(), _lib.DONE_U8, 4, tvs.data_ptr(), 4 * len(hz), len(hz), 128, 4,
                             len(hz), len(hz), ctypes.byref(ctab), rt.PATH_QSLAB, out.data_ptr(), 0, 0, 0,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and lib.ppo_tvf_last_path() & 0xff == rt.PATH_QSLAB
    torch.cuda.synchronize()
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    assert rel_err(out.cpu().numpy(), ref) <= TOL
