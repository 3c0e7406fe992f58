"""GPU parity of the column slab kernel (PPO_TVF_CSLAB, csrc/tvf_cslab.cu): the explicit path fails instead of
falling back, so every case below really runs that kernel.  It takes 128 or 256 steps, 32 / 64 / 96 / 128 value
horizons and 64, 96 or 128 required horizons; PATH_AUTO does not select it (the transposed slab kernel is faster
on B200), so these tests are what keeps it honest."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import _lib
from ppo_b200 import returns_truncated as rt
from ppo_b200 import synthetic
from util import TOL, rel_err

# a kernel built from hand-rolled barriers can hang instead of failing: kill the run rather than the GPU budget
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


@pytest.mark.parametrize("steps,envs,heads,max_h,n_step,samples,done_p,use_log", [
    (256, 24, 128, 30000, 20, 8, 0.03, False),   # the bench shape, fewer envs
    (128, 64, 128, 30000, 20, 8, 0.03, False),   # c3 shape
    (256, 8, 128, 30000, 20, 8, 0.0, False),     # no terminals: every window runs to the end of the rollout
    (256, 400, 128, 30000, 20, 8, 0.03, False),  # more envs than SMs: envs are drawn from the global counter
    (256, 40, 128, 30000, 20, 8, 0.3, False),    # many terminals
    (256, 12, 128, 30000, 20, 8, 0.03, True),    # log interpolation: tails through the record tables
    (128, 9, 64, 5000, 9, 5, 0.02, False),       # V = K = 64
    (256, 5, 64, 1000, 30, 4, 0.1, False),       # V = K = 64, long windows
    (128, 3, 128, 30000, 128, 6, 0.0, False),    # windows as long as the rollout
    (256, 7, 128, 30000, 3, 15, 0.05, False),    # 15 samples per horizon (the 4-bit multiplicity table's limit)
])
def test_cslab_vs_oracle(steps, envs, heads, max_h, n_step, samples, done_p, use_log):
    hz = oracle.get_value_head_horizons(heads, max_h)
    assert len(hz) % 32 == 0
    r, d, _ = synthetic.rollout(steps, envs, seed=steps + envs, done_prob=done_p)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=steps)
    for mode in ("advanced", "standard"):
        table = oracle.draw_n_step_samples("exponential", mode, hz, steps, n_step, samples, seed=5)
        if samples == 15:  # few distinct n-steps with large multiplicities (the kernel takes <= 8 distinct per horizon)
            table = np.random.RandomState(6).randint(1, 2 * n_step + 1, size=table.shape)
        ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table, use_log_interpolation=use_log)
        got = rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table,
                                                      use_log_interpolation=use_log, path=rt.PATH_CSLAB)
        assert _lib.load().ppo_tvf_last_path() & 0xff == rt.PATH_CSLAB
        assert rel_err(got, ref) <= TOL, mode


def test_cslab_fused_statistics():
    steps, envs = 256, 200
    hz = oracle.get_value_head_horizons(128, 30000)
    rng = np.random.RandomState(3)
    r = (rng.rand(steps, envs) * 3 + 1).astype(np.float32)
    d = np.zeros((steps, envs), dtype=bool)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=4)
    r[:, 5:11], vs[:, 5:11] = 0.0, 0.0
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=2)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    res, stats = rt._launch(0.99997, r, d, hz, hz, vs, table, False, path=rt.PATH_CSLAB, with_stats=True)
    assert _lib.load().ppo_tvf_last_path() & 0xff == rt.PATH_CSLAB
    assert rel_err(res, ref) <= TOL
    s = stats.cpu().numpy()
    last = res[:, :, -1].astype(np.float64)
    assert s[0] == last.size and s[3] == last.min() and s[4] == last.max()
    assert s[12] == np.count_nonzero(last == 0) == 6 * steps
    np.testing.assert_allclose(s[1], last.sum(), rtol=1e-12)
    np.testing.assert_allclose(s[2], (last ** 2).sum(), rtol=1e-12)


@pytest.mark.parametrize("seed", range(8))
def test_cslab_fuzz_against_the_generic_kernel(seed):
    """Random env counts, horizon counts, sample counts, terminal rates and horizon sets at 128 / 256 steps."""
    rng = np.random.RandomState(900 + seed)
    steps = int(rng.choice([128, 256]))
    envs = int(rng.choice([1, 3, 50, 149, 300]))
    n_val = int(rng.choice([32, 64, 96, 128]))
    n_req = int(rng.choice([64, 96, 128]))
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
                                                  use_log_interpolation=use_log, path=rt.PATH_CSLAB)
    assert rel_err(got, ref) <= TOL, (steps, envs, n_val, n_req, c)


def test_cslab_declines_what_it_cannot_take():
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(100, 4, seed=1)
    vs = synthetic.value_samples(100, 4, len(hz), seed=1)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, 100, 20, 8, seed=5)
    with pytest.raises(ValueError, match="128 or 256"):
        rt._calculate_sampled_return_multi_fast(0.99997, r, d, hz, hz, vs, n_step_samples=table, path=rt.PATH_CSLAB)
    hz32 = oracle.get_value_head_horizons(32, 1000)  # one team job per env: not supported
    r2, d2, _ = synthetic.rollout(128, 4, seed=1)
    with pytest.raises(ValueError, match="64, 96 or 128"):
        rt._calculate_sampled_return_multi_fast(0.99997, r2, d2, hz32, hz32, synthetic.value_samples(128, 4, 32, seed=1),
                                                n_step_samples=oracle.draw_n_step_samples("exponential", "advanced", hz32,
                                                                                           128, 20, 8, seed=5),
                                                path=rt.PATH_CSLAB)
    # tables built without the column slab section (what ReturnPass uploads per pass) are refused, not misread
    r, d, _ = synthetic.rollout(128, 4, seed=1)
    vs = synthetic.value_samples(128, 4, len(hz), seed=1)
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, 128, 20, 8, seed=5)
    lean = rt.build_tables(0.99997, 128, hz, hz, table, False, sections=rt.TABLES_TSLAB)
    assert lean.layout.sections == rt.TABLES_TSLAB
    import ctypes
    dev = torch.device("cuda")
    ctab, keep = rt._upload_tables(lean, dev)
    out = torch.empty((128, 4, len(hz)), device=dev)
    tr, td, tvs = torch.from_numpy(r).to(dev), torch.from_numpy(d).to(dev).view(torch.uint8), torch.from_numpy(vs).to(dev)
    lib = _lib.load()
    rc = lib.ppo_tvf_targets(tr.data_ptr(), 4, td.data_ptr(), _lib.DONE_U8, 4, tvs.data_ptr(), 4 * len(hz), len(hz), 128, 4,
                             len(hz), len(hz), ctypes.byref(ctab), rt.PATH_CSLAB, out.data_ptr(), 0, 0, 0,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.PPO_ERR_UNSUPPORTED
    # ... while the kernels PATH_AUTO selects take them
    rc = lib.ppo_tvf_targets(tr.data_ptr(), 4, td.data_ptr(), _lib.DONE_U8, 4, tvs.data_ptr(), 4 * len(hz), len(hz), 128, 4,
                             len(hz), len(hz), ctypes.byref(ctab), rt.PATH_AUTO, out.data_ptr(), 0, 0, 0,
                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and lib.ppo_tvf_last_path() & 0xff in (rt.PATH_TSLAB, rt.PATH_QSLAB)
    torch.cuda.synchronize()
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    assert rel_err(out.cpu().numpy(), ref) <= TOL
