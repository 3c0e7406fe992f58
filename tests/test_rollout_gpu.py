"""GPU parity of the return stage (ppo_b200.rollout): fused GAE + TD(lambda) + statistics, advantage
normalisation (rl/rollout.py:1887-1897) and the device-resident ReturnPass against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import returns_truncated as rt
from ppo_b200 import rollout, synthetic
from util import TOL, rel_err

pytestmark = pytest.mark.gpu


def test_advantage_normalisation_golden(golden):
    adv = golden["gae_bool"]
    assert rel_err(rollout.normalize_advantages(adv), golden["advnorm"]) <= TOL
    got = rollout.normalize_advantages(adv, epsilon=1e-8, clipping=1.5)
    assert rel_err(got, golden["advnorm_clip"]) <= TOL
    assert got.min() >= -1.5 and got.max() <= 1.5


@pytest.mark.parametrize("steps,envs", [(256, 4096), (128, 128), (17, 5), (1, 1)])
def test_normalise_vs_oracle(steps, envs):
    r, d, v = synthetic.rollout(steps, envs, seed=envs)
    adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    ref = oracle.normalize_advantages(adv)
    got = rollout.normalize_advantages(adv)
    assert got.shape == ref.shape and got.dtype == np.float32
    if steps * envs > 1:
        assert rel_err(got, ref) <= TOL
    clipped = rollout.normalize_advantages(adv, clipping=0.5)
    assert rel_err(clipped, oracle.normalize_advantages(adv, clipping=0.5)) <= TOL


def test_calculate_returns_matches_oracle():
    steps, envs = 96, 160
    hz = oracle.get_value_head_horizons(64, 3000)
    r, d, v = synthetic.rollout(steps, envs, seed=21)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=21)
    res = rollout.calculate_returns(r, v, d, 0.999, 0.95, 0.9, advantage_clipping=2.0,
                                    tvf=dict(value_samples=vs, horizons=hz, gamma=0.9999, mode="advanced",
                                             distribution="exponential", n_step=12, max_samples=6, seed=4))
    adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    assert rel_err(res["advantage"], adv) <= TOL
    assert rel_err(res["ext_returns"], oracle.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.9)) <= TOL
    assert rel_err(res["advantages"], oracle.normalize_advantages(adv, clipping=2.0)) <= TOL
    ref = oracle.get_return_estimate("exponential", "advanced", 0.9999, r, d, hz, hz, vs, n_step=12, max_samples=6, seed=4)
    assert rel_err(res["tvf_returns"], ref) <= TOL
    summary = rollout.stats_summary(res["stats"])
    assert summary["count"] == adv.size
    np.testing.assert_allclose(summary["adv_mean"], adv.astype(np.float64).mean(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(summary["adv_std"], adv.astype(np.float64).std(), rtol=1e-9)
    np.testing.assert_allclose(summary["value_mean"], v.astype(np.float64).mean(), rtol=1e-9)
    np.testing.assert_allclose(summary["value_std"], v.astype(np.float64).std(), rtol=1e-9)
    assert summary["adv_sparsity"] == np.mean(adv == 0) and summary["ret_sparsity"] == np.mean(res["ext_returns"] == 0)
    t = rollout.stats_summary(res["tvf_stats"])
    last = np.asarray(res["tvf_returns"])[:, :, -1].astype(np.float64)
    assert t["adv_min"] == last.min() and t["adv_max"] == last.max()


def test_return_pass_resident_and_pinned_out():
    steps, envs = 64, 200
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=33)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=33)
    v = vs[:, :, -1]  # the trainer's value estimate is the longest horizon of the TVF buffer (rl/tvf.py:325-340)
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    dr, dd, dvs = torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda().view(torch.uint8), torch.from_numpy(vs).cuda()
    rp = rollout.ReturnPass(steps, envs, hz, 0.999, 0.99997, 0.95, 0.9)
    rp.set_samples(table)
    for _ in range(2):  # twice: tickets and workspaces must be left reusable
        rp.run(dr, dvs[:, :, -1], dd, dvs)
    torch.cuda.synchronize()
    adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    assert rel_err(rp.advantage.cpu().numpy(), adv) <= TOL
    assert rel_err(rp.ext_returns.cpu().numpy(), oracle.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.9)) <= TOL
    assert rel_err(rp.advantages.cpu().numpy(), oracle.normalize_advantages(adv)) <= TOL
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    assert rel_err(rp.tvf_returns.cpu().numpy(), ref) <= TOL
    # the host path with pinned buffers delivers the same bits as the resident pass
    out = {"advantage": torch.empty((steps, envs)).pin_memory(), "tvf_returns": torch.empty((steps, envs, len(hz))).pin_memory()}
    res = rollout.calculate_returns(torch.from_numpy(r).pin_memory(), torch.from_numpy(np.ascontiguousarray(v)).pin_memory(),
                                    torch.from_numpy(d).pin_memory(), 0.999, 0.95, 0.9,
                                    tvf=dict(value_samples=torch.from_numpy(vs).pin_memory(), horizons=hz, gamma=0.99997,
                                             mode="advanced", distribution="exponential", n_step=20, max_samples=8, seed=1),
                                    out=out)
    assert res["tvf_returns"] is out["tvf_returns"]
    assert torch.equal(out["tvf_returns"], rp.tvf_returns.cpu())
    assert torch.equal(out["advantage"], rp.advantage.cpu())


def test_return_pass_rebuilds_lean_tables_per_pass():
    """A pass that redraws its sample table every rollout: after the first pass on the transposed slab kernel only the
    slab section of the tables is rebuilt and uploaded; the targets must be those of a full build, a table the slab
    kernel cannot serve (fixed n-step: the ring kernel's case) must still run, and the ring of blobs must survive
    more rebuilds than it has slots."""
    steps, envs = 128, 96
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, _ = synthetic.rollout(steps, envs, seed=5)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=5)
    dr, dd, dvs = torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda().view(torch.uint8), torch.from_numpy(vs).cuda()
    rp = rollout.ReturnPass(steps, envs, hz, 0.999, 0.99997, 0.95, 0.9)
    for i in range(5):
        table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=10 + i)
        tabs = rp.set_samples(table)
        assert bool(tabs.layout.sections & rt.TABLES_GENERIC) == (i == 0)
        rp.run(dr, None, dd, dvs)
        assert rp.tvf_kernel in ("tslab", "qslab")  # (qslab at 128 / 256 steps x 128 horizons)
        fresh = rollout.ReturnPass(steps, envs, hz, 0.999, 0.99997, 0.95, 0.9)
        fresh.set_samples(table)
        fresh.run(dr, None, dd, dvs)
        torch.cuda.synchronize()
        assert torch.equal(rp.tvf_returns, fresh.tvf_returns) and torch.equal(rp.advantages, fresh.advantages)
        assert torch.equal(rp.tvf_stats, fresh.tvf_stats)
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    assert rel_err(rp.tvf_returns.cpu().numpy(), ref) <= TOL
    fixed = rt.draw_n_step_samples("fixed", "advanced", hz, steps, 20, 8)
    tabs = rp.set_samples(fixed)
    assert tabs.layout.sections & rt.TABLES_GENERIC  # the builder kept the generic tables for this one
    rp.run(dr, None, dd, dvs)
    torch.cuda.synchronize()
    assert rp.tvf_kernel not in ("tslab", "qslab")
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=fixed)
    assert rel_err(rp.tvf_returns.cpu().numpy(), ref) <= TOL


def test_return_pass_combines_intrinsic_advantage():
    """rl/rollout.py:1225-1231: advantage = ext_advantage + ir.scale * int_advantage, normalised as a whole."""
    steps, envs = 96, 160
    hz = oracle.get_value_head_horizons(128, 30000)
    r, d, v = synthetic.rollout(steps, envs, seed=21)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=21)
    rng = np.random.RandomState(3)
    int_adv = rng.normal(0.05, 0.3, [steps, envs]).astype(np.float32)
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    rp = rollout.ReturnPass(steps, envs, hz, 0.999, 0.99997, 0.95, 0.9, advantage_clipping=3.0)
    rp.set_samples(table)
    for _ in range(2):
        rp.run(cu(r), cu(v), cu(d).view(torch.uint8), cu(vs), int_advantage=cu(int_adv), ir_scale=0.3)
    torch.cuda.synchronize()
    ext = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    comb = ext + np.float32(0.3) * int_adv
    assert np.array_equal(rp.ext_advantage.cpu().numpy(), ext)
    assert np.array_equal(rp.advantage.cpu().numpy(), comb)
    want = np.clip(oracle.normalize_advantages(comb), -3.0, 3.0)
    assert rel_err(rp.advantages.cpu().numpy(), want.reshape(-1)) <= TOL
    assert rel_err(rp.ext_returns.cpu().numpy(), oracle.td_lambda(r, v[:-1], v[-1], d, 0.999, 0.9)) <= TOL
    ref = oracle.sampled_return_multi(0.99997, r, d, hz, hz, vs, n_step_samples=table)
    assert rel_err(rp.tvf_returns.cpu().numpy(), ref) <= TOL
    s = rollout.stats_summary(rp.combined_stats)
    np.testing.assert_allclose(s["adv_mean"], comb.astype(np.float64).mean(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(s["ret_std"], (np.float32(0.3) * int_adv).astype(np.float64).std(), rtol=1e-9)
    want_ratio = ((ext.astype(np.float64) ** 2).mean() / ((np.float32(0.3) * int_adv).astype(np.float64) ** 2).mean()) ** 0.5
    np.testing.assert_allclose(rp.adv_ratio(), want_ratio, rtol=1e-9)
    # the plain pass on the same object still normalises the extrinsic advantage alone
    rp.run(cu(r), cu(v), cu(d).view(torch.uint8), cu(vs))
    torch.cuda.synchronize()
    assert rel_err(rp.advantages.cpu().numpy(), np.clip(oracle.normalize_advantages(ext), -3.0, 3.0).reshape(-1)) <= TOL


@pytest.mark.parametrize("method,mode", [("off", "average"), ("timelimit", "interpolate"), ("timelimit", "average"),
                                         ("timelimit", "substitute")])
def test_rollout_buffers_written_on_the_device(method, mode):
    """The per-step writes of Runner.generate_rollout (rl/rollout.py:790-813, :878-892) with the model outputs left on
    the device, against the same writes in NumPy through the oracle's trim_horizons; then one return pass straight off
    the buffers against the oracle on the NumPy buffers."""
    import types
    from ppo_b200 import tvf as tvf_mod
    steps, envs, timeout = 32, 40, 300
    hz = oracle.get_value_head_horizons(128, 30000)
    log = types.SimpleNamespace(watch_stats=lambda *a, **k: None, watch_mean=lambda *a, **k: None)
    runner = types.SimpleNamespace(ext_rewards=np.zeros([steps, envs], np.float32), terminals=None, all_obs=None,
                                   tvf_horizons=list(hz), value_heads=["ext"], N=steps, log=log, episode_length_buffer=[])
    args = types.SimpleNamespace(env=types.SimpleNamespace(timeout=timeout),
                                 tvf=types.SimpleNamespace(eta_percentile=90, eta_buffer=32, eta_minh=128, trim_clip=-1))
    mod = tvf_mod.TVFRunnerModule(runner, args)
    bufs = rollout.RolloutBuffers(steps, envs, hz, mod, trimming=method, trimming_mode=mode)
    rng = np.random.RandomState(4)
    r, d, _ = synthetic.rollout(steps, envs, seed=4)
    time = rng.randint(0, timeout - 20, size=envs)
    want_v = np.zeros([steps + 1, envs, len(hz), 1], np.float32)
    want_u, want_f = np.zeros_like(want_v), np.zeros([steps + 1, envs], np.float32)
    for t in range(steps + 1):
        out = rng.normal(0.3, 0.5, [envs, len(hz), 1]).astype(np.float32)
        ref_in = out.copy()
        ref_in[:, 0, :] = 0
        want_u[t] = ref_in
        new, fin, _ = oracle.trim_horizons(hz, ref_in, time, method, mode, timeout=timeout)
        want_v[t], want_f[t] = new, fin
        model_out = {"tvf_value": torch.from_numpy(out).cuda(), "value": torch.from_numpy(out[:, -1, :]).cuda()}
        if t < steps:
            bufs.write_step(t, model_out, time, r[t], d[t])
            time = np.where(d[t], 0, time + 1)
        else:
            bufs.write_final(model_out, time)
    torch.cuda.synchronize()
    assert np.array_equal(bufs.tvf_untrimmed_value.cpu().numpy(), want_u)
    assert rel_err(bufs.tvf_value.cpu().numpy(), want_v) <= 1e-6
    assert rel_err(bufs.tvf_final_value.cpu().numpy(), want_f) <= 1e-6 or method == "off"
    assert np.array_equal(bufs.ext_rewards.cpu().numpy(), r) and np.array_equal(bufs.terminals.cpu().numpy() != 0, d)
    assert bufs.value.shape == (steps + 1, envs, 1)
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, 8, seed=1)
    rp = rollout.ReturnPass(steps, envs, hz, 0.999, 0.999, 0.95, 0.9)
    rp.set_samples(table)
    bufs.returns(rp)
    torch.cuda.synchronize()
    vs = bufs.tvf_value.cpu().numpy()[..., 0]
    assert rel_err(rp.tvf_returns.cpu().numpy(), oracle.sampled_return_multi(0.999, r, d, hz, hz, vs, n_step_samples=table)) <= TOL
    assert rel_err(rp.advantage.cpu().numpy(), oracle.gae(r, vs[:-1, :, -1], vs[-1, :, -1], d, 0.999, 0.95)) <= TOL


def test_env_sharding_is_exact_on_one_gpu():
    """Two env shards processed separately + summed moments == the unsharded pass (the N > 1 data path without NCCL)."""
    steps, envs = 80, 130
    r, d, v = synthetic.rollout(steps, envs, seed=8)
    full = rollout.calculate_returns(r, v, d, 0.999, 0.95, 0.95)
    parts, stats = [], []
    for rank in range(2):
        rs, vs_, ds = (rollout.shard_envs(x, rank, 2) for x in (r, v, d))
        res = rollout.calculate_returns(np.ascontiguousarray(rs), np.ascontiguousarray(vs_), np.ascontiguousarray(ds),
                                        0.999, 0.95, 0.95)
        parts.append(res)
        stats.append(res["local_stats"])
    merged = rollout.merge_stats(stats)
    assert np.array_equal(np.concatenate([p["advantage"] for p in parts], axis=1), full["advantage"])
    s_full, s_merged = full["stats"].cpu().numpy(), merged.cpu().numpy()
    assert s_full[0] == s_merged[0] and s_full[3] == s_merged[3] and s_full[4] == s_merged[4]
    np.testing.assert_allclose(s_merged[1:3], s_full[1:3], rtol=1e-12)
    norm = [rollout.normalize_advantages(p["advantage"], stats=merged) for p in parts]
    got = np.concatenate([n.reshape(steps, -1) for n in norm], axis=1).reshape(-1)
    assert rel_err(got, full["advantages"]) <= 1e-6


@pytest.mark.parametrize("steps,envs,heads,block", [(96, 700, 64, 148), (64, 90, 32, 40), (40, 10, 32, None)])
def test_host_blocks_pipeline_equals_one_launch(steps, envs, heads, block):
    """Pinned host buffers: env blocks uploaded / computed / downloaded on three streams give exactly the targets
    and statistics of one whole-buffer launch (ragged last block included)."""
    hz = oracle.get_value_head_horizons(heads, 3000)
    r, d, _ = synthetic.rollout(steps, envs, seed=5)
    vs = torch.from_numpy(synthetic.value_samples(steps, envs, len(hz), seed=6)).pin_memory()
    table = oracle.draw_n_step_samples("exponential", "advanced", hz, steps, 12, 6, seed=3)
    out = torch.empty((steps, envs, len(hz)), dtype=torch.float32).pin_memory()
    assert rt.host_pipeline_ok(vs, out)
    stats = rt._launch_host_blocks(0.9999, r, d, hz, hz, vs, table, False, out, block_envs=block)
    torch.cuda.current_stream().synchronize()
    whole, wstats = rt._launch(0.9999, r, d, hz, hz, vs.numpy(), table, False, with_stats=True)
    assert np.array_equal(out.numpy(), whole)
    s, w = stats.cpu().numpy(), wstats.cpu().numpy()
    assert s[0] == w[0] and s[3] == w[3] and s[4] == w[4]
    np.testing.assert_allclose(s[1:3], w[1:3], rtol=1e-12)
    ref = oracle.sampled_return_multi(0.9999, r, d, hz, hz, vs.numpy(), n_step_samples=table)
    assert rel_err(out.numpy(), ref) <= TOL
    # a second call re-uses the cached scratch while the first results are still being read: must stay correct
    out2 = torch.empty_like(out).pin_memory()
    rt._launch_host_blocks(0.9999, r, d, hz, hz, vs, table, False, out2, block_envs=block, with_stats=False)
    torch.cuda.current_stream().synchronize()
    assert torch.equal(out, out2)


def test_registered_numpy_buffers_take_the_pipeline():
    """ppo_b200.pin_numpy: a persistent NumPy value buffer is page-locked in place once; the drop-in call then runs
    the env-blocked pipeline and still returns a plain NumPy array with the same bits."""
    import ppo_b200
    from ppo_b200 import _device as dev
    steps, envs = 64, 400
    hz = oracle.get_value_head_horizons(64, 3000)
    r, d, _ = synthetic.rollout(steps, envs, seed=9)
    vs = synthetic.value_samples(steps, envs, len(hz), seed=9)
    kw = dict(n_step=12, max_samples=6, seed=3)
    plain = rt.get_return_estimate("exponential", "advanced", 0.9999, r, d, hz, hz, vs, **kw)
    assert dev.pinned_tensor_of(vs) is None
    assert ppo_b200.pin_numpy(vs) is vs and dev.pinned_tensor_of(vs) is not None
    piped = rt.get_return_estimate("exponential", "advanced", 0.9999, r, d, hz, hz, vs, **kw)
    assert isinstance(piped, np.ndarray) and piped.dtype == np.float32 and np.array_equal(piped, plain)
    view = vs[:, :200]  # a view that is not env-contiguous per row falls back to the plain path, same result
    assert np.array_equal(rt.get_return_estimate("exponential", "advanced", 0.9999, r[:, :200], d[:, :200], hz, hz,
                                                 view, **kw), plain[:, :200])
