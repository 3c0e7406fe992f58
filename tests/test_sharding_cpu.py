"""CPU: the N > 1 host logic -- env sharding and the statistics exchange -- over gloo with world_size 2.

The kernels themselves are per-env independent; what multi-GPU adds is (a) the contiguous env split and
(b) the all-reduce of the statistics record between the GAE scan and the normalisation.  Both are plain
host / torch.distributed code and run here without a GPU.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from ppo_b200 import _lib, rollout, synthetic


def _record(adv, ret):
    s = np.zeros(_lib.STATS_LEN)
    a, r = adv.astype(np.float64), ret.astype(np.float64)
    s[_lib.STAT_COUNT] = a.size
    s[_lib.STAT_ADV_SUM], s[_lib.STAT_ADV_SUMSQ], s[_lib.STAT_ADV_MIN], s[_lib.STAT_ADV_MAX] = a.sum(), (a * a).sum(), a.min(), a.max()
    s[_lib.STAT_RET_SUM], s[_lib.STAT_RET_SUMSQ], s[_lib.STAT_RET_MIN], s[_lib.STAT_RET_MAX] = r.sum(), (r * r).sum(), r.min(), r.max()
    return torch.from_numpy(s)


def test_shard_envs_partitions_the_env_axis():
    x = np.arange(3 * 11 * 2).reshape(3, 11, 2)
    for world in (1, 2, 3, 4, 8, 11):
        parts = [rollout.shard_envs(x, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate(parts, axis=1), x)
        sizes = [p.shape[1] for p in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert rollout.shard_envs(np.zeros([4, 2]), 5, 8).shape == (4, 0)


def test_merge_stats_matches_unsharded_moments():
    r, d, v = synthetic.rollout(40, 37, seed=1)
    adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    ret = adv + v[:-1]
    whole = _record(adv, ret)
    parts = [_record(rollout.shard_envs(adv, k, 3), rollout.shard_envs(ret, k, 3)) for k in range(3)]
    merged = rollout.merge_stats(parts)
    np.testing.assert_allclose(merged.numpy(), whole.numpy(), rtol=1e-13)
    summary = rollout.stats_summary(merged)
    np.testing.assert_allclose(summary["adv_mean"], adv.astype(np.float64).mean(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(summary["adv_std"], adv.astype(np.float64).std(), rtol=1e-10)


def _worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, d, v = synthetic.rollout(48, 50, seed=2)
        adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
        ret = adv + v[:-1]
        mine = _record(rollout.shard_envs(adv, rank, world), rollout.shard_envs(ret, rank, world))
        full = rollout.allreduce_stats(mine.clone())
        only = rollout.allreduce_moments(mine.clone())
        np.save(os.path.join(result_dir, f"full{rank}.npy"), full.numpy())
        np.save(os.path.join(result_dir, f"moments{rank}.npy"), only.numpy())
        np.save(os.path.join(result_dir, f"local{rank}.npy"), mine.numpy())
        # uneven shards (51 envs over 2 ranks = 26 + 25): the row folds divide by the true env count
        from ppo_b200 import normalization
        mine_envs = rollout.shard_envs(np.zeros([1, 51]), rank, world).shape[1]
        counts = [normalization.global_env_count(mine_envs), normalization.global_env_count(mine_envs)]  # 2nd: cached
        np.save(os.path.join(result_dir, f"envs{rank}.npy"), np.asarray([mine_envs] + counts))
    finally:
        dist.destroy_process_group()


def test_allreduce_stats_world_size_2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r, d, v = synthetic.rollout(48, 50, seed=2)
    adv = oracle.gae(r, v[:-1], v[-1], d, 0.999, 0.95)
    whole = _record(adv, adv + v[:-1]).numpy()
    for rank in range(2):
        full = np.load(tmp_path / f"full{rank}.npy")
        np.testing.assert_allclose(full, whole, rtol=1e-13)
        mom = np.load(tmp_path / f"moments{rank}.npy")
        local = np.load(tmp_path / f"local{rank}.npy")
        np.testing.assert_allclose(mom[:3], whole[:3], rtol=1e-13)   # the three moments are global ...
        assert np.array_equal(mom[3:], local[3:])                      # ... everything else stays local
    assert np.array_equal(np.load(tmp_path / "full0.npy"), np.load(tmp_path / "full1.npy"))
    assert np.load(tmp_path / "envs0.npy").tolist() == [26, 51, 51]
    assert np.load(tmp_path / "envs1.npy").tolist() == [25, 51, 51]
