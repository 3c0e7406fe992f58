"""Multi-GPU consistency check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py

Every rank computes its env shard with the moment exchange over the whole group, then the WHOLE rollout on its own
GPU with a single-rank group, and compares the shard of the second with the first: the sharded pass must reproduce
the unsharded one (normalised advantages, return statistics, intrinsic-return and reward normalisation state).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppo_b200 import normalization as nz, rollout, synthetic, tvf  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    solo = [dist.new_group([r]) for r in range(world)][rank]
    steps, envs, heads = 64, 48 * world + 1, 32  # uneven shards: rank 0 takes one env more than the others
    hz = tvf.get_value_head_horizons(heads, 1000)
    r, d, v = synthetic.rollout(steps, envs, seed=5)
    vs = synthetic.value_samples(steps, envs, heads, seed=5)
    kw = dict(horizons=hz, gamma=0.9999, mode="advanced", distribution="exponential", n_step=12, max_samples=6, seed=4)
    failures = []

    def check(name, got, ref, tol=1e-6):
        got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
        err = float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30)) if ref.size else 0.0
        if not err <= tol:
            failures.append(f"rank {rank}: {name}: rel err {err:.3e} > {tol}")

    def shard(x):
        return rollout.shard_envs(x, rank, world)

    local = shard(np.zeros([1, envs])).shape[1]

    # ---- the return stage: GAE / TD(lambda) + globally normalised advantages + TVF targets
    full = rollout.calculate_returns(r, v, d, 0.999, 0.95, 0.9, advantage_clipping=2.0,
                                     tvf=dict(value_samples=vs, **kw), group=solo)
    part = rollout.calculate_returns(shard(r), shard(v), shard(d), 0.999, 0.95, 0.9, advantage_clipping=2.0,
                                     tvf=dict(value_samples=shard(vs), **kw))
    check("advantage", part["advantage"], shard(full["advantage"]), 0.0)
    check("advantages (normalised)", part["advantages"].reshape(steps, -1), shard(full["advantages"].reshape(steps, envs)))
    check("tvf_returns", part["tvf_returns"], shard(full["tvf_returns"]), 0.0)
    a, b = rollout.stats_summary(part["stats"]), rollout.stats_summary(full["stats"])
    for key in b:
        check("stats." + key, a[key], b[key], 1e-9)
    a, b = rollout.stats_summary(part["tvf_stats"]), rollout.stats_summary(full["tvf_stats"])
    for key in ("count", "adv_mean", "adv_std", "adv_min", "adv_max", "adv_sparsity"):
        check("tvf_stats." + key, a[key], b[key], 1e-9)

    # ---- intrinsic-return normalisation and the reward wrapper: running moments over ALL envs of every step
    rng = np.random.RandomState(11)
    ir_r = np.abs(rng.normal(0.2, 0.3, [steps, envs])).astype(np.float32)
    ir_v = rng.normal(0.5, 0.4, [steps + 1, envs]).astype(np.float32)
    whole, mine = nz.IntrinsicReturns(envs), nz.IntrinsicReturns(local)
    for _ in range(2):
        f = whole.calculate_intrinsic_returns(ir_r, ir_v, d, 0.99, 0.95, normalize=True, center=True, group=solo)
        p = mine.calculate_intrinsic_returns(shard(ir_r), shard(ir_v), shard(d), 0.99, 0.95, normalize=True, center=True)
        check("int_advantage", p["int_advantage"], shard(f["int_advantage"]))
        check("int_rewards", p["int_rewards"], shard(f["int_rewards"]))
        check("ir rms", mine.intrinsic_returns_rms.save_state(), whole.intrinsic_returns_rms.save_state(), 1e-11)
    w_all, w_mine = nz.VecNormalizeRewards(envs, gamma=0.99, clip=1.0), nz.VecNormalizeRewards(local, gamma=0.99, clip=1.0)
    f, fc = w_all.normalize_rollout(r * 5, d, group=solo)
    p, pc = w_mine.normalize_rollout(shard(r * 5), shard(d))
    check("scaled rewards", p, shard(f))
    check("ret_rms", w_mine.ret_rms.save_state(), w_all.ret_rms.save_state(), 1e-11)
    total = pc.clone()
    dist.all_reduce(total)
    check("clip count", total.item(), fc.item(), 0.0)

    # ---- device-resident pass: the moment exchange inside the GAE launch over NVLink peer mailboxes must give the
    # same bits as the NCCL all-reduce between two launches, and both the unsharded normalisation
    from ppo_b200 import returns_truncated as rt
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 12, 6, seed=4)
    dev_t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    sr, sv, sd, svs = dev_t(shard(r)), dev_t(shard(v)), dev_t(shard(d)).view(torch.uint8), dev_t(shard(vs))
    passes = {}
    for name, env in (("peers", "1"), ("nccl", "0")):
        os.environ["PPO_PEER_EXCHANGE"] = env
        rp = rollout.ReturnPass(steps, local, hz, 0.999, 0.9999, 0.95, 0.9, advantage_clipping=2.0)
        rp.set_samples(table)
        for _ in range(3):  # several epochs through the same mailboxes
            rp.run(sr, sv, sd, svs)
        torch.cuda.synchronize()
        passes[name] = (rp.advantages.clone(), rp.kernel_launches_per_pass, rp.peers is not None,
                        rp.global_stats.clone(), rp.stats.clone())
        rp.check_peers()
        dist.barrier()
        if rp.peers is not None:
            rp.peers.close()
    adv_p, launches_p, used_p, gstats, _ = passes["peers"]
    adv_n, launches_n, _, _, nstats = passes["nccl"]
    if rank == 0:
        print(f"dist_check: peer exchange {'active' if used_p else 'UNAVAILABLE'}: {launches_p} launches per pass "
              f"against {launches_n} with the all-reduce", flush=True)
    check("peers vs nccl normalised advantages", adv_p.cpu().numpy(), adv_n.cpu().numpy(), 0.0 if used_p else 1e-6)
    check("peers normalised advantages vs unsharded", adv_p.cpu().numpy().reshape(steps, -1),
          shard(np.asarray(full["advantages"]).reshape(steps, envs)))
    if used_p:
        a, b = rollout.stats_summary(gstats), rollout.stats_summary(full["stats"])
        for key in ("count", "adv_mean", "adv_std", "adv_min", "adv_max", "ret_mean", "ret_min", "ret_max"):
            check("peer global stats." + key, a[key], b[key], 1e-9)
        if not launches_p < launches_n:
            failures.append(f"rank {rank}: expected fewer launches with the peer exchange ({launches_p} vs {launches_n})")
    # both exchanges leave the local record in rp.stats and the same merged record in rp.global_stats
    check("global_stats: nccl merge vs unsharded count", rollout.stats_summary(passes["nccl"][3])["count"], steps * envs, 0.0)
    for key in ("adv_mean", "adv_std", "ret_mean", "ret_min", "ret_max"):
        check("nccl global stats." + key, rollout.stats_summary(passes["nccl"][3])[key], rollout.stats_summary(full["stats"])[key], 1e-9)
    check("local stats count", rollout.stats_summary(nstats)["count"], steps * local, 0.0)

    flags = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flags)
    for line in failures:
        print(line, flush=True)
    if rank == 0:
        print("dist_check: " + ("OK" if flags.item() == 0 else f"{int(flags.item())} FAILURES") + f" on {world} GPUs", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flags.item() == 0 else 1)


if __name__ == "__main__":
    main()
