"""The return / advantage stage of ``Runner.calculate_returns`` and the batch advantage normalisation
of ``Runner.train_policy`` (reference: ``rl/rollout.py:1182-1285`` and ``:1886-1900``), as free
functions over rollout buffers, with the env axis optionally sharded over ``torch.distributed`` ranks.

Envs are independent, so each rank owns a contiguous block of envs of every ``[N, A, ...]`` buffer and
runs the same kernels on its block.  The only exchange is the sum of ``{count, sum, sum of squares}``
of the raw advantages (three float64, plus min / max for the logged statistics) before the
normalisation: inside the GAE launch over NVLink peer mailboxes for the resident pass (``ReturnPass``,
``PeerMailbox``), one NCCL all-reduce otherwise.
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _device as dev
from . import _lib
from . import returns as _returns
from . import returns_truncated as _rt


def shard_envs(array, rank, world, axis=1):
    """Contiguous env block of ``array`` owned by ``rank`` (even split, remainder to the low ranks)."""
    n = array.shape[axis]
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    index = [slice(None)] * array.ndim
    index[axis] = slice(lo, hi)
    return array[tuple(index)]


def gather_rows(buffer, sample, check=True):
    """``buffer[sample]`` for a device-resident ``[M, ...]`` buffer and an int64 index vector: the minibatch selection
    of ``rl/rollout.py:2349-2372`` without the host round trip.  Rows are gathered as bytes (any dtype, any trailing
    shape; the row axis may be strided).  ``check`` raises ``IndexError`` for an out-of-range index like NumPy (one
    synchronisation); with ``check=False`` such rows are zero-filled."""
    device = dev.require_cuda()
    lib = _lib.load()
    if not (isinstance(buffer, torch.Tensor) and buffer.is_cuda):
        raise ValueError("gather_rows: buffer must be a CUDA tensor (device-resident rollout buffer)")
    rows = buffer.shape[0]
    if buffer.dim() > 1 and not buffer[0].is_contiguous():
        buffer = buffer.contiguous()
    index = dev.to_device(sample, torch.int64, device).reshape(-1).contiguous()
    row_bytes = (buffer[0].numel() if rows else 0) * buffer.element_size()
    out = torch.empty((index.numel(),) + tuple(buffer.shape[1:]), dtype=buffer.dtype, device=device)
    bad = torch.zeros(1, dtype=torch.int32, device=device)
    pitch = buffer.stride(0) * buffer.element_size() if rows > 1 else row_bytes
    rc = lib.ppo_gather_rows(dev.ptr(buffer), pitch, row_bytes, rows, dev.ptr(index), index.numel(), dev.ptr(out),
                             row_bytes, dev.ptr(bad), dev.stream_ptr())
    _lib.check(rc, "ppo_gather_rows")
    if check and int(bad.item()):
        raise IndexError(f"index out of bounds for axis 0 with size {rows}")
    return out


_SUM_SLOTS = [_lib.STAT_COUNT, _lib.STAT_ADV_SUM, _lib.STAT_ADV_SUMSQ, _lib.STAT_RET_SUM, _lib.STAT_RET_SUMSQ,
              _lib.STAT_VAL_SUM, _lib.STAT_VAL_SUMSQ, _lib.STAT_VAL_COUNT, _lib.STAT_ADV_ZEROS, _lib.STAT_RET_ZEROS]
_MIN_SLOTS = [_lib.STAT_ADV_MIN, _lib.STAT_RET_MIN]
_MAX_SLOTS = [_lib.STAT_ADV_MAX, _lib.STAT_RET_MAX]


def merge_stats(records):
    """Combine statistics records of disjoint env blocks: sums are added, min / max are reduced (``ppo_merge_stats``;
    CPU tensors -- the gloo tests of the sharding logic -- take the same reduction in torch)."""
    stack = records if isinstance(records, torch.Tensor) and records.dim() == 2 else torch.stack(list(records))
    stack = stack.contiguous()
    merged = torch.zeros_like(stack[0])
    if stack.is_cuda:
        rc = _lib.load().ppo_merge_stats(dev.ptr(stack), stack.shape[0], dev.ptr(merged), dev.stream_ptr())
        _lib.check(rc, "ppo_merge_stats")
        return merged
    merged[_SUM_SLOTS] = stack[:, _SUM_SLOTS].sum(dim=0)
    merged[_MIN_SLOTS] = stack[:, _MIN_SLOTS].min(dim=0).values
    merged[_MAX_SLOTS] = stack[:, _MAX_SLOTS].max(dim=0).values
    return merged


def allreduce_stats(stats, group=None):
    """Combine per-rank statistics records in place across ``group`` (same result as :func:`merge_stats`): one
    all-gather of the 16-double records, then the merge on every rank in rank order (identical bits everywhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    world = dist.get_world_size(group)
    gathered = torch.empty((world, stats.numel()), dtype=stats.dtype, device=stats.device)
    dist.all_gather(list(gathered.unbind(0)), stats.contiguous(), group=group)  # rows of one buffer (NCCL and gloo)
    stats.copy_(merge_stats(gathered))
    return stats


def allreduce_moments(stats, group=None):
    """The one exchange the normalisation needs: sum of {count, sum, sum of squares} (3 x float64)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    dist.all_reduce(stats[_lib.STAT_COUNT:_lib.STAT_ADV_SUMSQ + 1], op=dist.ReduceOp.SUM, group=group)
    return stats


def moments(x):
    """Device statistics record (count / sum / sumsq / min / max in the ADV slots) of a float32 array."""
    device = dev.require_cuda()
    lib = _lib.load()
    t = dev.to_device(x, torch.float32, device).contiguous()
    stats = torch.empty(_lib.STATS_LEN, dtype=torch.float64, device=device)
    ws = dev.workspace(lib.ppo_gae_workspace_bytes(0), device)
    rc = lib.ppo_moments(dev.ptr(t), t.numel(), dev.ptr(stats), dev.ptr(ws), ws.numel(), dev.stream_ptr())
    _lib.check(rc, "ppo_moments")
    return stats


def normalize_advantages(advantage, epsilon=1e-8, clipping=None, stats=None, group=None, out=None):
    """``(adv - mean) / (std + eps)`` over the whole (global) batch, optional clip; returns the flattened
    float32 vector like ``rl/rollout.py:1887-1897``.  ``stats`` is the record returned by
    ``returns.gae_td_lambda(..., with_stats=True)`` (local to this rank); it is summed over ``group``."""
    device = dev.require_cuda()
    lib = _lib.load()
    np_like = dev.is_numpy_like(advantage)
    adv = dev.to_device(advantage, torch.float32, device).contiguous()
    if stats is None:
        stats = moments(adv)
    else:
        stats = stats.clone()
    allreduce_moments(stats, group)
    if out is None:
        out = torch.empty(adv.numel(), dtype=torch.float32, device=device)
    rc = lib.ppo_adv_normalize(dev.ptr(adv), adv.numel(), dev.ptr(stats), float(epsilon),
                               float(clipping) if clipping is not None else 0.0, dev.ptr(out), dev.stream_ptr())
    _lib.check(rc, "ppo_adv_normalize")
    return dev.like_input(out, np_like)


def calculate_returns(ext_rewards, ext_value_estimates, terminals, gamma, lambda_policy=0.95, lambda_value=0.95,
                      advantage_epsilon=1e-8, advantage_clipping=None, tvf=None, group=None, out=None):
    """One pass of the return stage (``rl/rollout.py:1182-1285`` + ``:1886-1897``) over this rank's envs.

    ``ext_value_estimates`` is ``[N+1, A]`` (row N is the final estimate, ``rl/rollout.py:1207-1214``); ``None`` takes
    it off the value curves in ``tvf`` like ``rl/rollout.py:1196`` does (``get_tvf_ext_value_estimate``, rediscounted
    when ``gamma`` differs from the TVF gamma), fused into the target launch.
    ``tvf`` (optional) is a dict with ``value_samples [N+1, A, K]``, ``horizons [K]``, ``gamma`` and the
    ``get_return_estimate`` keywords (``mode``, ``distribution``, ``n_step``, ``max_samples``,
    ``use_log_interpolation``, ``seed``), mirroring ``TVFRunnerModule.calculate_tvf_returns``.
    Returns a dict: ``advantage`` (raw), ``ext_returns``, ``advantages`` (normalised, flattened),
    ``stats`` (global record on the device) and, with ``tvf``, ``tvf_returns`` and ``tvf_stats``.
    Host inputs may be NumPy arrays (NumPy out) or CPU tensors, ideally pinned; ``out`` (optional) maps result
    names to pre-allocated pinned CPU tensors that receive the results by asynchronous copies, and the call
    returns after one stream synchronisation -- the zero-allocation host path for a trainer whose rollout
    buffers live in pinned memory.
    """
    np_like = dev.is_numpy_like(ext_rewards)
    device = dev.require_cuda()
    steps = ext_rewards.shape[0]
    tvf_join = tvf_launch = None
    from_curves = ext_value_estimates is None
    assert not from_curves or tvf is not None, "ext_value_estimates=None needs the value curves in `tvf`"
    if tvf is not None:
        kw = dict(tvf)
        vs = kw.pop("value_samples")
        hz = kw.pop("horizons")
        g = kw.pop("gamma")
        mode = kw.pop("mode", "advanced")
        distribution = kw.pop("distribution", "exponential")
        samples = _rt.draw_n_step_samples(distribution, mode, np.asarray(hz), steps, kw.pop("n_step", 40),
                                          kw.pop("max_samples", 40), kw.pop("seed", None))
        if distribution != "fixed" and mode == "full":
            raise ValueError("calculate_returns: use returns_truncated.get_return_estimate for mode='full'")
        use_log = kw.pop("use_log_interpolation", False)
        to_host = out is not None and "tvf_returns" in out
        vkw = {}
        if from_curves:
            # rl/rollout.py:1196: the value estimate GAE reads comes off the value curves (rl/tvf.py:303-360): the last
            # horizon for the TVF gamma, the rediscounted estimate otherwise -- produced by the target launch itself
            from . import tvf as _tvf
            values = torch.empty((steps + 1, ext_rewards.shape[1]), dtype=torch.float32, device=device)
            weights = None if abs(gamma - g) < 1e-8 else _tvf.rediscount_weights(g, gamma, [int(h) for h in hz])
            vkw = dict(value_out=values, value_weights=weights)
        if to_host and steps > 0 and _rt.host_pipeline_ok(vs, out["tvf_returns"]):
            # rollout buffers in pinned host memory: env blocks are uploaded, computed and downloaded concurrently on
            # side streams, started first so that the transfers overlap the GAE part below
            tvf_join = _rt._launch_host_blocks(g, ext_rewards, terminals, hz, hz, vs, samples, use_log,
                                               out["tvf_returns"], defer_join=True, **vkw)
            if from_curves:  # GAE needs every block's value estimate (not the downloads: those are joined at the end)
                tstats_early, join_all = tvf_join(kernels_only=True), tvf_join

                def tvf_join():
                    join_all()
                    return tstats_early
        elif from_curves:
            early = _rt._launch(g, ext_rewards, terminals, hz, hz, vs, samples, use_log, with_stats=True,
                                keep_on_device=to_host, **vkw)
            tvf_launch = lambda: early  # noqa: E731
        else:
            tvf_launch = lambda: _rt._launch(g, ext_rewards, terminals, hz, hz, vs, samples, use_log,  # noqa: E731
                                             with_stats=True, keep_on_device=to_host)
    if not from_curves:
        values = dev.to_device(ext_value_estimates, torch.float32, device)
    lib = _lib.load()
    clip = float(advantage_clipping) if advantage_clipping is not None else 0.0
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if sharded:
        adv, ret, stats = _returns._gae_launch(ext_rewards, values[:steps], values[steps], terminals, gamma,
                                               lambda_policy, lambda_value, True, True, True)
        local = stats.clone()
        allreduce_stats(stats, group)
        norm = torch.empty(adv.numel(), dtype=torch.float32, device=device)
        rc = lib.ppo_adv_normalize(dev.ptr(adv), adv.numel(), dev.ptr(stats), float(advantage_epsilon), clip,
                                   dev.ptr(norm), dev.stream_ptr())
        _lib.check(rc, "ppo_adv_normalize")
    else:
        # one GPU: scan, statistics and normalisation in one launch (ppo_gae_normalized)
        adv, ret, norm, stats = _returns._gae_launch(ext_rewards, values[:steps], values[steps], terminals, gamma,
                                                     lambda_policy, lambda_value, True, True, True,
                                                     normalize=(float(advantage_epsilon), clip))
        local = stats.clone()

    def deliver(name, tensor):
        if out is not None and name in out:
            out[name].copy_(tensor.view(out[name].shape), non_blocking=True)
            return out[name]
        return dev.like_input(tensor, np_like)

    result = {
        "value_estimates": values,
        "advantage": deliver("advantage", adv),
        "ext_returns": deliver("ext_returns", ret),
        "advantages": deliver("advantages", norm),
        "stats": stats,
        "local_stats": local,
    }
    if tvf is not None:
        if tvf_join is not None:
            tstats = tvf_join()
            result["tvf_returns"] = out["tvf_returns"]
        else:
            targets, tstats = tvf_launch()
            result["tvf_returns"] = deliver("tvf_returns", targets) if to_host else targets
        allreduce_stats(tstats, group)
        result["tvf_stats"] = tstats
    if out is not None:
        torch.cuda.current_stream().synchronize()
    return result


class PeerMailbox:
    """Statistics mailboxes of every rank of ``group`` mapped into this process over NVLink peer memory (CUDA IPC,
    ``ppo_peer_alloc`` / ``ppo_peer_open``), for ``ppo_gae_normalized_peers``: rank r owns ``[2, world, 16]`` float64
    records followed by ``2 x world`` 64-bit flag words (two halves used alternately by epoch parity);
    ``mail_ptrs[r]`` / ``flag_ptrs[r]`` are rank r's buffers as seen from here.  Construction is collective.  ``ok``
    is False when some rank cannot map a peer: callers then keep the all-reduce (and the mappings are released).

    A rank whose peers do not launch within ``timeout_ms`` (env ``PPO_PEER_TIMEOUT_MS``, default 4000: ranks that
    step their envs at different speeds should meet at a host barrier before the pass) gives up instead of hanging the
    GPU: the kernel stores the epoch into ``status`` (mapped pinned host memory) and writes NaN advantages;
    :meth:`check` raises ``TimeoutError`` for it."""

    def __init__(self, device, group=None):
        lib = _lib.load()
        self.lib = lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.global_stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=device)
        self.epoch = 0
        self.timeout_ms = float(os.environ.get("PPO_PEER_TIMEOUT_MS", "4000"))
        self.status = torch.zeros(1, dtype=torch.int64, pin_memory=True)  # written by the kernel on a timed-out wait
        mail_bytes = 2 * self.world * _lib.STATS_LEN * 8
        self._own, self._opened = ctypes.c_void_p(), []
        handle = (ctypes.c_ubyte * 64)()
        ok = self.world <= _lib.MAX_PEERS
        with torch.cuda.device(device):
            ok = ok and lib.ppo_peer_alloc(mail_bytes + 2 * self.world * 8, ctypes.byref(self._own), handle) == 0
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
            bases = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    bases.append(self._own.value or 0)
                    continue
                ptr = ctypes.c_void_p()
                if h is None or lib.ppo_peer_open((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(ptr)) != 0:
                    ok = False
                    bases.append(0)
                else:
                    self._opened.append(ptr)
                    bases.append(ptr.value)
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(flag.item())
        self.mail_ptrs = (ctypes.c_void_p * self.world)(*bases)
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*[b + mail_bytes if b else 0 for b in bases])
        if not self.ok:
            self.close()

    def next_epoch(self):
        self.epoch += 1
        return self.epoch

    def check(self):
        """Raise ``TimeoutError`` if a launch gave up waiting for a peer.  Reads a host word: call it once the stream
        has drained (``ReturnPass.run`` also looks at it on entry, which reports the failure of an earlier pass)."""
        epoch = int(self.status[0])
        if epoch:
            raise TimeoutError(
                f"ppo_gae_normalized_peers: rank {self.rank} waited more than {self.timeout_ms:.0f} ms for the statistics "
                f"record of a peer in pass {epoch} (of {self.epoch}); its normalised advantages are NaN.  Some rank did "
                f"not launch the pass: meet at a barrier before ReturnPass.run(), raise PPO_PEER_TIMEOUT_MS, or set "
                f"PPO_PEER_EXCHANGE=0 for the NCCL all-reduce")

    def close(self):
        """Unmap the peers and free the own mailbox (collective in spirit: no rank may still be launching)."""
        for ptr in self._opened:
            self.lib.ppo_peer_close(ptr)
        self._opened = []
        if self._own:
            self.lib.ppo_peer_free(self._own)
            self._own = ctypes.c_void_p()


class ReturnPass:
    """Device-resident form of the return stage for one rank's env block: buffers stay in HBM, outputs
    and workspaces are allocated once, and ``run()`` is exactly the launches of one pass
    (``rl/rollout.py:1207-1236`` + ``:1891``):

      1. ``ppo_gae``            advantages (lambda_policy) + TD(lambda) returns (lambda_value) + statistics
      2. all-reduce (sum) of {count, sum, sum of squares} of the raw advantages -- only when sharded
      3. ``ppo_adv_normalize``  normalised (optionally clipped) advantages
      4. ``ppo_tvf_targets_ex`` [N, A, K] truncated-return targets + statistics of the longest horizon
         (first, and also producing the value estimate GAE reads, when ``run(values=None)``)

    ``set_samples`` resolves an n-step sample table (host integers only) and uploads it; the reference
    draws a new table every rollout (``rl/returns_truncated.py:85-129``), so call it per pass when the
    table changes.  All launches go to the current CUDA stream; nothing synchronises.
    """

    def __init__(self, steps, envs, horizons, gamma, tvf_gamma, lambda_policy=0.95, lambda_value=0.95,
                 advantage_epsilon=1e-8, advantage_clipping=None, use_log_interpolation=False, group=None,
                 tvf_path=_rt.PATH_AUTO, device=None):
        self.device = device or dev.require_cuda()
        self.lib = _lib.load()
        self.steps, self.envs = int(steps), int(envs)
        self.horizons = np.asarray(horizons)
        self.n_h = len(self.horizons)
        self.gamma, self.tvf_gamma = float(gamma), float(tvf_gamma)
        self.lambda_policy, self.lambda_value = float(lambda_policy), float(lambda_value)
        self.eps = float(advantage_epsilon)
        self.clip = float(advantage_clipping) if advantage_clipping is not None else 0.0
        self.use_log = bool(use_log_interpolation)
        self.group = group
        self.tvf_path = tvf_path
        self.sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        f32 = dict(dtype=torch.float32, device=self.device)
        self.advantage = torch.empty((self.steps, self.envs), **f32)
        self.ext_returns = torch.empty((self.steps, self.envs), **f32)
        self.advantages = torch.empty(self.steps * self.envs, **f32)
        self.tvf_returns = torch.empty((self.steps, self.envs, self.n_h), **f32)
        self.stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=self.device)
        self.tvf_stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=self.device)
        # `stats` is always THIS rank's record; `global_stats` the record merged over the ranks of the group (the same
        # tensor on one GPU, the peer exchange's merged record, or an all-gather + ppo_merge_stats with NCCL)
        self._gstats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=self.device) if self.sharded else None
        self.global_stats = self._gstats if self.sharded else self.stats
        self.ws_gae = torch.zeros(max(int(self.lib.ppo_gae_workspace_bytes(self.envs)), 1 << 12),
                                  dtype=torch.uint8, device=self.device)
        self.ws_tvf = torch.zeros(max(int(self.lib.ppo_tvf_workspace_bytes(self.steps, self.envs)), 1 << 12),
                                  dtype=torch.uint8, device=self.device)
        # env-sharded: the moment exchange inside the GAE launch over NVLink peer memory when every rank can take the
        # single cooperative launch for its shard and map its peers (PPO_PEER_EXCHANGE=0 keeps the NCCL all-reduce)
        self.peers = None
        if self.sharded and os.environ.get("PPO_PEER_EXCHANGE", "1") != "0":
            fits = torch.tensor([int(self.lib.ppo_gae_normalized_fits(self.steps, self.envs, _lib.DONE_U8, 1))],
                                device=self.device)
            dist.all_reduce(fits, op=dist.ReduceOp.MIN, group=group)
            if fits.item() == 1 and dist.get_world_size(group) <= _lib.MAX_PEERS:
                box = PeerMailbox(self.device, group)
                self.peers = box if box.ok else None
                if self.peers is not None:
                    self.global_stats = box.global_stats
        self._ctab = self._blob = None
        self._ring = _rt.TableRing(self.device)
        # Every section of the tables until a pass has run on the transposed slab kernel; from then on only what that
        # kernel reads (a third of the host work per rebuild).  The builder itself falls back to all sections for a
        # table the slab kernel cannot serve, and run() rebuilds if the kernel declines a lean table at launch.
        self._sections = _rt.TABLES_ALL
        self._lean_tables = False
        self._samples = None
        self._hz64 = np.ascontiguousarray(self.horizons.reshape(-1), dtype=np.int64)
        # GAE inside the target launch (the transposed slab kernel's scan warp): measured slower than the separate
        # cooperative GAE launch on B200 (0.536 vs 0.506 ms per pass at 256 x 4096 x 128: the target kernel is issue
        # bound and every consumer waits for the scheduler that also carries the scan warp), so it is opt-in
        self.fuse_gae = os.environ.get("PPO_FUSE_GAE", "0") != "0"
        self.gae_fused = False
        self.kernel_launches_per_pass = 4
        self._tvf_launches = 2
        self.tvf_kernel = "?"
        self.value_estimates = self._value_weights = None
        self.ext_advantage = self.combined_stats = None  # allocated by the first pass with an intrinsic advantage

    def close(self):
        """Release the peer mailboxes (env-sharded passes).  Collective in spirit: call it on every rank once no rank
        launches this pass any more (e.g. after a barrier)."""
        if self.peers is not None:
            self.peers.close()
            self.peers = None

    def _merge_global(self):
        """NCCL form of the exchange: all-gather of the per-rank records + ``ppo_merge_stats`` in rank order (identical
        bits on every rank), into ``global_stats``; ``stats`` stays local."""
        if not self.sharded:
            return self.stats
        self.global_stats = self._gstats
        self._gstats.copy_(self.stats)
        allreduce_stats(self._gstats, self.group)
        return self._gstats

    def adv_ratio(self):
        """``((ext_advantage**2).mean() / (int_advantage**2).mean())**0.5`` of the last pass with an intrinsic advantage
        (``adv_ratio`` of ``rl/rollout.py:1230``), from the two statistics records (host float; synchronises)."""
        ext, comb = self.stats.cpu().numpy(), self.combined_stats.cpu().numpy()
        return float(np.sqrt(ext[_lib.STAT_ADV_SUMSQ] / comb[_lib.STAT_RET_SUMSQ]))

    def check_peers(self):
        """Raise ``TimeoutError`` if the in-kernel peer exchange of an earlier pass gave up on a rank that never launched
        (call after the stream has drained; see :class:`PeerMailbox`)."""
        if self.peers is not None:
            self.peers.check()

    def set_samples(self, n_step_samples):
        """Resolve and upload a new ``[K, C]`` sample table: host integers only (about 0.3 ms for 128 horizons x 8
        samples), into rotating pinned/device blobs, one asynchronous copy on the current stream -- call it per pass."""
        self._samples = n_step_samples
        tabs, self._ctab, self._blob = self._ring.build(self.tvf_gamma, self.steps, self._hz64, self._hz64,
                                                        n_step_samples, self.use_log, sections=self._sections)
        self._lean_tables = not tabs.layout.sections & _rt.TABLES_GENERIC
        return tabs

    def run(self, rewards, values, dones, value_samples, tvf_events=None, int_advantage=None, ir_scale=1.0):
        """``rewards [N, A] f32``, ``values [N+1, A] f32`` (row N = final estimate; any strides with a
        unit-stride env axis or a strided view of the value-sample buffer) or ``None``, ``dones [N, A] uint8``,
        ``value_samples [N+1, A, K] f32`` -- all CUDA tensors of this rank's env block.

        ``values=None`` is the reference's own data flow (``rl/rollout.py:1196``): the estimate GAE reads is taken
        off the value curves (``rl/tvf.py:303-360``: last horizon for the TVF gamma, rediscounted otherwise).  The
        target launch then runs first and writes ``self.value_estimates [N+1, A]`` as a by-product of the value rows
        it streams, and GAE reads that contiguous buffer.

        ``int_advantage [N, A] f32`` (``use_intrinsic_rewards``, ``rl/rollout.py:1225-1231``; e.g. the output of
        ``normalization.IntrinsicReturns``): ``advantage = ext_advantage + ir_scale * int_advantage`` and the
        normalisation runs on the sum.  ``ext_advantage`` keeps the extrinsic part, ``stats`` its record,
        ``combined_stats`` the record of the sum (ADV slots) and of the scaled intrinsic part (RET slots)."""
        assert self._ctab is not None, "call set_samples() first"
        if self.peers is not None:
            self.peers.check()  # an earlier pass that gave up on a peer (host word, no synchronisation)
        lib, s = self.lib, dev.stream_ptr()
        n, a, k = self.steps, self.envs, self.n_h

        def targets(extras):
            self._ring.before_launch()  # the tables' upload runs on a side stream
            if tvf_events is not None:  # (start, end) around the target launch(es)
                tvf_events[0].record()
            # targets + statistics of the longest horizon (+ value estimate): one launch on the transposed slab
            # kernel, up to three on the other paths
            rc = lib.ppo_tvf_targets_ex(
                rewards.data_ptr(), rewards.stride(0), dones.data_ptr(), _lib.DONE_U8, dones.stride(0),
                value_samples.data_ptr(), value_samples.stride(0), value_samples.stride(1), n, a, k, k,
                ctypes.byref(self._ctab), self.tvf_path, self.tvf_returns.data_ptr(), self.tvf_stats.data_ptr(),
                self.ws_tvf.data_ptr(), self.ws_tvf.numel(), ctypes.byref(extras) if extras is not None else None, s)
            if rc == _lib.PPO_ERR_UNSUPPORTED and self._lean_tables:
                # the slab kernel declined at launch (e.g. buffer alignment): all sections from now on, and once more
                self._sections = _rt.TABLES_ALL
                self.set_samples(self._samples)
                self._ring.before_launch()
                rc = lib.ppo_tvf_targets_ex(
                    rewards.data_ptr(), rewards.stride(0), dones.data_ptr(), _lib.DONE_U8, dones.stride(0),
                    value_samples.data_ptr(), value_samples.stride(0), value_samples.stride(1), n, a, k, k,
                    ctypes.byref(self._ctab), self.tvf_path, self.tvf_returns.data_ptr(), self.tvf_stats.data_ptr(),
                    self.ws_tvf.data_ptr(), self.ws_tvf.numel(), ctypes.byref(extras) if extras is not None else None, s)
            _lib.check(rc, "ppo_tvf_targets")
            last = lib.ppo_tvf_last_path()
            self.tvf_kernel = {1: "generic", 2: "ring", 3: "tslab", 4: "cslab", 5: "qslab"}.get(last & 0xff, "?")
            if self.tvf_kernel in ("tslab", "qslab") and self.tvf_path == _rt.PATH_AUTO and self._sections == _rt.TABLES_ALL \
                    and os.environ.get("PPO_TVF_LEAN_TABLES", "1") != "0":
                self._sections = _rt.TABLES_TSLAB
            self._tvf_launches = 1 + (0 if last & 0x100 else 1) + (1 if extras is not None and not last & 0x200 else 0)
            if tvf_events is not None:
                tvf_events[1].record()
            self._ring.after_launch()

        if values is None:
            if self.value_estimates is None:
                from . import tvf as _tvf
                self.value_estimates = torch.empty((n + 1, a), dtype=torch.float32, device=self.device)
                self._value_weights = None if abs(self.gamma - self.tvf_gamma) < 1e-8 else \
                    _tvf.rediscount_weights(self.tvf_gamma, self.gamma, [int(h) for h in self.horizons])
            # GAE + TD(lambda) + their statistics ride on the same launch (the kernel's scan warp runs the two
            # recurrences of an env as soon as its value column is complete): what is left is the normalisation
            gae_req = dict(gamma=self.gamma, lambda_adv=self.lambda_policy, lambda_ret=self.lambda_value,
                           adv=self.advantage, ret=self.ext_returns, stats=self.stats, workspace=self.ws_gae) \
                if self.fuse_gae and int_advantage is None else None
            extras, _keep = _rt._extras(self.value_estimates, self._value_weights, gae=gae_req)
            targets(extras)
            values = self.value_estimates
            if extras.gae_done:
                rc = lib.ppo_adv_normalize(self.advantage.data_ptr(), n * a, self._merge_global().data_ptr(), self.eps,
                                           self.clip, self.advantages.data_ptr(), s)
                _lib.check(rc, "ppo_adv_normalize")
                self.gae_fused = True
                self.kernel_launches_per_pass = self._tvf_launches + 1
                return self
        self.gae_fused = False
        if int_advantage is not None:
            # extrinsic scan -> combine (+ statistics of the sum) -> exchange when sharded -> normalisation
            if self.ext_advantage is None:
                self.ext_advantage = torch.empty_like(self.advantage)
                self.combined_stats = torch.zeros(_lib.STATS_LEN, dtype=torch.float64, device=self.device)
            assert tuple(int_advantage.shape) == (n, a) and int_advantage.is_contiguous() and \
                int_advantage.dtype == torch.float32, "int_advantage must be a contiguous float32 [N, A] CUDA tensor"
            rc = lib.ppo_gae(rewards.data_ptr(), rewards.stride(0), values.data_ptr(), values.stride(0),
                             values.stride(1), values[n].data_ptr(), values.stride(1), dones.data_ptr(), _lib.DONE_U8,
                             dones.stride(0), n, a, self.gamma, self.lambda_policy, self.lambda_value,
                             self.ext_advantage.data_ptr(), self.ext_returns.data_ptr(), self.stats.data_ptr(),
                             self.ws_gae.data_ptr(), self.ws_gae.numel(), s)
            _lib.check(rc, "ppo_gae")
            rc = lib.ppo_adv_combine(self.ext_advantage.data_ptr(), int_advantage.data_ptr(), float(ir_scale), n * a,
                                     self.advantage.data_ptr(), self.combined_stats.data_ptr(), self.ws_gae.data_ptr(),
                                     self.ws_gae.numel(), s)
            _lib.check(rc, "ppo_adv_combine")
            record = self.combined_stats
            if self.sharded:
                self.global_stats = self._gstats
                self._gstats.copy_(self.combined_stats)
                record = allreduce_stats(self._gstats, self.group)
            rc = lib.ppo_adv_normalize(self.advantage.data_ptr(), n * a, record.data_ptr(), self.eps, self.clip,
                                       self.advantages.data_ptr(), s)
            _lib.check(rc, "ppo_adv_normalize")
            if values is not self.value_estimates:
                targets(None)
            self.kernel_launches_per_pass = self._tvf_launches + 3 + (1 if self.sharded else 0)
            return self
        gae_launches = 2
        peer_rc = 0
        if self.sharded and self.peers is not None and values.stride(1) == 1:
            # moments exchanged inside the launch (NVLink peer mailboxes): GAE, TD(lambda), statistics, the cross-GPU
            # sum and the normalisation in ONE cooperative launch per rank
            box = self.peers
            peer_rc = lib.ppo_gae_normalized_peers(
                rewards.data_ptr(), rewards.stride(0), values.data_ptr(), values.stride(0), values.stride(1),
                values[n].data_ptr(), values.stride(1), dones.data_ptr(), _lib.DONE_U8, dones.stride(0), n, a,
                self.gamma, self.lambda_policy, self.lambda_value, self.eps, self.clip, self.advantage.data_ptr(),
                self.ext_returns.data_ptr(), self.advantages.data_ptr(), self.stats.data_ptr(),
                box.global_stats.data_ptr(), self.ws_gae.data_ptr(), self.ws_gae.numel(), box.rank, box.world,
                box.mail_ptrs, box.flag_ptrs, box.next_epoch(), box.status.data_ptr(), box.timeout_ms, s)
            _lib.check(peer_rc, "ppo_gae_normalized_peers")
            assert peer_rc == 1, "ppo_gae_normalized_peers declined a shard shape ppo_gae_normalized_fits accepted"
            self.global_stats = box.global_stats
            gae_launches = 1
        if peer_rc == 1:
            pass
        elif self.sharded:
            rc = lib.ppo_gae(rewards.data_ptr(), rewards.stride(0), values.data_ptr(), values.stride(0),
                             values.stride(1), values[n].data_ptr(), values.stride(1), dones.data_ptr(), _lib.DONE_U8,
                             dones.stride(0), n, a, self.gamma, self.lambda_policy, self.lambda_value,
                             self.advantage.data_ptr(), self.ext_returns.data_ptr(), self.stats.data_ptr(),
                             self.ws_gae.data_ptr(), self.ws_gae.numel(), s)
            _lib.check(rc, "ppo_gae")
            rc = lib.ppo_adv_normalize(self.advantage.data_ptr(), n * a, self._merge_global().data_ptr(), self.eps,
                                       self.clip, self.advantages.data_ptr(), s)
            _lib.check(rc, "ppo_adv_normalize")
            gae_launches = 3  # scan, merge of the gathered records, normalisation
        else:
            # one GPU: no exchange between the scan and the normalisation -> one cooperative launch when the grid fits
            rc = lib.ppo_gae_normalized(rewards.data_ptr(), rewards.stride(0), values.data_ptr(), values.stride(0),
                                        values.stride(1), values[n].data_ptr(), values.stride(1), dones.data_ptr(),
                                        _lib.DONE_U8, dones.stride(0), n, a, self.gamma, self.lambda_policy,
                                        self.lambda_value, self.eps, self.clip, self.advantage.data_ptr(),
                                        self.ext_returns.data_ptr(), self.advantages.data_ptr(), self.stats.data_ptr(),
                                        self.ws_gae.data_ptr(), self.ws_gae.numel(), s)
            _lib.check(rc, "ppo_gae_normalized")
            gae_launches = 1 if rc == 1 else 2
        if values is not self.value_estimates:
            targets(None)
        self.kernel_launches_per_pass = self._tvf_launches + gae_launches
        return self


def stats_summary(stats):
    """mean / std / min / max of the advantages and returns from a statistics record (host floats),
    the quantities the reference logs at ``rl/rollout.py:1252-1256``."""
    s = stats.detach().cpu().numpy()
    n = max(s[_lib.STAT_COUNT], 1.0)

    def ms(total, sq):
        mean = total / n
        return mean, float(np.sqrt(max(sq / n - mean * mean, 0.0)))

    adv_mean, adv_std = ms(s[_lib.STAT_ADV_SUM], s[_lib.STAT_ADV_SUMSQ])
    ret_mean, ret_std = ms(s[_lib.STAT_RET_SUM], s[_lib.STAT_RET_SUMSQ])
    nv = max(s[_lib.STAT_VAL_COUNT], 1.0)
    val_mean = s[_lib.STAT_VAL_SUM] / nv
    val_std = float(np.sqrt(max(s[_lib.STAT_VAL_SUMSQ] / nv - val_mean * val_mean, 0.0)))
    return {"count": int(s[_lib.STAT_COUNT]), "value_mean": val_mean, "value_std": val_std,
            "adv_sparsity": s[_lib.STAT_ADV_ZEROS] / n, "ret_sparsity": s[_lib.STAT_RET_ZEROS] / n,
            "adv_mean": adv_mean, "adv_std": adv_std, "adv_min": s[_lib.STAT_ADV_MIN], "adv_max": s[_lib.STAT_ADV_MAX],
            "ret_mean": ret_mean, "ret_std": ret_std, "ret_min": s[_lib.STAT_RET_MIN], "ret_max": s[_lib.STAT_RET_MAX]}


class RolloutBuffers:
    """Device-resident form of the rollout buffers the return stage consumes (``rl/rollout.py:646-680``,
    ``rl/tvf.py:27-30``) and of the writes ``Runner.generate_rollout`` makes into them once per env step
    (``rl/rollout.py:790-813``) and once after the last step (``:878-892``).  The reference pulls every model output
    to the host (``model_out["tvf_value"].cpu().numpy()``), trims it there and stores it in NumPy arrays, which the
    return stage then uploads again; here the model outputs stay where the forward pass left them:

      ``write_step(t, model_out, prev_time, ext_rewards, dones)``   rows ``t`` of ``tvf_untrimmed_value``, ``tvf_value``,
                                                                    ``tvf_final_value``, ``all_time``, ``ext_rewards``,
                                                                    ``terminals`` (+ ``value``, ``log_policy`` when given)
      ``write_final(model_out, time)``                              row ``N`` of the value buffers
      ``returns(pass_)``                                            ``ReturnPass.run`` on the buffers, no copies

    ``tvf`` is a :class:`ppo_b200.tvf.TVFRunnerModule` (its ``trim_horizons`` does the trimming with the reference's
    arguments); rewards, terminals and times come from the env workers as host arrays, 5 bytes per env and step.
    """

    def __init__(self, steps, envs, horizons, tvf, value_heads=1, trimming="off", trimming_mode="average", device=None):
        self.device = device or dev.require_cuda()
        self.steps, self.envs, self.tvf = int(steps), int(envs), tvf
        self.trimming, self.trimming_mode = trimming, trimming_mode
        k = len(horizons)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.tvf_value = torch.zeros((steps + 1, envs, k, value_heads), **f32)
        self.tvf_untrimmed_value = torch.zeros((steps + 1, envs, k, value_heads), **f32)
        self.tvf_final_value = torch.zeros((steps + 1, envs), **f32)
        self.ext_rewards = torch.zeros((steps, envs), **f32)
        self.terminals = torch.zeros((steps, envs), dtype=torch.uint8, device=self.device)
        self.all_time = torch.zeros((steps + 1, envs), dtype=torch.int32, device=self.device)
        self.value = self.log_policy = None  # allocated by the first write that carries them
        self.ttt_predictions = [[] for _ in range(envs)]

    def _tvf_row(self, t, tvf_value, time):
        tv = dev.to_device(tvf_value, torch.float32, self.device).detach()
        if tv.dim() == 2:
            tv = tv.unsqueeze(-1)
        row = self.tvf_untrimmed_value[t]
        row.copy_(tv)
        row[:, 0, :] = 0  # "first value head should always be zero" (rl/rollout.py:792, :883)
        trimmed, final, ttt = self.tvf.trim_horizons(row, time, method=self.trimming, mode=self.trimming_mode)
        self.tvf_value[t].copy_(trimmed if trimmed.shape == row.shape else trimmed.expand_as(row))
        if torch.is_tensor(final) or isinstance(final, np.ndarray):
            self.tvf_final_value[t].copy_(dev.to_device(final, torch.float32, self.device))
        else:
            self.tvf_final_value[t].fill_(float(final))
        if ttt is not None:  # rl/rollout.py:800-802
            for a, x in enumerate(np.asarray(ttt)):
                self.ttt_predictions[a].append(x)
        self.all_time[t].copy_(torch.as_tensor(np.asarray(time), dtype=torch.int32), non_blocking=True)

    def write_step(self, t, model_out, prev_time, ext_rewards, dones):
        self._tvf_row(t, model_out["tvf_value"], prev_time)
        self.ext_rewards[t].copy_(torch.as_tensor(np.asarray(ext_rewards, dtype=np.float32)), non_blocking=True)
        self.terminals[t].copy_(torch.as_tensor(np.asarray(dones).astype(np.uint8)), non_blocking=True)
        for name in ("value", "log_policy"):
            if name in model_out:
                x = model_out[name].detach()
                if getattr(self, name) is None:
                    rows = self.steps + 1 if name == "value" else self.steps
                    setattr(self, name, torch.zeros((rows,) + tuple(x.shape), dtype=x.dtype, device=self.device))
                getattr(self, name)[t].copy_(x)

    def write_final(self, model_out, time):
        self._tvf_row(self.steps, model_out["tvf_value"], time)
        if "value" in model_out and self.value is not None:
            self.value[self.steps].copy_(model_out["value"].detach())

    def returns(self, pass_, value_head=0, **kwargs):
        """One return pass over the buffers (``ReturnPass.run`` with ``values=None``: the estimate GAE reads comes off
        the value curves, ``rl/rollout.py:1196``)."""
        vs = self.tvf_value[..., value_head]
        if not vs.is_contiguous():
            vs = vs.contiguous()
        return pass_.run(self.ext_rewards, None, self.terminals, vs, **kwargs)
