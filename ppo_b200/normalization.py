"""Return trackers and running moments around the hot path, on the device (``ppo_b200/csrc/return_tracker.cu``).

Mirrors, with the reference's names and state:

* ``RunningMeanStd``                       ``rl/utils.py:414-443`` (scalar shape); state {mean, var, count} lives on the GPU
* ``IntrinsicReturns.calculate_intrinsic_returns``   ``Runner.calculate_intrinsic_returns``, ``rl/rollout.py:1140-1180``
  (state ``ems_norm`` [A] float64 and ``intrinsic_returns_rms``, ``rl/rollout.py:279-280``)
* ``VecNormalizeRewards.normalize_rollout``          ``VecNormalizeRewardWrapper.step`` (``rl/wrappers.py:795-895``, mode
  "rms") applied to the N rows of a rollout in order (state ``current_returns`` [A] and ``ret_rms``)

Both are the forward recurrence ``x_t = r_t + gamma * (1 - done_t) * x_{t-1}`` followed, step by step, by a merge of the
batch moments of ``x_t`` over the envs into the running moments.  The reference does N NumPy passes; here one scan
launch, two row-reduction launches and one fold.  With envs sharded over ranks (``group``; uneven shards allowed) the only exchange is
the all-reduce of the two [N] row vectors (``tools/dist_check.py`` checks the sharded result against the unsharded one
on real GPUs).  No CPU fallback.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _device as dev
from . import _lib
from . import returns as _returns
from . import rollout as _rollout


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


_env_totals = {}


def global_env_count(envs, group=None, device=None):
    """Envs over all ranks of ``group``.  Shards may be uneven (``rollout.shard_envs`` gives the remainder to the low
    ranks), so the count is summed over the ranks -- one tiny all-reduce the first time a (group, local count) pair is
    seen, cached afterwards: every rank must keep its shard size between calls (collective on first use)."""
    if _world(group) == 1:
        return int(envs)
    key = (id(group) if group is not None else 0, int(envs))
    total = _env_totals.get(key)
    if total is None:
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor([int(envs)], dtype=torch.int64, device=(device or dev.require_cuda()) if on_gpu else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        total = _env_totals[key] = int(t.item())
    return total


class RunningMeanStd:
    """Running mean / variance book-keeping (``rl/utils.py:414-443``), scalar shape, state on the device."""

    def __init__(self, epsilon=1e-4, shape=(), device=None):
        if tuple(shape) != ():
            raise ValueError("only the scalar RunningMeanStd (shape=()) of the return trackers is provided")
        self.device = device or dev.require_cuda()
        self.state = torch.tensor([0.0, 1.0, float(epsilon)], dtype=torch.float64, device=self.device)

    mean = property(lambda self: float(self.state[0].item()))
    var = property(lambda self: float(self.state[1].item()))
    count = property(lambda self: float(self.state[2].item()))

    def update(self, x, group=None):
        """One merge of the batch ``x`` (1-D; a float / int counts as a batch of one, ``rl/utils.py:427-430``)."""
        if type(x) in (float, int):
            self.update_from_moments(x, 0.0, 1)
            return
        t = dev.to_device(x, torch.float64, self.device).reshape(1, -1).contiguous()
        _fold_rows(t, self.state, group, None)

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        lib = _lib.load()
        n = float(batch_count)
        sums = torch.tensor([float(batch_mean) * n], dtype=torch.float64, device=self.device)
        sq = torch.tensor([float(batch_var) * n], dtype=torch.float64, device=self.device)
        _lib.check(lib.ppo_rms_fold(dev.ptr(sums), dev.ptr(sq), 1, n, dev.ptr(self.state), 0, dev.stream_ptr()),
                   "ppo_rms_fold")

    def save_state(self):
        s = self.state.cpu().numpy()
        return tuple([s[0], s[1], s[2]])

    def restore_state(self, state):
        self.state.copy_(torch.as_tensor([float(v) for v in state], dtype=torch.float64))


def _fold_rows(x, rms_state, group, row_var, ema=None):
    """Merge every row of the float64 [N, A] tracker ``x`` into ``rms_state`` in row order.  ``ema``: ``(ema_var device
    scalar, ema_horizon)`` for the wrapper's mode "ema" (``row_var`` then records the EMA variance per step)."""
    lib = _lib.load()
    steps, envs = x.shape
    s = dev.stream_ptr()
    sums = torch.empty(steps, dtype=torch.float64, device=x.device)
    sq = torch.empty(steps, dtype=torch.float64, device=x.device)
    if _world(group) == 1:  # nothing to exchange: both row passes and the merges in one launch
        ws = dev.workspace(256, x.device)
        rc = lib.ppo_row_moments_fold(dev.ptr(x), steps, envs, dev.ptr(sums), dev.ptr(sq), dev.ptr(rms_state),
                                      dev.ptr(ema[0]) if ema is not None else None, float(ema[1]) if ema is not None else 0.0,
                                      dev.ptr(row_var), dev.ptr(ws), s)
        _lib.check(rc, "ppo_row_moments_fold")
        return
    total = float(global_env_count(envs, group, x.device))  # the true count: shards need not be equal
    _lib.check(lib.ppo_row_sums(dev.ptr(x), steps, envs, 0, total, dev.ptr(sums), s), "ppo_row_sums")
    if _world(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    _lib.check(lib.ppo_row_sums(dev.ptr(x), steps, envs, dev.ptr(sums), total, dev.ptr(sq), s), "ppo_row_sums")
    if _world(group) > 1:
        dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=group)
    if ema is not None:
        _lib.check(lib.ppo_rms_fold_ema(dev.ptr(sums), dev.ptr(sq), steps, total, dev.ptr(rms_state), dev.ptr(ema[0]),
                                        float(ema[1]), dev.ptr(row_var), s), "ppo_rms_fold_ema")
        return
    _lib.check(lib.ppo_rms_fold(dev.ptr(sums), dev.ptr(sq), steps, total, dev.ptr(rms_state), dev.ptr(row_var), s),
               "ppo_rms_fold")


def track_returns(rewards, dones, gamma, carry, rms, group=None, want_row_var=False, ema=None, transform=None):
    """The tracker scan + the N moment merges.  ``carry`` (float64 device tensor [A]) and ``rms`` are updated in
    place; returns (rewards on the device, running variance after each step [N] or None).  ``transform`` (optional): a
    callable applied to the float64 [N, A] tracker rows on the device before they enter the running moments
    (``returns_transform(current_returns / norms)``, ``rl/wrappers.py:876``); the EMA of the wrapper's mode "ema" keeps
    following the variance of the untransformed tracker (:881)."""
    device = dev.require_cuda()
    lib = _lib.load()
    r = dev.env_contiguous(dev.to_device(rewards, torch.float32, device))
    if r.dim() != 2:
        raise ValueError(f"rewards must be [N, A], got {tuple(r.shape)}")
    steps, envs = r.shape
    if tuple(carry.shape) != (envs,) or carry.dtype != torch.float64 or not carry.is_cuda:
        raise ValueError(f"carry must be a float64 CUDA tensor of shape {(envs,)}")
    d, kind = dev.done_tensor(dones, device)
    if d is not None:
        if tuple(d.shape) != (steps, envs):
            raise ValueError(f"operands could not be broadcast together: dones {tuple(d.shape)} vs {(steps, envs)}")
        d = dev.env_contiguous(d)
    x = torch.empty((steps, envs), dtype=torch.float64, device=device)
    rc = lib.ppo_return_tracker_scan(dev.ptr(r), r.stride(0) if steps else 0, dev.ptr(d), kind,
                                     d.stride(0) if d is not None and steps else 0, steps, envs, float(gamma),
                                     dev.ptr(carry), dev.ptr(x), dev.stream_ptr())
    _lib.check(rc, "ppo_return_tracker_scan")
    row_var = torch.empty(steps, dtype=torch.float64, device=device) if want_row_var else None
    if steps and envs:
        if transform is None:
            _fold_rows(x, rms.state, group, row_var, ema=ema)
        else:
            y = transform(x).to(torch.float64).contiguous()
            if tuple(y.shape) != (steps, envs):
                raise ValueError(f"the returns transform must keep the shape {(steps, envs)}, got {tuple(y.shape)}")
            if ema is None:
                _fold_rows(y, rms.state, group, row_var)
            else:
                # the moments take the transformed returns; the EMA weights only need the counts after each merge, which
                # are the same for any data, so the raw tracker is folded into a scratch copy of the state
                scratch = rms.state.clone()
                _fold_rows(y, rms.state, group, None)
                _fold_rows(x, scratch, group, row_var, ema=ema)
    return r, row_var


def episodic_normalization_constant(times, discount_type, gamma=1.0, discount_bias=1.0):
    """``EpisodicDiscounting.get_normalization_constant`` (``rl/wrappers.py:52-72``) on a device tensor of env times.
    With the wrapper's ``gamma=1.0`` (:872) "finite" raises ZeroDivisionError (plain Python arithmetic, :55) and
    "geometric" gives infinite normalisers (:60), like the reference."""
    k = times.to(torch.float64) + float(discount_bias)
    if discount_type == "finite":
        m = 1 / (1 - gamma)
        return torch.clamp(m - k, min=0.0) + 1
    if discount_type == "geometric":
        return (gamma ** k) / (1 - gamma)
    if discount_type == "quadratic":
        return 1 / k
    if discount_type == "power":
        return (1 / 1e-1) * (k ** -1e-1)
    if discount_type == "harmonic":
        return 1 / torch.log(k)
    if discount_type == "none":
        return torch.ones_like(k)
    raise ValueError(f"Invalid discount_type {discount_type}")


class ReturnTracker:
    """``ReturnTracker`` of ``rl/wrappers.py:737-755`` (one per env type in the multi-env reward normaliser, :946-971):
    float32 ``current_returns`` [A] and ``ret_rms`` on the device.  ``update`` is the reference's per-step call;
    ``update_rollout`` takes the N steps of a rollout in one go (two launches, ``ppo_return_tracker_masked``) and
    returns the running variance after every step."""

    def __init__(self, num_envs, gamma, device=None):
        self.device = device or dev.require_cuda()
        self.ret_rms = RunningMeanStd(shape=(), device=self.device)
        self.current_returns = torch.zeros(num_envs, dtype=torch.float32, device=self.device)
        self.gamma = gamma

    def reset(self):
        self.current_returns.zero_()

    def update(self, rewards, dones, mask):
        """One step: rewards / dones / mask are [A] (``rl/wrappers.py:749-755``)."""
        def row(a, dtype=None):
            t = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a)
            return t.reshape(1, -1) if dtype is None else t.to(dtype).reshape(1, -1)
        self.update_rollout(row(rewards), row(dones), row(mask))

    def update_rollout(self, rewards, dones, mask):
        """[N, A] rewards, dones and per-step masks -> the running variance after each step [N] (float64, device)."""
        lib = _lib.load()
        device = self.device
        r = dev.env_contiguous(dev.to_device(rewards, torch.float32, device))
        if r.dim() != 2 or r.shape[1] != self.current_returns.numel():
            raise ValueError(f"rewards must be [N, {self.current_returns.numel()}], got {tuple(r.shape)}")
        steps, envs = r.shape
        d, kind = dev.done_tensor(dones, device)
        if d is not None:
            if tuple(d.shape) != (steps, envs):
                raise ValueError(f"operands could not be broadcast together: dones {tuple(d.shape)} vs {(steps, envs)}")
            d = dev.env_contiguous(d)
        m = torch.as_tensor(np.asarray(mask) if not isinstance(mask, torch.Tensor) else mask)
        if tuple(m.shape) != (steps, envs):
            raise ValueError(f"mask must be [N, A] = {(steps, envs)}, got {tuple(m.shape)}")
        m = (m != 0).to(device=device, dtype=torch.uint8).contiguous()
        row_var = torch.empty(steps, dtype=torch.float64, device=device)
        if steps == 0 or envs == 0:
            return row_var
        x = torch.empty((steps, envs), dtype=torch.float64, device=device)
        rec = torch.empty(3 * steps, dtype=torch.float64, device=device)
        ws = dev.workspace(256, device)
        rc = lib.ppo_return_tracker_masked(dev.ptr(r), r.stride(0), dev.ptr(d), kind, d.stride(0) if d is not None else 0,
                                           dev.ptr(m), m.stride(0), steps, envs, float(self.gamma),
                                           dev.ptr(self.current_returns), dev.ptr(x), dev.ptr(rec),
                                           dev.ptr(self.ret_rms.state), dev.ptr(row_var), dev.ptr(ws), dev.stream_ptr())
        _lib.check(rc, "ppo_return_tracker_masked")
        return row_var


def _scale(r, var, per_row, mode, eps, clip, scale, center_stats, clips):
    lib = _lib.load()
    steps, envs = r.shape
    out = torch.empty((steps, envs), dtype=torch.float32, device=r.device)
    rc = lib.ppo_scale_rewards(dev.ptr(r), r.stride(0) if steps else 0, steps, envs, dev.ptr(var), int(per_row), mode,
                               float(eps), float(clip), float(scale), dev.ptr(center_stats), dev.ptr(out),
                               dev.ptr(clips), dev.stream_ptr())
    _lib.check(rc, "ppo_scale_rewards")
    return out


class IntrinsicReturns:
    """The runner state and method of ``rl/rollout.py:279-280, 1140-1180``."""

    def __init__(self, envs, device=None):
        self.device = device or dev.require_cuda()
        self.ems_norm = torch.zeros(envs, dtype=torch.float64, device=self.device)
        self.intrinsic_returns_rms = RunningMeanStd(shape=(), device=self.device)
        self.intrinsic_reward_norm_scale = 1.0

    def calculate_intrinsic_returns(self, int_rewards, int_value, terminals, gamma_int, lambda_policy=0.95,
                                    normalize=True, center=False, propagation=False, group=None):
        """``args.ir.normalize / center / propagation`` are the keyword arguments.  Returns a dict with
        ``int_advantage`` [N, A] (the reference's return value), ``int_returns`` [N, A] (``self.int_returns``),
        ``int_rewards`` (the normalised rewards the reference writes back to ``self.int_rewards``) and ``scale``
        (a device float64 scalar; ``self.intrinsic_reward_norm_scale`` reads it)."""
        np_like = dev.is_numpy_like(int_rewards)
        device = self.device
        r = dev.env_contiguous(dev.to_device(int_rewards, torch.float32, device))
        steps, envs = r.shape
        value = dev.to_device(int_value, torch.float32, device)
        if tuple(value.shape) != (steps + 1, envs):
            raise ValueError(f"int_value must be [N + 1, A] = {(steps + 1, envs)}, got {tuple(value.shape)}")
        dones = None if propagation else terminals  # rl/rollout.py:1151,1174: `(not propagation) * terminals`
        scale = torch.ones((), dtype=torch.float64, device=device)
        if normalize:
            track_returns(r, dones, gamma_int, self.ems_norm, self.intrinsic_returns_rms, group=group)
            var = self.intrinsic_returns_rms.state[1:2]
            scale = 1e-5 + var[0].sqrt()  # rl/rollout.py:1164 (device scalar, no synchronisation)
        else:
            var = torch.zeros(1, dtype=torch.float64, device=device)
        if normalize or center:
            stats = None
            if center:  # rl/rollout.py:1167-1168: the mean over every env of every rank
                stats = _rollout.allreduce_moments(_rollout.moments(r), group)
            # eps 1e-5 (rl/rollout.py:1164); without normalisation the divisor is exactly 1
            r = _scale(r, var if normalize else torch.ones_like(var), False, _lib.SCALE_INTRINSIC,
                       1e-5 if normalize else 0.0, -1.0, 1.0, stats, None)
        adv, ret, _ = _returns._gae_launch(r, value[:steps], value[steps], dones, gamma_int, lambda_policy,
                                           lambda_policy, True, True, False)
        self.intrinsic_reward_norm_scale = scale
        return {"int_advantage": dev.like_input(adv, np_like), "int_returns": dev.like_input(ret, np_like),
                "int_rewards": dev.like_input(r, np_like), "scale": scale}


class VecNormalizeRewards:
    """``VecNormalizeRewardWrapper`` (``rl/wrappers.py:795-919``) over whole rollouts: rewards are
    divided by the return std AS OF THEIR OWN STEP, clipped and scaled.  ``mode``: "rms" (running variance over the
    whole history, what ``rl/envs.py:160-165`` configures), "ema" (``ret_var`` follows the variance of the tracker over
    the envs with ``alpha = 1 - A / min(count, ema_horizon)``, :878-881) or "custom" (the caller sets ``ret_var``).
    ``returns_transform`` (a callable on the device tensor of tracker rows) and the episodic-discounting normalisers
    ``ed_type`` / ``ed_bias`` (:870-876; ``normalize_rollout`` then takes the envs' ``info['time']`` as ``times``
    [N, A]) shape what enters the running moments, like in the reference."""

    def __init__(self, num_envs, initial_state=None, gamma=1.0, clip=10.0, scale=1.0, returns_transform=None, mode="rms",
                 ed_type=None, ed_bias=1.0, ema_horizon=5e6, device=None):
        if mode not in ("rms", "ema", "custom"):
            raise ValueError(f"Invalid mode {mode}")
        self.returns_transform, self.ed_type, self.ed_bias = returns_transform, ed_type, ed_bias
        self.device = device or dev.require_cuda()
        self.clip, self.epsilon, self.gamma, self.scale = clip, 1e-2, gamma, scale
        self.mode, self.ema_horizon = mode, float(ema_horizon)
        self.current_returns = torch.zeros(num_envs, dtype=torch.float64, device=self.device)
        self.ret_rms = RunningMeanStd(shape=(), device=self.device)
        self._ret_var = torch.zeros(1, dtype=torch.float64, device=self.device)
        if initial_state is not None:
            self.ret_rms.restore_state(initial_state)

    def reset(self):
        self.current_returns.zero_()

    @property
    def mean(self):
        return self.ret_rms.mean

    @property
    def ret_var(self):
        return float(self._ret_var.item())

    @ret_var.setter
    def ret_var(self, value):
        self._ret_var.fill_(float(value))

    @property
    def std(self):
        if self.mode == "rms":
            return float(np.sqrt(self.ret_rms.var + self.epsilon))
        return float(np.sqrt(self.ret_var + self.epsilon))

    def normalize_rollout(self, rewards, dones, group=None, times=None):
        """[N, A] raw rewards and dones -> (scaled rewards [N, A] float32, number of clipped elements as a device
        counter).  ``times`` [N, A]: the envs' ``info['time']`` per step, needed with ``ed_type`` (missing infos count
        as 0 like ``info.get('time', 0)``, :871)."""
        np_like = dev.is_numpy_like(rewards)
        ema = (self._ret_var, self.ema_horizon) if self.mode == "ema" else None
        transform = None
        if self.ed_type is not None or self.returns_transform is not None:
            norms = None
            if self.ed_type is not None:
                if times is None:
                    raise ValueError("ed_type needs the envs' times [N, A]")
                t = torch.as_tensor(np.asarray(times) if not isinstance(times, torch.Tensor) else times).to(self.device)
                norms = episodic_normalization_constant(t, self.ed_type, discount_bias=self.ed_bias)  # :872
            f = self.returns_transform if self.returns_transform is not None else (lambda v: v)
            transform = lambda x: f(x if norms is None else x / norms)  # noqa: E731  (:876)
        r, row_var = track_returns(rewards, dones, self.gamma, self.current_returns, self.ret_rms, group=group,
                                   want_row_var=self.mode != "custom", ema=ema, transform=transform)
        clips = torch.zeros(1, dtype=torch.int64, device=self.device)
        clip = -1.0 if self.clip is None else float(self.clip)
        if self.mode == "custom":  # one fixed variance for the whole rollout
            out = _scale(r, self._ret_var, False, _lib.SCALE_WRAPPER, self.epsilon, clip, self.scale, None, clips)
        else:
            out = _scale(r, row_var, True, _lib.SCALE_WRAPPER, self.epsilon, clip, self.scale, None, clips)
        return dev.like_input(out, np_like), clips

    def save_state(self, buffer):
        buffer["ret_rms"] = self.ret_rms.save_state()
        buffer["ret_var"] = self.ret_var
        buffer["current_returns"] = self.current_returns.cpu().numpy()

    def restore_state(self, buffer):
        self.ret_var = buffer.get("ret_var", 0.0)
        self.ret_rms.restore_state(buffer["ret_rms"])
        self.current_returns.copy_(torch.as_tensor(np.asarray(buffer["current_returns"], dtype=np.float64)))
