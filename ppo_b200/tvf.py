"""Drop-in for the return-side pieces of the reference's ``rl/tvf.py``: the horizon table generator and the
``TVFRunnerModule`` methods that feed / call the hot path (``calculate_tvf_returns`` :210-271,
``get_tvf_ext_value_estimate`` :303-360), trimming (:91-208) and the value loss that consumes the targets
(``on_train_value_minibatch`` :32-77, forward and gradient in one CUDA pass behind ``torch.autograd``).
"""
import collections
import math
import time as clock

import numpy as np
import torch

from . import _device as dev
from . import _lib
from .returns_truncated import get_return_estimate


def get_value_head_horizons(n_heads: int, max_horizon: int, spacing: str = "geometric", include_weight=False):
    """Horizons spaced (approximately) geometrically with H[0] = 0 and H[-1] = max_horizon; duplicates caused
    by rounding are merged and counted as weights (reference: ``rl/tvf.py:576-610``).  Host-only integers."""
    if spacing == "linear":
        result = np.asarray(np.round(np.linspace(0, max_horizon, n_heads)), dtype=np.int32)
        return (result, np.ones([n_heads], dtype=np.float32)) if include_weight else result
    elif spacing == "geometric":
        def get_heads(x, remove_duplicates=True):
            result = np.asarray(np.round(np.geomspace(1, max_horizon + 1, x)) - 1, dtype=np.int32)
            return np.asarray(sorted(set(result))) if remove_duplicates else np.asarray(sorted(result))

        try_heads = n_heads
        while len(get_heads(try_heads)) != n_heads:
            if len(get_heads(try_heads)) < n_heads:
                try_heads += int(math.sqrt(n_heads))
            else:
                try_heads -= 1
        counts = collections.OrderedDict()
        for head in get_heads(try_heads, remove_duplicates=False):
            counts[head] = counts.get(head, 0) + 1
        result = np.asarray(list(counts.keys()), dtype=np.int32)
        weights = np.asarray(list(counts.values()), dtype=np.float32)
        return (result, weights) if include_weight else result
    raise ValueError(f"Invalid spacing value {spacing}")


def rediscount_ratios(old_gamma, new_gamma, horizons, clipping=10):
    """Clipped ratios ``rho_i = min((new_gamma / old_gamma)^mid_i, clipping)`` with
    ``mid_i = (h_{i-1} + 1 + h_i) / 2 - 1`` for i = 1..K-1, as Python floats exactly like ``rl/tvf.py:420-429``;
    ``rho_0 = 0`` (unused).  Host integers / floats only."""
    rho = np.zeros(len(horizons), dtype=np.float64)
    prev_h = 0
    for i in range(1, len(horizons)):
        h = horizons[i]
        mid_h = ((prev_h + 1 + h) / 2) - 1
        prev_h = h
        rho[i] = min((new_gamma ** mid_h) / (old_gamma ** mid_h), clipping)
    return rho


def rediscount_weights(old_gamma, new_gamma, horizons, clipping=10):
    """Per-horizon weights ``w`` with ``sum_i (V_i - V_{i-1}) rho_i == sum_j w_j V_j``: the rediscounted estimate as
    one dot product per value curve (what ``ppo_value_estimate`` and ``ppo_tvf_targets_ex`` consume).  float32 [K]."""
    rho = rediscount_ratios(old_gamma, new_gamma, horizons, clipping)
    w = np.zeros(len(rho), dtype=np.float64)
    if len(rho) > 1:
        w[0] = -rho[1]
        w[1:-1] = rho[1:-1] - rho[2:]
        w[-1] = rho[-1]
    return w.astype(np.float32)


def get_rediscounted_value_estimate(values, old_gamma: float, new_gamma: float, horizons, clipping=10):
    """Rediscounted return at the longest horizon, ``[B, K] -> [B]`` float32 (reference: ``rl/tvf.py:388-433``;
    NumPy in -> NumPy out, tensor in -> tensor out, ``old_gamma == new_gamma`` -> ``values[:, -1]``)."""
    n_rows, n_h = values.shape
    if old_gamma == new_gamma:
        return values[:, -1]
    assert n_h == len(horizons), f"missmatch {n_h} {horizons}"
    assert horizons[0] == 0, 'first horizon must be 0'
    device = dev.require_cuda()
    lib = _lib.load()
    np_like = dev.is_numpy_like(values)
    v = dev.env_contiguous(dev.to_device(values, torch.float32, device))
    w = rediscount_weights(old_gamma, new_gamma, [int(h) for h in horizons], clipping)
    out = torch.empty(n_rows, dtype=torch.float32, device=device)
    rc = lib.ppo_value_estimate(dev.ptr(v), n_rows, v.stride(0) if n_rows else n_h, n_h, w.ctypes.data, dev.ptr(out),
                                dev.stream_ptr())
    _lib.check(rc, "ppo_value_estimate")
    return dev.like_input(out, np_like)


def rediscount_horizons(values, old_gamma: float, new_gamma: float, horizons, clipping=10):
    """Every horizon rediscounted: ``out[:, k]`` is the estimate from the first ``k + 1`` horizons, ``[B, K] ->
    [B, K]`` (reference: the loop of ``TVFRunnerModule.rediscount_horizons``, ``rl/tvf.py:362-380``)."""
    if old_gamma == new_gamma:
        return values
    n_rows, n_h = values.shape
    assert n_h == len(horizons) and horizons[0] == 0, 'first horizon must be 0'
    device = dev.require_cuda()
    lib = _lib.load()
    np_like = dev.is_numpy_like(values)
    v = dev.env_contiguous(dev.to_device(values, torch.float32, device))
    rho = rediscount_ratios(old_gamma, new_gamma, [int(h) for h in horizons], clipping).astype(np.float32)
    out = torch.empty((n_rows, n_h), dtype=torch.float32, device=device)
    rc = lib.ppo_rediscount_horizons(dev.ptr(v), n_rows, v.stride(0) if n_rows else n_h, n_h, rho.ctypes.data,
                                     dev.ptr(out), n_h, dev.stream_ptr())
    _lib.check(rc, "ppo_rediscount_horizons")
    return dev.like_input(out, np_like)


def horizon_interpolate(horizons, values, target_horizons):
    """Linearly interpolated value at a per-example target horizon: ``values [*shape, K]``, ``target_horizons
    [*shape]`` -> ``[*shape]`` float64 (reference: ``rl/tvf.py:527-573``, same assertions)."""
    horizons = np.asarray(horizons)
    assert len(set(horizons.tolist())) == len(horizons), f"Horizons duplicates not supported {horizons}"
    assert np.all(np.diff(horizons) > 0), \
        f"Horizons must be sorted and unique horizons:{horizons}, targets:{target_horizons}"
    assert horizons[0] == 0, "first horizon must be 0"
    *shape, n_h = values.shape
    shape = tuple(shape)
    assert horizons.shape == (n_h,)
    assert tuple(target_horizons.shape) == shape, f"{tuple(target_horizons.shape)} != {shape}"
    device = dev.require_cuda()
    lib = _lib.load()
    np_like = dev.is_numpy_like(values)
    v = dev.to_device(values, torch.float32, device).reshape(-1, n_h).contiguous()
    t = dev.to_device(target_horizons, torch.float64, device).reshape(-1).contiguous()
    out = torch.empty(v.shape[0], dtype=torch.float64, device=device)
    axis = np.ascontiguousarray(horizons, dtype=np.float64)
    rc = lib.ppo_horizon_interpolate(axis.ctypes.data, n_h, dev.ptr(v), v.shape[0], n_h, dev.ptr(t), dev.ptr(out),
                                     dev.stream_ptr())
    _lib.check(rc, "ppo_horizon_interpolate")
    return dev.like_input(out.reshape(shape), np_like)


def tvf_head_loss_weights(tvf_weights, horizons=None, required_tvf_heads=None, single_value_head=None,
                          head_weighting="off", max_horizon=None, n_step=None):
    """Per-head loss weights of ``rl/tvf.py:51-62`` folded into one float32 vector on the host: the duplicate-head
    weights (``runner.tvf_weights[head_filter]``) times, for ``head_weighting == "h_weighted"`` with every head trained,
    ``(1 + (max_horizon - h) / n_step) * 2 / (min + max)``."""
    head_filter = required_tvf_heads if required_tvf_heads is not None else slice(None, None)
    weights = np.asarray(tvf_weights, dtype=np.float32)[head_filter]
    if head_weighting == "h_weighted" and single_value_head is None:
        weighting = np.asarray([1 + ((max_horizon - h) / n_step) for h in horizons], dtype=np.float32)
        adjustment = 2 / (np.min(weighting) + np.max(weighting))
        weights = (weights * weighting) * np.float32(adjustment)
    return np.ascontiguousarray(weights, dtype=np.float32)


class _TVFValueLoss(torch.autograd.Function):
    """``ppo_tvf_value_loss``: the per-sample loss and d loss / d prediction from one pass; backward is the chain rule
    step ``ppo_scale_rows``."""

    @staticmethod
    def forward(ctx, predictions, targets, head_weights, head_index, mask, coef):
        lib = _lib.load()
        rows, heads = predictions.shape
        loss = torch.empty(rows, dtype=torch.float32, device=predictions.device)
        grad = torch.empty((rows, heads), dtype=torch.float32, device=predictions.device) \
            if predictions.requires_grad else None
        rc = lib.ppo_tvf_value_loss(dev.ptr(targets), targets.stride(0) if rows else 0, dev.ptr(head_index),
                                    dev.ptr(predictions), predictions.stride(0) if rows else 0, predictions.stride(1),
                                    dev.ptr(head_weights), dev.ptr(mask), rows, heads, float(coef), dev.ptr(loss),
                                    dev.ptr(grad), dev.stream_ptr())
        _lib.check(rc, "ppo_tvf_value_loss")
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        grad = ctx.grad
        if grad is None:
            return None, None, None, None, None, None
        lib = _lib.load()
        rows, heads = grad.shape
        out = torch.empty_like(grad)
        g = grad_loss.to(torch.float32).contiguous()
        _lib.check(lib.ppo_scale_rows(dev.ptr(grad), dev.ptr(g), rows, heads, dev.ptr(out), dev.stream_ptr()),
                   "ppo_scale_rows")
        return out, None, None, None, None, None


def tvf_value_loss(targets, predictions, head_weights, coef=1.0, required_tvf_heads=None, mask=None):
    """Per-sample TVF value loss ``[B]`` of ``rl/tvf.py:44-73``:
    ``sqrt(K) * mean_k(coef * 0.5 * (targets[:, heads] - predictions)^2 * head_weights * mask)``.

    ``targets`` ``[B, K_all]`` and ``predictions`` ``[B, K]`` (any strides, e.g. ``model_out["tvf_value"][:, :, 0]``)
    are CUDA tensors; ``head_weights`` ``[K]`` comes from :func:`tvf_head_loss_weights`; ``mask`` is the optional
    horizon-dropout mask ``[B, K]`` (already divided by the keep probability, ``rl/tvf.py:67-69``).  Differentiable
    with respect to ``predictions``."""
    device = dev.require_cuda()
    if not (isinstance(predictions, torch.Tensor) and predictions.is_cuda):
        raise ValueError("tvf_value_loss: predictions must be a CUDA tensor (the model output)")
    if predictions.dim() != 2:
        raise ValueError(f"tvf_value_loss: predictions must be [B, K], got {tuple(predictions.shape)}")
    rows, heads = predictions.shape
    t = dev.to_device(targets, torch.float32, device)
    if t.dim() != 2 or t.shape[0] != rows:
        raise ValueError(f"tvf_value_loss: targets must be [B, K], got {tuple(t.shape)} for B={rows}")
    if t.stride(1) != 1:
        t = t.contiguous()
    index = None
    if required_tvf_heads is not None:
        idx = np.asarray(required_tvf_heads, dtype=np.int64).reshape(-1)
        idx = np.where(idx < 0, idx + t.shape[1], idx)
        if idx.size != heads or np.any((idx < 0) | (idx >= t.shape[1])):
            raise IndexError(f"required_tvf_heads {list(required_tvf_heads)} do not select {heads} of {t.shape[1]} heads")
        index = torch.from_numpy(idx.astype(np.int32)).to(device)
    elif t.shape[1] != heads:
        raise ValueError(f"tvf_value_loss: {t.shape[1]} target heads vs {heads} predicted heads")
    w = dev.to_device(head_weights, torch.float32, device).contiguous()
    if tuple(w.shape) != (heads,):
        raise ValueError(f"tvf_value_loss: head_weights must be [{heads}], got {tuple(w.shape)}")
    m = None
    if mask is not None:
        m = dev.to_device(mask, torch.float32, device).contiguous()
        if tuple(m.shape) != (rows, heads):
            raise ValueError(f"tvf_value_loss: mask must be {(rows, heads)}, got {tuple(m.shape)}")
    p = predictions if predictions.dtype == torch.float32 else predictions.float()
    return _TVFValueLoss.apply(p, t, w, index, m, coef)


class TVFRunnerModule:
    """Return-side mirror of ``rl.tvf.TVFRunnerModule``.

    ``parent`` is the runner (anything with ``ext_rewards``, ``terminals``, ``all_obs``, ``tvf_horizons``,
    ``value_heads``, ``N`` and a ``log`` with ``watch_mean``); ``args`` is the config namespace the reference
    reads globally (``args.tvf.*``, ``args.gamma``, ``args.tvf_return_n_step``).  Buffers have the reference's
    shapes (``rl/tvf.py:27-30``).
    """

    def __init__(self, parent, args):
        self.runner = parent
        self.args = args
        n_steps, n_envs = parent.ext_rewards.shape
        n_vh = len(parent.value_heads)
        n_h = len(parent.tvf_horizons)
        self.tvf_value = np.zeros([n_steps + 1, n_envs, n_h, n_vh], dtype=np.float32)
        self.tvf_untrimmed_value = np.zeros([n_steps + 1, n_envs, n_h, n_vh], dtype=np.float32)
        self.tvf_final_value = np.zeros([n_steps + 1, n_envs], dtype=np.float32)
        self.tvf_returns = np.zeros([n_steps, n_envs, n_h, n_vh], dtype=np.float32)

    def trim_horizons(self, tvf_value_estimates, time, method: str = "timelimit", mode: str = "interpolate"):
        """Pull horizons that extend over the time limit back to the limit (reference: ``rl/tvf.py:91-208``, same
        arguments and return values: trimmed estimates ``[A, K, VH]`` (``[A, K, 1]`` in substitute mode), the value
        estimate used for advantages ``[A]`` and the time till termination ``[A]``; ``(copy, 0, None)`` for
        ``method="off"``).  Indices (trimming point, interpolation slot, random picks) are resolved on the host
        with the reference's NumPy calls, the arithmetic on the estimates runs in ``ppo_tvf_trim_horizons``."""
        args = self.args
        horizons = np.asarray(self.runner.tvf_horizons)
        assert horizons[0] == 0, "First horizon must be zero"
        if method == "off":
            out = np.array(tvf_value_estimates, copy=True) if dev.is_numpy_like(tvf_value_estimates) \
                else tvf_value_estimates.clone()
            out[:, 0, :] = 0
            return out, 0, None
        time = np.asarray(time.cpu() if isinstance(time, torch.Tensor) else time)
        if method == "timelimit":
            time_till_termination = np.maximum(args.env.timeout - time, 0)
        elif method == "est_term":
            est_ep_length = np.percentile(list(self.runner.episode_length_buffer) + list(time),
                                          args.tvf.eta_percentile).astype(int) + args.tvf.eta_buffer
            est_ep_length += args.tvf.eta_minh / 4
            time_till_termination = np.maximum(args.env.timeout - time, 0)
            est_time_till_termination = np.maximum(est_ep_length - time, args.tvf.eta_minh)
            time_till_termination = np.minimum(time_till_termination, est_time_till_termination)
            self.runner.log.watch_mean("*ttt_ep_length", np.percentile(self.runner.episode_length_buffer,
                                                                     args.tvf.eta_percentile).astype(int))
            self.runner.log.watch_mean("*ttt_ep_std", np.std(self.runner.episode_length_buffer))
        else:
            raise ValueError(f"Invalid trimming method {method}")
        self.runner.log.watch_stats("ttt", time_till_termination, display_width=0, history_length=4)

        n_envs, n_h, n_vh = tvf_value_estimates.shape
        modes = {"interpolate": 0, "average": 1, "substitute": 2, "random": 3}
        if mode not in modes:
            raise ValueError(f"Invalid trimming mode {mode}")
        trimmed_ks = np.searchsorted(horizons, time_till_termination).astype(np.int32)
        pre = post = frac = picks = None
        if mode == "interpolate":
            # the slot of horizon_interpolate (rl/tvf.py:527-573) on the log-scaled axis (:143-152)
            axis = np.log10(10 + horizons) - 1
            target = np.clip(np.log10(10 + time_till_termination) - 1, axis[0], axis[-1])
            post_i = np.searchsorted(axis, target, side='left')
            pre_i = np.maximum(post_i - 1, 0)
            dx = axis[post_i] - axis[pre_i]
            dx[dx == 0] = 1.0
            frac = np.ascontiguousarray((target - axis[pre_i]) / dx, dtype=np.float64)
            pre, post = pre_i.astype(np.int32), post_i.astype(np.int32)
        elif mode == "random":
            # same np.random.randint calls in the same order as the reference (:182-188)
            picks = np.tile(np.arange(n_h, dtype=np.int32), (n_envs, 1))
            for a, trimmed_k in zip(range(n_envs), trimmed_ks):
                for k in range(trimmed_k, n_h):
                    picks[a, k] = np.random.randint(trimmed_k, k + 1)

        device = dev.require_cuda()
        lib = _lib.load()
        np_like = dev.is_numpy_like(tvf_value_estimates)
        v = dev.to_device(tvf_value_estimates, torch.float32, device).contiguous()
        up = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(device)  # noqa: E731
        d_tk, d_pre, d_post, d_frac, d_picks = up(trimmed_ks), up(pre), up(post), up(frac), up(picks)
        out = torch.empty((n_envs, n_h, 1 if mode == "substitute" else n_vh), dtype=torch.float32, device=device)
        final = torch.empty(n_envs, dtype=torch.float32, device=device)
        rc = lib.ppo_tvf_trim_horizons(dev.ptr(v), n_envs, n_h, n_vh, modes[mode], dev.ptr(d_tk), dev.ptr(d_pre),
                                       dev.ptr(d_post), dev.ptr(d_frac), dev.ptr(d_picks), float(args.tvf.trim_clip),
                                       dev.ptr(out), dev.ptr(final), dev.stream_ptr())
        _lib.check(rc, "ppo_tvf_trim_horizons")
        new, fin = dev.like_input(out, np_like), dev.like_input(final, np_like)
        if np_like:
            delta = np.abs(new - np.where(np.arange(n_h)[None, :, None] == 0, 0, np.asarray(tvf_value_estimates))
                           [:, :, :new.shape[2]])
            self.runner.log.watch_stats("ved", delta.ravel(), display_width=0, history_length=4)
        return new, fin, time_till_termination

    def calculate_tvf_returns(self, value_head: str = "ext", obs=None, rewards=None, dones=None,
                              tvf_return_mode=None, tvf_return_distribution=None, tvf_n_step=None):
        """(tvf_gamma discounted) return estimates ``[N, A, K]`` for the rollout (reference: ``rl/tvf.py:210-271``)."""
        args = self.args
        obs = obs if obs is not None else self.runner.all_obs
        rewards = rewards if rewards is not None else self.runner.ext_rewards
        dones = dones if dones is not None else self.runner.terminals
        tvf_return_mode = tvf_return_mode or args.tvf.return_mode
        tvf_return_distribution = tvf_return_distribution or args.tvf.return_distribution
        tvf_n_step = tvf_n_step or args.tvf_return_n_step

        n_steps, n_envs, *state_shape = obs[:-1].shape
        assert obs.shape == (n_steps + 1, n_envs, *state_shape)
        assert rewards.shape == (n_steps, n_envs)
        assert dones.shape == (n_steps, n_envs)

        start_time = clock.time()
        values = self.tvf_value[..., self.runner.value_heads.index(value_head)]
        returns = get_return_estimate(
            mode=tvf_return_mode,
            distribution=tvf_return_distribution,
            gamma=args.tvf.gamma,
            rewards=rewards,
            dones=dones,
            required_horizons=np.asarray(self.runner.tvf_horizons),
            value_sample_horizons=np.asarray(self.runner.tvf_horizons),
            value_samples=values,
            n_step=tvf_n_step,
            max_samples=args.tvf.return_samples,
            use_log_interpolation=args.tvf.return_use_log_interpolation,
        )
        self.runner.log.watch_mean("time_return_estimate", clock.time() - start_time,
                                   display_precision=3, display_name="t_re")
        return returns

    def get_tvf_ext_value_estimate(self, new_gamma: float):
        """``[N+1, A]`` value estimate fed to GAE (reference: ``rl/tvf.py:303-360``): slices of the value buffers for
        the TVF gamma, the rediscounted estimate (``ppo_value_estimate``) for any other gamma."""
        args = self.args
        assert args.tvf.enabled
        head = self.runner.value_heads.index('ext')
        trimmed = self.tvf_value[:, :, :, head]
        untrimmed = self.tvf_untrimmed_value[:, :, :, head]
        if abs(new_gamma - args.tvf.gamma) < 1e-8:
            if args.tvf.trimming_mode == "off":
                return untrimmed[:, :, -1]
            if args.tvf.trim_advantages == "trimmed":
                return trimmed[:, :, -1]
            elif args.tvf.trim_advantages == "average":
                return self.tvf_final_value
            elif args.tvf.trim_advantages == "untrimmed":
                return untrimmed[:, :, -1]
            raise ValueError(f"Invalid advantage trimming mode {args.tvf.trim_advantages}.")
        assert args.tvf.trim_advantages != "average", "Average advantage trimming not supported with rediscounting."
        if args.tvf.trim_advantages == "trimmed":
            tvf_values = trimmed
        elif args.tvf.trim_advantages == "untrimmed":
            tvf_values = untrimmed
        else:
            raise ValueError()
        n1, n_envs, n_h = tvf_values.shape
        return get_rediscounted_value_estimate(
            values=tvf_values.reshape([n1 * n_envs, -1]),
            old_gamma=args.tvf.gamma,
            new_gamma=new_gamma,
            horizons=self.runner.tvf_horizons,
        ).reshape([n1, n_envs])

    def rediscount_horizons(self, old_value_estimates):
        """``[B, K] -> [B, K]`` (reference: ``rl/tvf.py:362-380``)."""
        args = self.args
        if args.tvf.gamma == args.gamma:
            return old_value_estimates
        return rediscount_horizons(old_value_estimates, args.tvf.gamma, args.gamma, self.runner.tvf_horizons)

    def on_train_value_minibatch(self, model_out, data, **kwargs):
        """Per-sample TVF value loss ``[B]`` of one minibatch (reference: ``rl/tvf.py:32-77``, same arguments).
        ``data["tvf_returns"]`` ``[B, K]`` and ``model_out["tvf_value"]`` ``[B, K, VH]`` are CUDA tensors (the
        targets never left the device, see :func:`ppo_b200.rollout.gather_rows`)."""
        args = self.args
        assert "tvf_returns" in data, "TVF returns were not uploaded with batch."
        required_tvf_heads = kwargs['required_tvf_heads']
        single_value_head = kwargs['single_value_head']
        predictions = model_out["tvf_value"][:, :, 0]
        weights = tvf_head_loss_weights(self.runner.tvf_weights, self.runner.tvf_horizons, required_tvf_heads,
                                        single_value_head, args.tvf.head_weighting,
                                        getattr(args.tvf, "max_horizon", None), getattr(args, "tvf_return_n_step", None))
        mask = None
        if args.tvf.horizon_dropout > 0:
            keep_prob = (1 - args.tvf.horizon_dropout)
            mask = torch.bernoulli(torch.ones_like(predictions) * keep_prob) / keep_prob
        tvf_loss = tvf_value_loss(data["tvf_returns"], predictions, weights, args.tvf.coef, required_tvf_heads, mask)
        self.runner.log.watch_mean("loss_tvf", tvf_loss.mean(), history_length=64 * args.value_opt.epochs,
                                   display_name="ls_tvf", display_width=8)
        return tvf_loss

    def save(self):
        pass

    def load(self):
        pass
