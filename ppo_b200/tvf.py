"""Drop-in for the return-side pieces of the reference's ``rl/tvf.py``: the horizon table generator and the
``TVFRunnerModule`` methods that feed / call the hot path (``calculate_tvf_returns`` :210-271,
``get_tvf_ext_value_estimate`` :303-360).  The loss, trimming and logging parts of the module are not on
the return path and are out of scope.
"""
import collections
import math
import time as clock

import numpy as np

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
        """``[N+1, A]`` value estimate fed to GAE (reference: ``rl/tvf.py:303-360``).  Same-gamma cases are
        slices of the value buffers; rediscounting to a different gamma is a "next" row and
        raises until that kernel lands."""
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
        raise NotImplementedError("rediscounting to a different gamma (rl/tvf.py:355-360) is not built yet")

    def save(self):
        pass

    def load(self):
        pass
