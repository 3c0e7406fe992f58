"""CPU oracle (NumPy): return trackers and running moments around the hot path.  TEST INFRASTRUCTURE ONLY.

Restates, in the reference's own order of operations:

* ``RunningMeanStd`` / ``update_mean_var_count_from_moments``  (``rl/utils.py:379-396, 414-443``)
* ``Runner.calculate_intrinsic_returns``                      (``rl/rollout.py:1140-1180``)
* ``VecNormalizeRewardWrapper.step`` run over a whole rollout  (``rl/wrappers.py:795-899``), modes "rms" / "ema" / "custom", no episodic
  discounting (``ed_type=None``), identity ``returns_transform``

Pinned against outputs of the reference's own source for these pieces, executed by ``tests/golden/make_golden.py``
(``tests/golden/normalization_vectors.npz``; ``tests/test_oracle_golden.py``).
"""
import math

import numpy as np

from .returns_np import gae


class RunningMeanStd:
    """rl/utils.py:414-443 (scalar shape)."""

    def __init__(self, epsilon=1e-4):
        self.mean = np.float64(0.0)
        self.var = np.float64(1.0)
        self.count = epsilon

    def update(self, x):  # rl/utils.py:425-435
        x = np.asarray(x)
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):  # rl/utils.py:379-396
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m2 = m_a + m_b + np.square(delta) * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, m2 / tot_count, tot_count

    def state(self):
        return np.asarray([self.mean, self.var, self.count], dtype=np.float64)


def intrinsic_returns(int_rewards, int_value, terminals, gamma_int, lambda_policy, ems_norm, rms,
                      normalize=True, center=False, propagation=False):
    """rl/rollout.py:1140-1180.  ``ems_norm`` [A] float64 and ``rms`` are updated in place like the runner's state.
    Returns (int_advantage, int_returns, normalised int_rewards, scale)."""
    int_rewards = np.asarray(int_rewards)
    steps = int_rewards.shape[0]
    scale = 1.0
    if normalize:
        for t in range(steps):  # :1150-1153
            term = (not propagation) * terminals[t, :]
            ems_norm[:] = (1 - term) * gamma_int * ems_norm + int_rewards[t, :]
            rms.update(ems_norm.reshape(-1))
        scale = 1e-5 + rms.var ** 0.5  # :1164
        int_rewards = int_rewards / scale  # :1165
    if center:
        int_rewards = int_rewards - int_rewards.mean()  # :1167-1168
    adv = gae(int_rewards, int_value[:steps], int_value[steps], (not propagation) * terminals,
              gamma=gamma_int, lamb=lambda_policy)  # :1170-1177
    return adv, adv + int_value[:steps], int_rewards, scale


def episodic_normalization_constant(k, discount_type, gamma=1.0, discount_bias=1.0):
    """``EpisodicDiscounting.get_normalization_constant`` (rl/wrappers.py:52-72).  The reward wrapper calls it with the
    default ``gamma=1.0`` (:872), for which "finite" raises ZeroDivisionError (plain Python, :55) and "geometric" gives
    infinite normalisers (an array divided by zero, :60) -- there too."""
    k = k + discount_bias  # :53
    if discount_type == "finite":
        m = 1 / (1 - gamma)  # :55
        return np.clip(m - k, 0, float("inf")) + 1  # :56-58
    if discount_type == "geometric":
        return (gamma ** k) / (1 - gamma)  # :60
    if discount_type == "quadratic":
        return 1 / k  # :62
    if discount_type == "power":
        return (1 / 1e-1) * (k ** -1e-1)  # :64-65
    if discount_type == "harmonic":
        return 1 / np.log(k)  # :67
    if discount_type == "none":
        return 1.0  # :69
    raise ValueError(f"Invalid discount_type {discount_type}")  # :71


def return_tracker_update(current_returns, rms, rewards, dones, mask, gamma):
    """``ReturnTracker.update`` (rl/wrappers.py:749-755): only the masked envs move, ``current_returns`` keeps its dtype
    (float32 in the reference, :743), the running moments take the masked envs' batch."""
    mask = np.asarray(mask, dtype=bool)
    if sum(mask) == 0:  # :750-751
        return
    current_returns[mask] = rewards[mask] + gamma * current_returns[mask] * (1 - dones[mask])  # :754
    rms.update(current_returns[mask])  # :755


def vec_normalize_rewards(rewards, dones, gamma, current_returns, rms, clip=10.0, scale=1.0, epsilon=1e-2, mode="rms",
                          ret_var=0.0, ema_horizon=5e6, ed_type=None, ed_bias=1.0, returns_transform=lambda x: x,
                          times=None):
    """rl/wrappers.py:825-885 applied to each row of a [N, A] rollout in order.  ``current_returns`` [A] and ``rms``
    are updated in place.  Returns (scaled rewards [N, A], number of clipped elements) for mode "rms" and, for the
    modes "ema" / "custom" (:878-881, :893-899), additionally the final ``ret_var``.  ``ed_type`` / ``times`` [N, A]:
    the episodic-discounting normalisers of :870-876 (``times`` are the envs' ``info['time']``)."""
    rewards = np.asarray(rewards)
    out = np.zeros(rewards.shape, dtype=np.float64)
    clips = 0
    for t in range(rewards.shape[0]):
        current_returns[:] = rewards[t] + gamma * current_returns * (1 - dones[t])  # :847
        norms = 1 if ed_type is None else \
            episodic_normalization_constant(np.asarray(times[t]), ed_type, discount_bias=ed_bias)  # :870-874
        rms.update(returns_transform(current_returns / norms))  # :876
        if mode == "ema":  # :878-881
            alpha = 1 - (len(dones[t]) / min(rms.count, ema_horizon))
            ret_var = alpha * ret_var + (1 - alpha) * np.var(current_returns)
        if mode == "rms":
            std = math.sqrt(rms.var + epsilon)  # :893-895
        elif mode in ("ema", "custom"):
            std = math.sqrt(ret_var + epsilon)  # :896-897
        else:
            raise ValueError(f"Invalid mode {mode}")  # :899
        scaled = rewards[t] / std  # :863
        if clip is not None and clip >= 0:  # :866-873
            clipped = np.clip(scaled, -clip, +clip)
            clips += int(np.sum(clipped != scaled))
            scaled = clipped
        out[t] = scaled * scale  # :875
    if mode == "rms":
        return out, clips
    return out, clips, ret_var
