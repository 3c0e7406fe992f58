"""Oracle (test infrastructure): GAE / TD(lambda) / bootstrapped returns / advantage normalisation.

NumPy restatement of ``rl/returns.py`` and the inline normalisation at
``rl/rollout.py:1887-1897`` of the reference.  The arithmetic below keeps the
reference's *promotion behaviour* on purpose, because parity is judged against
it:

* with boolean terminals ``1.0 - terminal`` is float64, so the carried value
  is float64 and only the stored row is rounded to float32
  (``rl/returns.py:26-28``: the chained assignment keeps the un-rounded RHS);
* ``gamma * next_value`` is a float32 product (python scalars are weak);
* ``calculate_bootstrapped_returns`` first turns a python-float gamma into a
  float32 array (``rl/returns.py:49-50``), i.e. it discounts with
  ``float32(gamma)``, whereas ``gae`` uses the float64 ``gamma * lamb``.
"""
import numpy as np


def gae(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma: float, lamb=0.95):
    """Reverse-time GAE scan over ``[N, A]``; follows ``rl/returns.py:7-29``."""
    steps, _envs = batch_rewards.shape
    out = np.zeros_like(batch_rewards, dtype=np.float32)
    carry = np.zeros(batch_rewards.shape[1:], dtype=np.float32)
    t = steps
    while t > 0:
        t -= 1
        # rl/returns.py:23-24 -- no terminal array means "never terminal"
        ended = batch_terminal[t] if batch_terminal is not None else False
        alive = 1.0 - ended
        # rl/returns.py:25 -- bootstrap from the next row, or the final estimate on the last row
        v_next = final_value_estimate if t == steps - 1 else batch_value[t + 1]
        # rl/returns.py:26
        td_error = batch_rewards[t] + gamma * v_next * alive - batch_value[t]
        # rl/returns.py:27-28 -- carry keeps the un-rounded value, the stored row is float32
        carry = td_error + gamma * lamb * alive * carry
        out[t] = carry
    return out


def td_lambda(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma: float, lamb=0.95):
    """TD(lambda) returns = GAE + value; follows ``rl/returns.py:58-67``."""
    return gae(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma, lamb) + batch_value


def calculate_bootstrapped_returns(rewards, dones, final_value_estimate, gamma):
    """Discounted returns bootstrapped from a final value; follows ``rl/returns.py:32-55``.

    ``gamma`` is a python float or an ``[N, A]`` array (only ``type(gamma) is float``
    is broadcast, ``rl/returns.py:49-50``).
    """
    steps, envs = rewards.shape
    out = np.zeros([steps, envs], dtype=np.float32)
    if type(gamma) is float:
        gamma = np.ones([steps, envs], dtype=np.float32) * gamma
    running = final_value_estimate
    for t in range(steps - 1, -1, -1):
        # rl/returns.py:53
        running = rewards[t] + running * gamma[t] * (1.0 - dones[t])
        out[t] = running
    return out


def normalize_advantages(advantage, epsilon=1e-8, clipping=None):
    """Batch-level advantage normalisation; follows ``rl/rollout.py:1887-1897``.

    Population std (ddof=0) over the flattened batch, optional symmetric clip.
    Returns the flattened float32 vector exactly like the reference.
    """
    flat = advantage.reshape(-1).copy()
    flat = (flat - flat.mean()) / (flat.std() + epsilon)
    if clipping is not None:
        flat = np.clip(flat, -clipping, +clipping)
    return flat
