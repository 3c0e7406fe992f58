"""Seeded synthetic rollouts shaped like the reference's own test data.

Distributions follow ``tests/test_tvf.py:17-24`` of the reference: rewards are
integers in {-1, 0, 1, 2} stored as float32, dones are Bernoulli(3/101) booleans,
value samples are N(0.1, 0.4) float32 with the horizon-0 column zeroed.  Used by
the tests, ``bench.py`` and ``__graft_entry__.smoke()``; pure NumPy, no GPU.
"""
import numpy as np


def rollout(steps, envs, seed=0, done_prob=3 / 101):
    """rewards [N,A] f32, dones [N,A] bool, values [N+1,A] f32 (row N is the final estimate)."""
    rng = np.random.RandomState(seed)
    rewards = rng.randint(-1, 3, size=[steps, envs]).astype(np.float32)
    dones = rng.random_sample([steps, envs]) < done_prob
    values = rng.normal(0.1, 0.4, size=[steps + 1, envs]).astype(np.float32)
    return rewards, dones, values


def value_samples(steps, envs, n_horizons, seed=0):
    """[N+1, A, V] f32 value curve samples with the first (h=0) column zeroed."""
    rng = np.random.RandomState(seed + 7919)
    vs = rng.normal(0.1, 0.4, size=[steps + 1, envs, n_horizons]).astype(np.float32)
    vs[:, :, 0] = 0
    return vs


def policies(steps, envs, n_actions=18, seed=0):
    """behaviour/target log-policies [N,B,ACT] f32 (log of Dirichlet(1) rows) and int64 actions [N,B]."""
    rng = np.random.RandomState(seed + 104729)
    beh = np.log(rng.dirichlet(np.ones(n_actions), size=[steps, envs])).astype(np.float32)
    tgt = np.log(rng.dirichlet(np.ones(n_actions), size=[steps, envs])).astype(np.float32)
    actions = rng.randint(0, n_actions, size=[steps, envs]).astype(np.int64)
    return beh, tgt, actions
