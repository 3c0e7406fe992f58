"""Oracle (test infrastructure): V-trace targets and the trust-region weighting.

NumPy restatement of ``rl/vtrace.py`` of the reference.  rho-bar = c-bar = 1 are
hard-coded there (``rl/vtrace.py:93-94``); all arithmetic is float32 when the
inputs are float32 because the not-terminal mask is cast to float32 (:78).
"""
import numpy as np


def v_trace_trust_region(behaviour_log_policy, target_log_policy):
    """``1 / (1 + KL(pi || pi_hat))`` per (t, b); follows ``rl/vtrace.py:3-39``.

    ``pi_hat = min(mu, pi)`` renormalised over the action axis.
    """
    mu = np.exp(behaviour_log_policy)
    pi = np.exp(target_log_policy)
    implied = np.minimum(1 * mu, pi)  # :31
    implied /= np.sum(implied, axis=2, keepdims=True)  # :34
    kl = (np.log(pi / implied) * pi).sum(axis=2)  # :37
    return 1 / (1 + kl)


def importance_sampling_v_trace(behaviour_log_policy, target_log_policy, actions, rewards, dones,
                                target_value_estimates, target_value_final_estimate, gamma, lamb=1.0):
    """V-trace value targets, rho-weighted advantages and ``c`` weights; follows ``rl/vtrace.py:44-114``."""
    steps, batch, _n_act = behaviour_log_policy.shape
    alive = np.asarray(1 - dones, dtype=np.float32)  # :78
    # :89-91 -- log-probabilities of the taken actions (output dtype follows the value estimates, :80-81)
    t_idx = np.arange(steps)[:, None]
    b_idx = np.arange(batch)[None, :]
    tgt = np.zeros_like(target_value_estimates)
    beh = np.zeros_like(target_value_estimates)
    tgt[...] = target_log_policy[t_idx, b_idx, actions]
    beh[...] = behaviour_log_policy[t_idx, b_idx, actions]
    rhos = np.minimum(1, np.exp(tgt - beh))  # :93
    cs = lamb * rhos  # :94
    acc = np.zeros_like(target_value_estimates)
    for t in range(steps - 1, -1, -1):  # :96-102
        v_next = target_value_estimates[t + 1] if t + 1 != steps else target_value_final_estimate
        td = rhos[t] * (rewards[t] + (gamma * alive[t]) * v_next - target_value_estimates[t])
        tail = acc[t + 1] if t + 1 != steps else 0
        acc[t] = td + (gamma * alive[t]) * cs[t] * tail
    vs = acc + target_value_estimates  # :105
    vs_next = np.concatenate((vs[1:], [target_value_final_estimate]))  # :106
    adv = rhos * (rewards + gamma * alive * vs_next - target_value_estimates)  # :109
    return vs, adv.astype(dtype=np.float32), cs  # :112-114
