"""Drop-in for the reference's ``rl/returns.py`` (same names, signatures and error behaviour),
computed by the sm_100a scan kernels behind the C ABI (``ppo_gae``, ``ppo_discounted_returns``).

NumPy arrays in -> NumPy arrays out (host<->device copies inside the call, like a user of the
reference would see); CUDA tensors in -> CUDA tensors out with no host traffic.
"""
import numpy as np
import torch

from . import _device as dev
from . import _lib


def _gae_launch(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma, lamb_adv, lamb_ret,
                want_adv, want_ret, want_stats, normalize=None):
    """``normalize=(eps, clip)``: also the batch-normalised advantages (``ppo_gae_normalized``, one launch when the
    grid is resident); returns ``(adv, ret, norm, stats)`` then."""
    device = dev.require_cuda()
    lib = _lib.load()
    rewards = dev.env_contiguous(dev.to_device(batch_rewards, torch.float32, device))
    if rewards.dim() != 2:
        raise ValueError(f"not enough values to unpack: rewards must be [N, A], got {tuple(rewards.shape)}")
    steps, envs = rewards.shape
    values = dev.to_device(batch_value, torch.float32, device)
    final = dev.to_device(final_value_estimate, torch.float32, device)
    if tuple(values.shape) != (steps, envs):
        raise ValueError(f"operands could not be broadcast together: values {tuple(values.shape)} vs rewards {(steps, envs)}")
    if final.dim() == 0:
        final = final.expand(envs)  # rl/unit_tests.py:206 passes a 0-d final estimate
    if tuple(final.shape) != (envs,):
        raise ValueError(f"operands could not be broadcast together: final value {tuple(final.shape)} vs envs {envs}")
    dones, kind = dev.done_tensor(batch_terminal, device)
    if dones is not None:
        if tuple(dones.shape) != (steps, envs):
            raise ValueError(f"operands could not be broadcast together: terminals {tuple(dones.shape)} vs {(steps, envs)}")
        dones = dev.env_contiguous(dones)
    adv = torch.empty((steps, envs), dtype=torch.float32, device=device) if want_adv else None
    ret = torch.empty((steps, envs), dtype=torch.float32, device=device) if want_ret else None
    stats = torch.empty(_lib.STATS_LEN, dtype=torch.float64, device=device) if want_stats else None
    ws = dev.workspace(lib.ppo_gae_workspace_bytes(envs), device) if want_stats else None
    if normalize is not None:
        norm = torch.empty(steps * envs, dtype=torch.float32, device=device)
        rc = lib.ppo_gae_normalized(
            dev.ptr(rewards), rewards.stride(0) if steps else 0,
            dev.ptr(values), values.stride(0) if steps else 0, values.stride(1) if envs else 0,
            dev.ptr(final), final.stride(0) if envs else 0,
            dev.ptr(dones), kind, dones.stride(0) if dones is not None and steps else 0,
            steps, envs, float(gamma), float(lamb_adv), float(lamb_ret), float(normalize[0]), float(normalize[1]),
            dev.ptr(adv), dev.ptr(ret), dev.ptr(norm), dev.ptr(stats),
            dev.ptr(ws), ws.numel() if ws is not None else 0, dev.stream_ptr())
        _lib.check(rc, "ppo_gae_normalized")
        return adv, ret, norm, stats
    rc = lib.ppo_gae(
        dev.ptr(rewards), rewards.stride(0) if steps else 0,
        dev.ptr(values), values.stride(0) if steps else 0, values.stride(1) if envs else 0,
        dev.ptr(final), final.stride(0) if envs else 0,
        dev.ptr(dones), kind, dones.stride(0) if dones is not None and steps else 0,
        steps, envs, float(gamma), float(lamb_adv), float(lamb_ret),
        dev.ptr(adv), dev.ptr(ret), dev.ptr(stats),
        dev.ptr(ws), ws.numel() if ws is not None else 0, dev.stream_ptr())
    _lib.check(rc, "ppo_gae")
    return adv, ret, stats


def gae(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma: float, lamb=0.95):
    """Calculates GAE based on rollout data (reference: ``rl/returns.py:7-29``)."""
    adv, _, _ = _gae_launch(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma, lamb, lamb,
                            True, False, False)
    return dev.like_input(adv, dev.is_numpy_like(batch_rewards))


def td_lambda(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma: float, lamb=0.95):
    """TD(lambda) returns = GAE + value (reference: ``rl/returns.py:58-67``)."""
    _, ret, _ = _gae_launch(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma, lamb, lamb,
                            False, True, False)
    return dev.like_input(ret, dev.is_numpy_like(batch_rewards))


def gae_td_lambda(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma: float,
                  lamb_policy=0.95, lamb_value=0.95, with_stats=False):
    """Fused form of the two calls the trainer makes back to back (``rl/rollout.py:1207-1223``):
    one read of the rollout, advantages with ``lamb_policy`` and TD(lambda) returns with ``lamb_value``.
    With ``with_stats`` also returns the device statistics record (count / sum / sum of squares / min / max
    of both outputs) that feeds :func:`ppo_b200.rollout.normalize_advantages`."""
    adv, ret, stats = _gae_launch(batch_rewards, batch_value, final_value_estimate, batch_terminal, gamma,
                                  lamb_policy, lamb_value, True, True, with_stats)
    np_like = dev.is_numpy_like(batch_rewards)
    out = (dev.like_input(adv, np_like), dev.like_input(ret, np_like))
    return out + (stats,) if with_stats else out


def calculate_bootstrapped_returns(rewards, dones, final_value_estimate, gamma) -> np.ndarray:
    """Discounted returns bootstrapped from a final value estimate (reference: ``rl/returns.py:32-55``).

    ``gamma`` is a python float or an ``[N, A]`` array, exactly as in the reference."""
    device = dev.require_cuda()
    lib = _lib.load()
    r = dev.env_contiguous(dev.to_device(rewards, torch.float32, device))
    steps, envs = r.shape
    d, kind = dev.done_tensor(dones, device)
    if d is None or tuple(d.shape) != (steps, envs):
        raise ValueError(f"dones must have shape {(steps, envs)}")
    d = dev.env_contiguous(d)
    final = dev.to_device(final_value_estimate, torch.float32, device)
    if final.dim() == 0:
        final = final.expand(envs)
    if tuple(final.shape) != (envs,):
        raise ValueError(f"final_value_estimate must have shape {(envs,)}")
    garr = None
    if type(gamma) is not float:  # rl/returns.py:49-50: anything but a python float is used as an [N, A] array
        garr = dev.env_contiguous(dev.to_device(gamma, torch.float32, device))
        if garr.dim() == 0:
            raise TypeError("'float' gamma expected or an [N, A] array (rl/returns.py:49-53 indexes gamma[i])")
        if tuple(garr.shape) != (steps, envs):
            garr = garr.expand(steps, envs).contiguous()
    out = torch.empty((steps, envs), dtype=torch.float32, device=device)
    rc = lib.ppo_discounted_returns(
        dev.ptr(r), r.stride(0) if steps else 0, dev.ptr(d), kind, d.stride(0) if steps else 0,
        dev.ptr(final), final.stride(0) if envs else 0,
        float(gamma) if garr is None else 0.0, dev.ptr(garr), garr.stride(0) if garr is not None and steps else 0,
        steps, envs, dev.ptr(out), dev.stream_ptr())
    _lib.check(rc, "ppo_discounted_returns")
    return dev.like_input(out, dev.is_numpy_like(rewards))
