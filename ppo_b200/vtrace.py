"""Drop-in for the reference's ``rl/vtrace.py`` (same names and signatures), computed by the sm_100a
kernels behind ``ppo_vtrace`` / ``ppo_vtrace_trust_region``.

All arithmetic is float32 like the reference's float32 path (``rl/vtrace.py:78``); float64 inputs (the
reference's stale unit test builds them from Python lists) are rounded to float32 on the way in.
"""
import numpy as np
import torch

from . import _device as dev
from . import _lib


def v_trace_trust_region(behaviour_log_policy, target_log_policy):
    """Trust-region weights ``[N, B]`` in [0, 1] (reference: ``rl/vtrace.py:3-39``)."""
    device = dev.require_cuda()
    lib = _lib.load()
    beh = dev.to_device(behaviour_log_policy, torch.float32, device).contiguous()
    tgt = dev.to_device(target_log_policy, torch.float32, device).contiguous()
    if beh.dim() != 3 or beh.shape != tgt.shape:
        raise ValueError(f"log policies must both be [N, B, A], got {tuple(beh.shape)} and {tuple(tgt.shape)}")
    steps, batch, n_act = beh.shape
    out = torch.empty((steps, batch), dtype=torch.float32, device=device)
    rc = lib.ppo_vtrace_trust_region(dev.ptr(beh), dev.ptr(tgt), steps * batch, n_act, dev.ptr(out), dev.stream_ptr())
    _lib.check(rc, "ppo_vtrace_trust_region")
    return dev.like_input(out, dev.is_numpy_like(behaviour_log_policy))


def importance_sampling_v_trace(behaviour_log_policy, target_log_policy, actions, rewards, dones,
                                target_value_estimates, target_value_final_estimate, gamma, lamb=1.0, check=True):
    """V-trace targets ``(vs, weighted_advantages, cs)``, each ``[N, B]`` (reference: ``rl/vtrace.py:44-114``).

    An action outside ``[-A, A)`` raises ``IndexError`` like the reference's fancy index (``rl/vtrace.py:90``): host
    arrays are checked before the upload, device tensors by the kernel itself (it clamps the index and raises a flag
    word that is read back after the launch -- one 4-byte copy and a stream synchronisation; ``check=False`` skips the
    read for callers that must stay asynchronous, the clamp stays)."""
    device = dev.require_cuda()
    lib = _lib.load()
    np_like = dev.is_numpy_like(behaviour_log_policy)
    beh = dev.to_device(behaviour_log_policy, torch.float32, device).contiguous()
    tgt = dev.to_device(target_log_policy, torch.float32, device).contiguous()
    if beh.dim() != 3 or beh.shape != tgt.shape:
        raise ValueError(f"log policies must both be [N, B, A], got {tuple(beh.shape)} and {tuple(tgt.shape)}")
    steps, batch, n_act = beh.shape
    act_on_host = dev.is_numpy_like(actions)
    if act_on_host:
        a_host = np.asarray(actions)
        if a_host.size and (a_host.min() < -n_act or a_host.max() >= n_act):  # rl/vtrace.py:90 would raise IndexError
            raise IndexError(f"index {int(a_host.max())} is out of bounds for axis 2 with size {n_act}")
    act = dev.to_device(actions, None, device)
    if act.dtype not in (torch.int64, torch.int32):
        act = act.to(torch.int64)
    act = act.contiguous()
    r = dev.to_device(rewards, torch.float32, device).contiguous()
    v = dev.to_device(target_value_estimates, torch.float32, device).contiguous()
    fin = dev.to_device(target_value_final_estimate, torch.float32, device).contiguous()
    d, kind = dev.done_tensor(dones, device)
    for name, t in (("actions", act), ("rewards", r), ("dones", d), ("target_value_estimates", v)):
        if tuple(t.shape) != (steps, batch):
            raise ValueError(f"{name} must be [N, B] = {(steps, batch)}, got {tuple(t.shape)}")
    if tuple(fin.shape) != (batch,):
        raise ValueError(f"target_value_final_estimate must be [B] = {(batch,)}, got {tuple(fin.shape)}")
    d = d.contiguous()
    vs = torch.empty((steps, batch), dtype=torch.float32, device=device)
    adv = torch.empty_like(vs)
    cs = torch.empty_like(vs)
    bad = torch.zeros(1, dtype=torch.int32, device=device) if check and not act_on_host else None
    rc = lib.ppo_vtrace(dev.ptr(beh), dev.ptr(tgt), dev.ptr(act), act.element_size(), dev.ptr(r), dev.ptr(d), kind,
                        dev.ptr(v), dev.ptr(fin), steps, batch, n_act, float(gamma), float(lamb),
                        dev.ptr(vs), dev.ptr(adv), dev.ptr(cs), dev.ptr(bad), dev.stream_ptr())
    _lib.check(rc, "ppo_vtrace")
    if bad is not None and int(bad.item()):
        raise IndexError(f"an action index is out of bounds for axis 2 with size {n_act}")
    return dev.like_input(vs, np_like), dev.like_input(adv, np_like), dev.like_input(cs, np_like)
