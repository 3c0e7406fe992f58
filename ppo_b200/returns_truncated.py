"""Drop-in for the reference's ``rl/returns_truncated.py``: same dispatcher signature, same NumPy RNG
calls for the n-step sample table, targets computed by the sm_100a kernels behind ``ppo_tvf_targets``.

Reference semantics kept: NumPy in -> NumPy ``[N, A, K]`` float32 out; ``ValueError`` for an unknown
distribution / mode (``rl/returns_truncated.py:78,139``); ``AssertionError`` for an n-step outside
``1..N`` (:597); ``IndexError`` when a bootstrap horizon lies beyond the last value horizon (:157).
"""
import ctypes
import os

import numpy as np
import torch

from . import _device as dev
from . import _lib
from . import _tvf_tables

PATH_AUTO, PATH_GENERIC, PATH_RING, PATH_TSLAB, PATH_CSLAB, PATH_QSLAB = 0, 1, 2, 3, 4, 5


class _CTables(ctypes.Structure):
    """``ppo_tvf_tables`` of include/ppo_b200.h."""
    _fields_ = [("n_entries", ctypes.c_int32), ("nmax", ctypes.c_int32),
                ("ent_n", ctypes.c_void_p), ("ent_cnt", ctypes.c_void_p), ("ent_idx", ctypes.c_void_p),
                ("ent_wlo", ctypes.c_void_p), ("ent_whi", ctypes.c_void_p), ("ent_g", ctypes.c_void_p),
                ("tail_idx", ctypes.c_void_p), ("tail_wlo", ctypes.c_void_p), ("tail_whi", ctypes.c_void_p),
                ("tail_cwlo", ctypes.c_void_p), ("tail_cwhi", ctypes.c_void_p),
                ("gpow", ctypes.c_void_p), ("inv_c", ctypes.c_float), ("tail_pitch", ctypes.c_int32),
                ("tailT_idx", ctypes.c_void_p), ("tailT_cwlo", ctypes.c_void_p), ("tailT_cwhi", ctypes.c_void_p),
                ("host_blob", ctypes.c_void_p), ("dev_blob", ctypes.c_void_p), ("gamma", ctypes.c_double),
                ("tslab_offset_words", ctypes.c_int64), ("cslab_offset_words", ctypes.c_int64)]


class _CLayout(ctypes.Structure):
    """``ppo_tvf_table_layout`` of include/ppo_b200.h."""
    _fields_ = [("n_entries", ctypes.c_int32), ("nmax", ctypes.c_int32), ("total_words", ctypes.c_int64),
                ("offset_words", ctypes.c_int64 * 16), ("inv_c", ctypes.c_float), ("tail_pitch", ctypes.c_int32),
                ("gamma", ctypes.c_double), ("cslab_offset_words", ctypes.c_int64),
                ("sections", ctypes.c_int32), ("reserved", ctypes.c_int32)]


TABLES_GENERIC, TABLES_TSLAB, TABLES_CSLAB, TABLES_ALL = 1, 2, 4, 7  # PPO_TVF_TABLES_* of include/ppo_b200.h


class _CExtras(ctypes.Structure):
    """``ppo_tvf_extras`` of include/ppo_b200.h."""
    _fields_ = [("value_weights", ctypes.c_void_p), ("value_out", ctypes.c_void_p),
                ("value_out_row_stride", ctypes.c_int64),
                ("gae_enable", ctypes.c_int32), ("gae_done", ctypes.c_int32),
                ("gae_gamma", ctypes.c_double), ("gae_lambda_adv", ctypes.c_double), ("gae_lambda_ret", ctypes.c_double),
                ("gae_adv_out", ctypes.c_void_p), ("gae_ret_out", ctypes.c_void_p), ("gae_stats", ctypes.c_void_p),
                ("gae_workspace", ctypes.c_void_p), ("gae_workspace_bytes", ctypes.c_int64)]


def _extras(value_out, value_weights, env_offset=0, gae=None):
    """ctypes extras for ``ppo_tvf_targets_ex`` (or None): ``value_out`` is a device ``[N+1, A]`` float32 tensor that
    receives the value estimate of the env block starting at ``env_offset``; ``value_weights`` a float32 NumPy
    vector ``[V]`` (rediscount weights, ``tvf.rediscount_weights``) or None for the last value horizon.
    ``gae``: optional dict(gamma, lambda_adv, lambda_ret, adv, ret, stats, workspace) of device tensors -- GAE and
    TD(lambda) over the value estimate in the same launch (check ``extras.gae_done`` after the call)."""
    if value_out is None:
        return None, None
    x = _CExtras()
    if gae is not None:
        x.gae_enable = 1
        x.gae_gamma, x.gae_lambda_adv, x.gae_lambda_ret = float(gae["gamma"]), float(gae["lambda_adv"]), float(gae["lambda_ret"])
        x.gae_adv_out, x.gae_ret_out = gae["adv"].data_ptr(), gae["ret"].data_ptr()
        x.gae_stats, x.gae_workspace = gae["stats"].data_ptr(), gae["workspace"].data_ptr()
        x.gae_workspace_bytes = gae["workspace"].numel()
    keep = None
    if value_weights is not None:
        keep = np.ascontiguousarray(value_weights, dtype=np.float32)
        x.value_weights = keep.ctypes.data
    x.value_out = value_out.data_ptr() + 4 * env_offset
    x.value_out_row_stride = value_out.stride(0)
    return x, keep


_TABLE_NAMES = ["ent_n", "ent_cnt", "ent_idx", "ent_wlo", "ent_whi", "ent_g", "tail_idx", "tail_wlo", "tail_whi",
                "tail_cwlo", "tail_cwhi", "gpow", "tailT_idx", "tailT_cwlo", "tailT_cwhi"]


class NativeTables:
    """Packed index tables built by ``ppo_tvf_tables_build`` (host integers and table arithmetic only):
    ``host`` is the blob (a pinned CPU int32 tensor when CUDA is present), ``layout`` its table offsets."""
    __slots__ = ("host", "layout", "n_entries", "nmax", "inv_c", "n_req", "steps")

    def table(self, name):
        """NumPy view of one table inside the host blob (for tests)."""
        i = _TABLE_NAMES.index(name)
        lo, hi = self.layout.offset_words[i], self.layout.offset_words[i + 1]
        words = self.host.numpy()[lo:hi]
        return words.view(np.float32) if name not in ("ent_n", "ent_idx", "tail_idx", "tailT_idx") else words


def build_tables(gamma, steps, required_horizons, value_sample_horizons, n_step_samples,
                 use_log_interpolation=False, host=None, sections=TABLES_ALL):
    """Resolve a ``[K, C]`` n-step sample table into the packed index tables (host side of ``ppo_tvf_targets``).
    Raises like the reference: ``AssertionError`` for an n-step outside ``1..N`` (rl/returns_truncated.py:597),
    ``IndexError`` for a bootstrap horizon beyond the last value horizon (:157).
    ``host``: an int32 CPU tensor to build into (reused when large enough; see :class:`TableRing`).
    ``sections``: ``TABLES_*`` bits (``ppo_tvf_tables_build_ex``): ``TABLES_TSLAB`` alone writes only what the transposed
    slab kernel reads -- when that kernel can serve the table; the builder adds the generic tables itself otherwise."""
    lib = _lib.load()
    req = required_horizons if _is_i64_vector(required_horizons) else \
        np.ascontiguousarray(np.asarray(required_horizons).reshape(-1), dtype=np.int64)
    vsh = value_sample_horizons if _is_i64_vector(value_sample_horizons) else \
        np.ascontiguousarray(np.asarray(value_sample_horizons).reshape(-1), dtype=np.int64)
    samples = np.asarray(n_step_samples)
    assert samples.ndim == 2 and samples.shape[0] == len(req), \
        f"n_step_samples must be [K, C] with K={len(req)}, got {samples.shape}"
    assert samples.shape[1] >= 1, "at least one n-step sample per horizon is required"
    samples = np.ascontiguousarray(samples, dtype=np.int64)
    words = int(lib.ppo_tvf_tables_words(len(req), samples.shape[1], int(steps)))
    if host is None or host.numel() < words:
        host = torch.empty(max(words, 16), dtype=torch.int32, pin_memory=torch.cuda.is_available())
    layout = _CLayout()
    rc = lib.ppo_tvf_tables_build_ex(float(gamma), int(steps), req.ctypes.data, len(req), vsh.ctypes.data, len(vsh),
                                     samples.ctypes.data, samples.shape[1], int(bool(use_log_interpolation)),
                                     int(sections), host.data_ptr(), host.numel(), ctypes.byref(layout))
    if rc == _lib.PPO_ERR_BAD_ARG:
        raise AssertionError(lib.ppo_last_error().decode("utf-8", "replace"))
    if rc == _lib.PPO_ERR_UNSUPPORTED:
        raise IndexError(lib.ppo_last_error().decode("utf-8", "replace"))
    _lib.check(rc, "ppo_tvf_tables_build")
    t = NativeTables()
    t.host, t.layout = host, layout
    t.n_entries, t.nmax, t.inv_c = layout.n_entries, layout.nmax, layout.inv_c
    t.n_req, t.steps = len(req), int(steps)
    return t


def _is_i64_vector(a):
    return isinstance(a, np.ndarray) and a.dtype == np.int64 and a.ndim == 1 and a.flags.c_contiguous


class TableRing:
    """Rotating (pinned host blob, device blob) pairs for a pass that rebuilds its tables every rollout
    (``rl/returns_truncated.py:85-129`` draws a new sample table per call): no allocation per rebuild, and the upload
    is one asynchronous copy on a side stream, so it overlaps the kernels of the previous pass instead of sitting
    between two passes on the compute stream.  Ordering, all with events and no host waits in the steady state:

    * the compute stream waits for the copy before the launch that reads the blob (:meth:`before_launch`);
    * the copy into a device blob waits for the last launch that read that blob (:meth:`after_launch` records it);
    * a host blob is rewritten only after the copy that read it has finished (three slots: over long before the slot
      comes round again)."""

    def __init__(self, device, slots=3):
        self.device = device
        self.slots = [dict(host=None, dev=None, copied=None, used=None) for _ in range(slots)]
        self.i = 0
        self.current = None
        self._pending = False
        self.copy_stream = torch.cuda.Stream(device=device)

    def build(self, gamma, steps, req, vsh, n_step_samples, use_log, sections=TABLES_ALL):
        slot = self.slots[self.i]
        self.i = (self.i + 1) % len(self.slots)
        if slot["copied"] is not None:
            slot["copied"].synchronize()
        tabs = build_tables(gamma, steps, req, vsh, n_step_samples, use_log, host=slot["host"], sections=sections)
        slot["host"] = tabs.host
        if slot["dev"] is None or slot["dev"].numel() < tabs.host.numel():
            slot["dev"] = torch.empty(tabs.host.numel(), dtype=torch.int32, device=self.device)
            slot["used"] = None
        with torch.cuda.stream(self.copy_stream):
            if slot["used"] is not None:
                self.copy_stream.wait_event(slot["used"])
            else:  # first use of this device blob: it was allocated on the compute stream
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            ctab, keep = _upload_tables(tabs, self.device, dev_blob=slot["dev"])
            if slot["copied"] is None:
                slot["copied"] = torch.cuda.Event()
            slot["copied"].record(self.copy_stream)
        self.current = slot
        self._pending = True
        return tabs, ctab, keep

    def before_launch(self):
        """The compute stream waits for the upload of the current tables (once per upload)."""
        if self.current is not None and self._pending:
            torch.cuda.current_stream(self.device).wait_event(self.current["copied"])
            self._pending = False

    def after_launch(self):
        """Marks the current device blob as read by the launches enqueued so far."""
        if self.current is not None:
            if self.current["used"] is None:
                self.current["used"] = torch.cuda.Event()
            self.current["used"].record(torch.cuda.current_stream(self.device))


def _upload_tables(tabs, device, dev_blob=None):
    """One packed H2D copy of all tables; returns (ctypes struct, tensors keeping the memory alive).
    ``tabs``: :class:`NativeTables`, or the NumPy reference builder's ``TvfTables`` (tests; no host mirror, so
    the transposed slab kernel is not available for it)."""
    c = _CTables()
    if isinstance(tabs, NativeTables):
        n = int(tabs.layout.total_words)
        generic = bool(tabs.layout.sections & TABLES_GENERIC)
        first = 0 if generic else int(tabs.layout.offset_words[15])  # unwritten sections are not uploaded either
        if dev_blob is not None:
            blob = dev_blob[:n]
            blob[first:].copy_(tabs.host[first:n], non_blocking=True)
        elif generic:
            blob = tabs.host[:n].to(device, non_blocking=True)
        else:
            blob = torch.empty(n, dtype=torch.int32, device=device)
            blob[first:].copy_(tabs.host[first:n], non_blocking=True)
        base = blob.data_ptr()
        for i, name in enumerate(_TABLE_NAMES):
            setattr(c, name, base + 4 * int(tabs.layout.offset_words[i]) if generic else None)
        c.n_entries, c.nmax, c.inv_c = tabs.n_entries, tabs.nmax, float(tabs.inv_c)
        c.tail_pitch, c.gamma = tabs.layout.tail_pitch, tabs.layout.gamma
        c.host_blob, c.dev_blob = tabs.host.data_ptr(), base
        c.tslab_offset_words = int(tabs.layout.offset_words[15])
        c.cslab_offset_words = int(tabs.layout.cslab_offset_words)
        return c, (blob, tabs.host)
    parts = [tabs.ent_n, tabs.ent_cnt, tabs.ent_idx, tabs.ent_wlo, tabs.ent_whi, tabs.ent_g,
             tabs.tail_idx, tabs.tail_wlo, tabs.tail_whi, tabs.tail_cwlo, tabs.tail_cwhi, tabs.gpow]
    words, offsets = 0, []
    for a in parts:
        offsets.append(words)
        words += (a.size + 3) // 4 * 4  # keep every table 16-byte aligned
    host = np.zeros(words, dtype=np.int32)
    for a, off in zip(parts, offsets):
        host[off:off + a.size] = np.ascontiguousarray(a).reshape(-1).view(np.int32)
    blob = torch.from_numpy(host).to(device, non_blocking=True)
    base = blob.data_ptr()
    c.n_entries, c.nmax = tabs.n_entries, tabs.nmax
    for name, off in zip(_TABLE_NAMES[:12], offsets):
        setattr(c, name, base + 4 * off)
    c.inv_c = float(tabs.inv_c)
    return c, blob


def _launch(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples, n_step_samples,
            use_log_interpolation, path=PATH_AUTO, with_stats=False, keep_on_device=False, numpy_tables=False,
            value_out=None, value_weights=None):
    device = dev.require_cuda()
    lib = _lib.load()
    np_like = dev.is_numpy_like(rewards)
    # Host buffers that are page-locked (pinned tensors, or NumPy arrays registered with ppo_b200.pin_numpy): the
    # env-blocked pipeline moves them at the PCIe rate of both directions together.
    pinned_vs = dev.pinned_tensor_of(value_samples)
    if (pinned_vs is not None and not numpy_tables and value_out is None and rewards.shape[0] > 0
            and len(required_horizons) > 0 and pinned_vs.dtype == torch.float32 and pinned_vs.dim() == 3
            and pinned_vs.stride(2) == 1 and pinned_vs.stride(1) == pinned_vs.shape[2] and not keep_on_device):
        out_host = torch.empty((rewards.shape[0], rewards.shape[1], len(required_horizons)), dtype=torch.float32,
                               pin_memory=True)
        stats = _launch_host_blocks(gamma, rewards, dones,
                                    np.asarray(required_horizons.cpu() if isinstance(required_horizons, torch.Tensor)
                                               else required_horizons),
                                    np.asarray(value_sample_horizons.cpu() if isinstance(value_sample_horizons, torch.Tensor)
                                               else value_sample_horizons),
                                    pinned_vs, n_step_samples, use_log_interpolation, out_host, path=path,
                                    with_stats=with_stats)
        torch.cuda.current_stream().synchronize()
        res = out_host.numpy() if np_like else out_host
        return (res, stats) if with_stats else res
    r = dev.env_contiguous(dev.to_device(rewards, torch.float32, device))
    steps, envs = r.shape
    req = np.asarray(required_horizons.cpu() if isinstance(required_horizons, torch.Tensor) else required_horizons)
    vsh = np.asarray(value_sample_horizons.cpu() if isinstance(value_sample_horizons, torch.Tensor)
                     else value_sample_horizons)
    n_req, n_val = len(req), len(vsh)
    vs = dev.to_device(value_samples, torch.float32, device)
    assert tuple(vs.shape) == (steps + 1, envs, n_val), \
        f"value_samples must be [N+1, A, V] = {(steps + 1, envs, n_val)}, got {tuple(vs.shape)}"
    vs = dev.env_contiguous(vs)  # unit stride on the horizon axis (a [..., vh] slice is copied once)
    d, kind = dev.done_tensor(dones, device)
    assert d is not None and tuple(d.shape) == (steps, envs), f"dones must be [N, A] = {(steps, envs)}"
    d = dev.env_contiguous(d)

    if numpy_tables or steps == 0 or envs == 0 or n_req == 0:  # the documented NumPy builder (tests compare both builders through the kernels)
        tabs = _tvf_tables.build(gamma, steps, req, vsh, n_step_samples, use_log_interpolation)
    else:
        tabs = build_tables(gamma, steps, req, vsh, n_step_samples, use_log_interpolation)
    ctab, blob = _upload_tables(tabs, device)

    out = torch.empty((steps, envs, n_req), dtype=torch.float32, device=device)
    stats = ws = None
    if with_stats:
        stats = torch.empty(_lib.STATS_LEN, dtype=torch.float64, device=device)
        ws = dev.workspace(lib.ppo_tvf_workspace_bytes(steps, envs), device)
    extras, keep = _extras(value_out, value_weights)
    rc = lib.ppo_tvf_targets_ex(
        dev.ptr(r), r.stride(0) if steps else 0, dev.ptr(d), kind, d.stride(0) if steps else 0,
        dev.ptr(vs), vs.stride(0), vs.stride(1) if envs else 0, steps, envs, n_req, n_val,
        ctypes.byref(ctab), path, dev.ptr(out), dev.ptr(stats),
        dev.ptr(ws), ws.numel() if ws is not None else 0, ctypes.byref(extras) if extras is not None else None,
        dev.stream_ptr())
    _lib.check(rc, "ppo_tvf_targets")
    del blob, keep  # same-stream allocator reuse is ordered after the kernel
    res = out if keep_on_device else dev.like_input(out, np_like)
    return (res, stats) if with_stats else res


_pipelines = {}


def _pipeline_state(device, steps, block, n_val, n_req):
    """Streams and double-buffered device scratch of the host-buffer path, cached per device and shape."""
    key = (device.index, steps, block, n_val, n_req)
    st = _pipelines.get(key)
    if st is None:
        st = {
            "s_in": torch.cuda.Stream(device=device), "s_out": torch.cuda.Stream(device=device),
            "s_comp": torch.cuda.Stream(device=device),
            "ws": torch.zeros(max(int(_lib.load().ppo_tvf_workspace_bytes(steps, block)), 1 << 12), dtype=torch.uint8,
                              device=device),
            "vs": [torch.empty((steps + 1, block, n_val), dtype=torch.float32, device=device) for _ in range(2)],
            "out": [torch.empty((steps, block, n_req), dtype=torch.float32, device=device) for _ in range(2)],
        }
        _pipelines.clear()  # one shape at a time: the scratch is ~2 x (value samples + targets) of one env block
        _pipelines[key] = st
    return st


def host_pipeline_ok(value_samples, out):
    """True when the env-blocked host path applies: pinned CPU float32 tensors with contiguous env blocks."""
    return (isinstance(value_samples, torch.Tensor) and isinstance(out, torch.Tensor)
            and value_samples.device.type == "cpu" and out.device.type == "cpu"
            and value_samples.is_pinned() and out.is_pinned()
            and value_samples.dtype == torch.float32 and out.dtype == torch.float32
            and value_samples.dim() == 3 and value_samples.stride(2) == 1
            and value_samples.stride(1) == value_samples.shape[2] and out.is_contiguous())


def _launch_host_blocks(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples,
                        n_step_samples, use_log_interpolation, out, path=PATH_AUTO, with_stats=True, block_envs=None,
                        defer_join=False, value_out=None, value_weights=None):
    """Targets for rollout buffers that live in pinned HOST memory (``value_samples [N+1, A, V]`` in, ``out
    [N, A, K]`` out): the env axis is cut into blocks, and the upload of block c + 1, the kernels of block c and
    the download of block c - 1 run concurrently on three streams (envs are independent, ``rl/returns_truncated.py``
    never mixes columns), so the call costs about max(upload, download) instead of their sum.  Returns the merged
    statistics record of the longest horizon (device tensor) or None.  The results are complete once the current
    stream has been synchronised.  ``defer_join=True`` returns a function instead that joins the side streams
    into the current stream and returns the record, so the caller can enqueue other work (the GAE part of the
    pass) on its own stream while the blocks are in flight."""
    device = dev.require_cuda()
    lib = _lib.load()
    steps, envs = rewards.shape
    req, vsh = np.asarray(required_horizons), np.asarray(value_sample_horizons)
    n_req, n_val = len(req), len(vsh)
    assert tuple(value_samples.shape) == (steps + 1, envs, n_val), \
        f"value_samples must be [N+1, A, V] = {(steps + 1, envs, n_val)}, got {tuple(value_samples.shape)}"
    assert tuple(out.shape) == (steps, envs, n_req), f"out must be [N, A, K] = {(steps, envs, n_req)}"

    sms = torch.cuda.get_device_properties(device).multi_processor_count
    if block_envs is None:
        # about `target` blocks of a whole number of envs per SM; the first and last blocks are smaller so that the
        # pipeline fills (first upload) and drains (last download) quickly
        target = int(os.environ.get("PPO_HOST_BLOCKS", "12"))
        block_envs = int(min(envs, sms * max(1, -(-envs // (target * sms)))))
        ramp = [b for b in (sms // 2, sms, 2 * sms) if 0 < b < block_envs]
        sizes, left = [], envs
        if envs >= 2 * sum(ramp) + block_envs:
            left -= 2 * sum(ramp)
            sizes = ramp + [block_envs] * (left // block_envs) + ([left % block_envs] if left % block_envs else []) + ramp[::-1]
        else:
            sizes = [block_envs] * (envs // block_envs) + ([envs % block_envs] if envs % block_envs else [])
    else:
        block_envs = int(min(block_envs, envs))
        sizes = [block_envs] * (envs // block_envs) + ([envs % block_envs] if envs % block_envs else [])
    starts = [sum(sizes[:c]) for c in range(len(sizes))]
    st = _pipeline_state(device, steps, block_envs, n_val, n_req)
    s_main, s_in, s_out, s_comp = torch.cuda.current_stream(), st["s_in"], st["s_out"], st["s_comp"]
    for s_ in (s_in, s_out, s_comp):
        s_.wait_stream(s_main)
    n_blocks = len(sizes)
    ev_in, ev_comp, ev_out = {}, [], []

    def upload(c):
        """value curves of block c -> scratch c & 1 (after the kernels of block c - 2 have released it)"""
        lo, ab = starts[c], sizes[c]
        if c >= 2:
            s_in.wait_event(ev_comp[c - 2])
        rc = lib.ppo_memcpy_2d_async(st["vs"][c & 1].data_ptr(), ab * n_val * 4,
                                     value_samples.data_ptr() + lo * n_val * 4, value_samples.stride(0) * 4,
                                     ab * n_val * 4, steps + 1, s_in.cuda_stream)
        _lib.check(rc, "ppo_memcpy_2d_async")
        ev_in[c] = torch.cuda.Event()
        ev_in[c].record(s_in)

    # the first two uploads do not depend on anything computed below: start them before the host-side table work
    for c in range(min(2, n_blocks)):
        upload(c)

    r = dev.env_contiguous(dev.to_device(rewards, torch.float32, device))
    d, kind = dev.done_tensor(dones, device)
    assert d is not None and tuple(d.shape) == (steps, envs), f"dones must be [N, A] = {(steps, envs)}"
    d = dev.env_contiguous(d)
    tabs = build_tables(gamma, steps, req, vsh, n_step_samples, use_log_interpolation)
    ctab, blob = _upload_tables(tabs, device)
    stats = torch.zeros((n_blocks, _lib.STATS_LEN), dtype=torch.float64, device=device) if with_stats else None
    ws = st["ws"] if with_stats else None
    s_comp.wait_stream(s_main)  # rewards / terminals / tables / zeroed records were issued on the caller's stream
    for c, ab in enumerate(sizes):
        lo = starts[c]
        vs_d, out_d = st["vs"][c & 1], st["out"][c & 1]
        if c not in ev_in:
            upload(c)
        # ---- kernels (their own stream: the caller's stream stays free for the GAE part of the pass)
        s_comp.wait_event(ev_in[c])
        if c >= 2:
            s_comp.wait_event(ev_out[c - 2])
        extras, keep = _extras(value_out, value_weights, env_offset=lo)
        rc = lib.ppo_tvf_targets_ex(
            r.data_ptr() + lo * 4, r.stride(0), d.data_ptr() + lo * d.element_size(), kind, d.stride(0),
            vs_d.data_ptr(), ab * n_val, n_val, steps, ab, n_req, n_val, ctypes.byref(ctab), path,
            out_d.data_ptr(), stats[c].data_ptr() if with_stats else 0,
            dev.ptr(ws), ws.numel() if ws is not None else 0,
            ctypes.byref(extras) if extras is not None else None, s_comp.cuda_stream)
        _lib.check(rc, "ppo_tvf_targets")
        e = torch.cuda.Event()
        e.record(s_comp)
        ev_comp.append(e)
        # ---- download
        s_out.wait_event(e)
        rc = lib.ppo_memcpy_2d_async(out.data_ptr() + lo * n_req * 4, envs * n_req * 4, out_d.data_ptr(),
                                     ab * n_req * 4, ab * n_req * 4, steps, s_out.cuda_stream)
        _lib.check(rc, "ppo_memcpy_2d_async")
        e = torch.cuda.Event()
        e.record(s_out)
        ev_out.append(e)
    # the inputs (r, d, tables) must outlive the kernels: tie their lifetime to the side stream
    for t_ in (r, d, blob[0] if isinstance(blob, tuple) else blob):
        t_.record_stream(s_comp)
    if stats is not None:
        stats.record_stream(s_comp)
    if value_out is not None:
        value_out.record_stream(s_comp)
    del blob

    def join(kernels_only=False):
        """Make the caller's stream wait for the kernels (value estimate, statistics) and, unless ``kernels_only``,
        for the downloads; returns the merged statistics record.  A caller that passes ``kernels_only=True`` (to
        enqueue work that only needs the kernels' results while the last downloads are still running) must call
        ``join()`` again before it synchronises its stream."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(s_comp)
        if not kernels_only:
            cur.wait_stream(s_out)
        if not with_stats:
            return None
        from . import rollout as _rollout  # merge of disjoint env blocks (sums add, min / max reduce)
        return _rollout.merge_stats(stats)  # on the caller's stream, which now waits for the kernels above

    return join if defer_join else join()


def _calculate_sampled_return_multi_fast(gamma: float, rewards, dones, required_horizons, value_sample_horizons,
                                         value_samples, n_step_samples=None, n_step_list=None,
                                         use_log_interpolation: bool = False, path=PATH_AUTO):
    """``[N, A, K]`` targets for a per-horizon sample table ``[K, C]`` or one list shared by all horizons
    (reference: ``rl/returns_truncated.py:623-693``)."""
    assert n_step_samples is not None or n_step_list is not None
    n_req = len(required_horizons)
    if n_step_list is not None:
        n_step_samples = np.asarray([list(n_step_list) for _ in range(n_req)])  # :645-646
    n_step_samples = np.asarray(n_step_samples)
    assert n_step_samples.ndim == 2 and n_step_samples.shape[0] == n_req
    return _launch(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples,
                   n_step_samples, use_log_interpolation, path)


_weight_cache = {}


def _sample_weights(distribution, steps, n_step):
    """Normalised float32 n-step weights of ``rl/returns_truncated.py:60-80`` (computed with the reference's own
    Python-float expressions, then cached: they depend on the three arguments only)."""
    key = (distribution, int(steps), n_step)
    w = _weight_cache.get(key)
    if w is None:
        lamb = 1 - (1 / n_step)

        def get_weights(f):
            return np.asarray([f(n) for n in range(1, steps + 1)], dtype=np.float32)

        if distribution == "exponential":
            w = get_weights(lambda x: lamb ** x)
        elif distribution == "uniform":
            w = get_weights(lambda x: 1)
        elif distribution == "hyperbolic":
            w = get_weights(lambda x: 1 / x)
        elif distribution == "quadratic":
            w = get_weights(lambda x: 1 / (steps + (x * x)))
        else:
            raise ValueError(f"Invalid distribution {distribution}")
        w /= np.sum(w)
        w.setflags(write=False)
        if len(_weight_cache) > 64:
            _weight_cache.clear()
        _weight_cache[key] = w
    return w


def _choice_1_to_n(weights, size):
    """``np.random.choice(range(1, len(weights) + 1), size=size, replace=True, p=weights)`` of the legacy global
    generator without its per-call argument checks (0.4 ms for 256 weights): the same inverse-CDF draw on the same
    ``random_sample`` stream (numpy/random/mtrand.pyx, ``RandomState.choice``), so the values and the generator
    state afterwards are identical (tests/test_tvf_tables_cpu.py)."""
    cdf = np.asarray(weights, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    idx = cdf.searchsorted(np.random.random_sample(size), side="right")
    return np.asarray(idx, dtype=np.int64) + 1


def draw_n_step_samples(distribution, mode, required_horizons, steps, n_step, max_samples, seed=None):
    """The reference's host-side draw of the n-step table (``rl/returns_truncated.py:60-129``), with the same
    global ``np.random`` draws in the same order so that a seed reproduces the reference's table.
    Returns a ``[K, C]`` int table, or for ``mode == "full"`` the normalised float32 weights ``[N]``."""
    n_req = len(required_horizons)
    if distribution == "fixed":
        return np.zeros([n_req, 1], dtype=np.int32) + n_step
    weights = _sample_weights(distribution, steps, n_step)
    if mode == "full":
        return weights.copy()
    if seed is not None:
        np.random.seed(seed)
    choices = range(1, len(weights) + 1)
    if mode == "standard":
        samples = _choice_1_to_n(weights, (1, max_samples))
        return np.repeat(samples, n_req, axis=0)
    elif mode == "advanced":
        return _choice_1_to_n(weights, (n_req, max_samples))
    elif mode in ("clipped", "adaptive"):
        samples = np.zeros([n_req, max_samples], dtype=np.int32)
        for k in range(n_req):
            max_h = max(required_horizons[k] if mode == "clipped" else required_horizons[k] // 2, 1)
            scaled = weights.copy()
            scaled[max_h:] = 0
            scaled = scaled / scaled.sum()
            samples[k, :] = np.random.choice(choices, size=max_samples, replace=True, p=scaled)
        return samples
    elif mode == "mcx":
        samples = np.zeros([n_req, max_samples], dtype=np.int32)
        for k in range(n_req):
            if required_horizons[k] <= 2 * n_step:
                samples[k, :] = required_horizons[k]
            else:
                samples[k, :] = np.random.choice(choices, size=max_samples, replace=True, p=weights)
        return samples
    elif mode == "full":
        return weights
    raise ValueError(f"Invalid return mode {mode}")


def get_return_estimate(distribution: str, mode: str, gamma: float, rewards, dones, required_horizons,
                        value_sample_horizons, value_samples, n_step: int = 40, max_samples: int = 40,
                        use_log_interpolation: bool = False, seed=None):
    """Weighted average of n-step truncated returns, ``[N, A, K]`` float32
    (reference: ``rl/returns_truncated.py:10-139``; same arguments, modes and error behaviour)."""
    steps = rewards.shape[0]
    n_req = len(required_horizons)
    req_host = np.asarray(required_horizons.cpu() if isinstance(required_horizons, torch.Tensor)
                          else required_horizons)
    drawn = draw_n_step_samples(distribution, mode, req_host, steps, n_step, max_samples, seed)
    if distribution != "fixed" and mode == "full":
        # rl/returns_truncated.py:130-137: sum_n w_n * (n-step return).  The n-step list with
        # multiplicities is exactly an entry table, so this is ONE pass with per-entry weights.
        return _launch_weighted_full(gamma, rewards, dones, required_horizons, value_sample_horizons,
                                     value_samples, drawn, use_log_interpolation)
    return _launch(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples, drawn,
                   use_log_interpolation)


def _launch_weighted_full(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples,
                          weights, use_log_interpolation):
    """``full`` mode: every n = 1..N with weight w_n.  Evaluated as N single-sample passes accumulated on
    the device in the reference's order (it is the reference's "very slow" diagnostic mode, :131-132)."""
    device = dev.require_cuda()
    np_like = dev.is_numpy_like(rewards)
    n_req = len(required_horizons)
    r = dev.to_device(rewards, torch.float32, device)
    d = dev.to_device(dones, None, device)
    vs = dev.to_device(value_samples, torch.float32, device)
    steps = r.shape[0]
    total = torch.zeros((steps, r.shape[1], n_req), dtype=torch.float32, device=device)
    for n, w in zip(range(1, steps + 1), weights):
        table = np.zeros([n_req, 1], dtype=np.int32) + n
        part = _launch(gamma, r, d, required_horizons, value_sample_horizons, vs, table, use_log_interpolation)
        total.add_(part, alpha=float(w))
    return dev.like_input(total, np_like)
