"""ctypes binding of the C ABI declared in ``include/ppo_b200.h``.

The product path has no CPU fallback: if the library cannot be loaded, or no CUDA device is
present, the first compute call raises.  ``symbols()`` lists every exported entry point with
its ctypes signature and is what the CPU test-suite checks against the header.
"""
import ctypes
import os
import threading

from . import _build

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p

PPO_OK = 0
PPO_ERR_BAD_ARG = -1
PPO_ERR_UNSUPPORTED = -2
PPO_ERR_CUDA = -3
PPO_ERR_WORKSPACE = -4

DONE_NONE, DONE_U8, DONE_F32 = 0, 1, 2
STATS_LEN = 16
STAT_COUNT, STAT_ADV_SUM, STAT_ADV_SUMSQ, STAT_ADV_MIN, STAT_ADV_MAX = 0, 1, 2, 3, 4
STAT_RET_SUM, STAT_RET_SUMSQ, STAT_RET_MIN, STAT_RET_MAX = 5, 6, 7, 8
STAT_VAL_SUM, STAT_VAL_SUMSQ, STAT_VAL_COUNT = 9, 10, 11
STAT_ADV_ZEROS, STAT_RET_ZEROS = 12, 13
SCALE_INTRINSIC, SCALE_WRAPPER = 0, 1
MAX_PEERS = 8

# name -> (restype, argtypes); must match include/ppo_b200.h
_SIGNATURES = {
    "ppo_version": (c_int, []),
    "ppo_last_error": (ctypes.c_char_p, []),
    "ppo_device_count": (c_int, []),
    "ppo_gae_workspace_bytes": (c_i64, [c_i64]),
    "ppo_gae": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_int, c_i64,
                        c_i64, c_i64, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ppo_gae_normalized": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_int, c_i64,
                                   c_i64, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                   c_i64, c_ptr]),
    "ppo_adv_normalize": (c_int, [c_ptr, c_i64, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr]),
    "ppo_adv_combine": (c_int, [c_ptr, c_ptr, c_dbl, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ppo_moments": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr]),
    "ppo_tvf_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "ppo_tvf_tables_words": (c_i64, [c_i64, c_i64, c_i64]),
    "ppo_tvf_tables_build": (c_int, [c_dbl, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_int,
                                     c_ptr, c_i64, c_ptr]),
    "ppo_tvf_tables_build_ex": (c_int, [c_dbl, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_int, c_int,
                                        c_ptr, c_i64, c_ptr]),
    "ppo_tvf_targets": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_i64, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "ppo_tvf_targets_ex": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_i64, c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64,
                                   c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "ppo_value_estimate": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "ppo_rediscount_horizons": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr]),
    "ppo_horizon_interpolate": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "ppo_tvf_trim_horizons": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr,
                                      c_ptr, c_ptr]),
    "ppo_tvf_debug_times": (c_int, [c_ptr]),
    "ppo_tvf_last_path": (c_int, []),
    "ppo_tvf_last_horizon_stats": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr]),
    "ppo_memcpy_2d_async": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_ptr]),
    "ppo_vtrace": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_i64, c_i64, c_i64,
                           c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ppo_vtrace_trust_region": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "ppo_discounted_returns": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64,
                                       c_i64, c_i64, c_ptr, c_ptr]),
    "ppo_return_tracker_scan": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_ptr, c_ptr]),
    "ppo_row_sums": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_dbl, c_ptr, c_ptr]),
    "ppo_rms_fold": (c_int, [c_ptr, c_ptr, c_i64, c_dbl, c_ptr, c_ptr, c_ptr]),
    "ppo_rms_fold_ema": (c_int, [c_ptr, c_ptr, c_i64, c_dbl, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr]),
    "ppo_row_moments_fold": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr]),
    "ppo_return_tracker_masked": (c_int, [c_ptr, c_i64, c_ptr, c_int, c_i64, c_ptr, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_ptr,
                                          c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "ppo_gather_rows": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr]),
    "ppo_tvf_value_loss": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_ptr,
                                   c_ptr, c_ptr]),
    "ppo_gae_normalized_peers": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_int, c_i64, c_i64, c_i64,
                                         c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                         c_i64, c_int, c_int, c_ptr, c_ptr, ctypes.c_uint64, c_ptr, c_dbl, c_ptr]),
    "ppo_gae_normalized_fits": (c_int, [c_i64, c_i64, c_int, c_i64]),
    "ppo_peer_alloc": (c_int, [c_i64, c_ptr, c_ptr]),
    "ppo_peer_open": (c_int, [c_ptr, c_ptr]),
    "ppo_peer_close": (c_int, [c_ptr]),
    "ppo_peer_free": (c_int, [c_ptr]),
    "ppo_merge_stats": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "ppo_scale_rows": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "ppo_scale_rewards": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_int, c_int, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr,
                                  c_ptr, c_ptr]),
}

_lock = threading.Lock()
_handle = None


def symbols():
    return dict(_SIGNATURES)


def library_path():
    return _build.LIB_PATH


def load():
    """Load (once) and return the ctypes handle of the in-tree library; raises if it is missing."""
    global _handle
    if _handle is not None:
        return _handle
    with _lock:
        if _handle is not None:
            return _handle
        path = _build.LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(
                f"ppo_b200: CUDA library {path} is not built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'`); there is no CPU fallback")
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _handle = lib
    return _handle


def check(status, what):
    """Translate a negative ppo_status into the exception the reference's callers would see."""
    if status >= 0:
        return status
    msg = load().ppo_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg}" if msg else f"{what}: status {status}"
    if status == PPO_ERR_CUDA:
        raise RuntimeError(text)
    raise ValueError(text)
