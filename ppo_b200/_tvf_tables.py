"""Host-side index tables for the TVF target kernel.

The reference evaluates, per required horizon ``h_k`` and n-step sample ``s``, the n-step return with
``n = min(s, h_k, N - t)`` bootstrapped from the value curve at horizon ``h_k - n``
(``rl/returns_truncated.py:558-620``).  Everything that does not depend on the rollout data is
resolved here, once per call, with the same NumPy calls the reference uses, so that horizon
indexing on the device is bit-exact by construction:

* entries: for every ``k`` the distinct clamped n-steps ``n' = min(s, h_k)`` with their
  multiplicity (the device adds ``count * term``; the reference adds the term ``count`` times);
* for each entry, and for each "tail" length ``m = N - t`` (rows that run off the end of the
  rollout bootstrap from the final row at ``h_k - m``, ``rl/returns_truncated.py:610-614``), the
  ``searchsorted`` slot and the two lerp weights of ``rl/returns_truncated.py:142-174``;
* the discount powers ``G[n]``: a float32 running product multiplied by the float64 gamma and
  rounded back to float32 every step, exactly like ``rl/returns_truncated.py:669``.

No rollout data is touched here; all arithmetic on rewards / dones / values happens on the GPU.
"""
import numpy as np


class TvfTables:
    __slots__ = ("n_entries", "nmax", "ent_n", "ent_cnt", "ent_idx", "ent_wlo", "ent_whi", "ent_g",
                 "tail_idx", "tail_wlo", "tail_whi", "tail_cwlo", "tail_cwhi", "gpow", "inv_c", "n_req", "n_val", "steps")


def _lerp_slots(axis, n_val, targets, positive):
    """Vectorised ``_interpolate`` bookkeeping (rl/returns_truncated.py:151-174).

    ``axis``: ascending float64/int horizon axis [V]; ``targets``: same-domain targets; ``positive``:
    mask of targets whose (linear-domain) horizon is > 0.  Returns (lo, hi, w_lo, w_hi) with
    ``value = v[lo] * w_lo + v[hi] * w_hi``; non-positive targets get zero weights.
    Raises IndexError like the reference when a positive target lies beyond the last horizon.
    """
    targets = np.asarray(targets)
    shape = targets.shape
    tg = targets.reshape(-1)
    pos = np.asarray(positive).reshape(-1)
    j = np.searchsorted(axis, tg, side='left')
    if np.any(pos & (j >= n_val)):
        raise IndexError(f"index {n_val} is out of bounds for axis 0 with size {n_val}")
    j = np.minimum(j, n_val - 1)
    exact = axis[j] == tg
    jm = np.maximum(j - 1, 0)
    width = (axis[j] - axis[jm]).astype(np.float64)
    safe = np.where(width == 0, 1.0, width)
    frac = (tg - axis[jm]) / safe
    single = exact | (j == 0)          # value[..., j] (exact hit or before the first horizon)
    left = (~single) & (width == 0)    # repeated horizons: leftmost value
    lo = np.where(single, j, jm)
    hi = np.where(single | left, lo, j)
    w_lo = np.where(single | left, 1.0, 1.0 - frac)
    w_hi = np.where(single | left, 0.0, frac)
    w_lo = np.where(pos, w_lo, 0.0)
    w_hi = np.where(pos, w_hi, 0.0)
    lo = np.where(pos, lo, 0)
    hi = np.where(pos, hi, 0)
    return (lo.reshape(shape).astype(np.int32), hi.reshape(shape).astype(np.int32),
            w_lo.reshape(shape).astype(np.float32), w_hi.reshape(shape).astype(np.float32))


def discount_powers(gamma, nmax):
    """G[n] as the reference's running float32 product (rl/returns_truncated.py:669)."""
    g = np.ones(nmax + 1, dtype=np.float32)
    step = np.float64(gamma) * np.int64(1)
    for n in range(1, nmax + 1):
        g[n] = np.float32(np.float64(g[n - 1]) * step)
    return g


def build(gamma, steps, required_horizons, value_sample_horizons, n_step_samples, use_log_interpolation=False):
    """Resolve a [K, C] sample table into device tables.  Shapes: entries [E, K], tails [nmax + 1, K]."""
    req = np.asarray(required_horizons).astype(np.int64).reshape(-1)
    vsh = np.asarray(value_sample_horizons).reshape(-1)
    samples = np.asarray(n_step_samples)
    n_req, n_val = len(req), len(vsh)
    assert samples.ndim == 2 and samples.shape[0] == n_req, \
        f"n_step_samples must be [K, C] with K={n_req}, got {samples.shape}"
    n_smp = samples.shape[1]
    assert n_smp >= 1, "at least one n-step sample per horizon is required"
    samples = samples.astype(np.int64)

    active = req > 0                                   # h == 0 -> zeros (rl/returns_truncated.py:580-583)
    clamped = np.minimum(samples, req[:, None])        # :590-591
    # :597 `assert 1 <= target_n_step <= N` for every sample of every non-zero horizon
    assert np.all((clamped[active] >= 1) & (clamped[active] <= steps)), \
        "n-step samples must satisfy 1 <= min(sample, horizon) <= N"

    # ---- de-duplicate per horizon
    order = np.sort(np.where(active[:, None], clamped, 0), axis=1)
    first = np.ones_like(order, dtype=bool)
    first[:, 1:] = order[:, 1:] != order[:, :-1]
    first &= active[:, None]
    per_k = first.sum(axis=1)
    n_entries = int(max(per_k.max() if n_req else 0, 1))
    ent_n = np.zeros([n_entries, n_req], dtype=np.int32)
    ent_cnt = np.zeros([n_entries, n_req], dtype=np.float32)
    rank = np.cumsum(first, axis=1) - 1                # run index of every (sorted) sample
    kk, cc = np.nonzero(first)
    ent_n[rank[kk, cc], kk] = order[kk, cc]
    ka, ca = np.nonzero(np.broadcast_to(active[:, None], order.shape))
    np.add.at(ent_cnt, (rank[ka, ca], ka), 1.0)
    nmax = int(ent_n.max()) if n_req else 0
    nmax = max(nmax, 1)

    # ---- interpolation axis (linear or the reference's log10(10 + h) - 1, :565,:573-574)
    if use_log_interpolation:
        axis = np.log10(10 + vsh) - 1

        def to_axis(x):
            return np.log10(10 + np.maximum(x, -9)) - 1
    else:
        axis = vsh

        def to_axis(x):
            return x

    left = req[None, :] - ent_n.astype(np.int64)       # h - n per entry
    valid = ent_cnt > 0
    lo, hi, w_lo, w_hi = _lerp_slots(axis, n_val, to_axis(left), (left > 0) & valid)
    ent_idx = (lo | (hi << 16)).astype(np.int32)

    ms = np.arange(nmax + 1, dtype=np.int64)[:, None]   # tail length m = N - t
    tail_left = req[None, :] - ms
    # the reference evaluates tails for i < n' only, i.e. m <= max entry of the horizon (:612-614)
    k_nmax = ent_n.max(axis=0).astype(np.int64)[None, :]
    tail_used = (ms >= 1) & (ms <= k_nmax) & active[None, :]
    t_lo, t_hi, t_wlo, t_whi = _lerp_slots(axis, n_val, to_axis(tail_left), (tail_left > 0) & tail_used)
    tail_idx = (t_lo | (t_hi << 16)).astype(np.int32)

    tabs = TvfTables()
    tabs.n_entries, tabs.nmax = n_entries, nmax
    tabs.ent_n, tabs.ent_cnt, tabs.ent_idx = ent_n, ent_cnt, ent_idx
    tabs.ent_wlo, tabs.ent_whi = w_lo, w_hi
    tabs.gpow = discount_powers(gamma, nmax)
    tabs.ent_g = tabs.gpow[ent_n]
    tabs.tail_idx, tabs.tail_wlo, tabs.tail_whi = tail_idx, t_wlo, t_whi
    # folded form for the slab kernel: (sum of multiplicities of the entries with n' > m) * gamma^m * weight
    beyond = (ent_n[None, :, :].astype(np.int64) > ms[:, :, None]) * ent_cnt[None, :, :].astype(np.float64)
    scale = beyond.sum(axis=1) * tabs.gpow.astype(np.float64)[:, None]      # [nmax + 1, K]
    tabs.tail_cwlo = (t_wlo.astype(np.float64) * scale).astype(np.float32)
    tabs.tail_cwhi = (t_whi.astype(np.float64) * scale).astype(np.float32)
    tabs.inv_c = np.float32(1 / n_smp)                   # :618 `returns *= 1 / len(target_n_steps)`
    tabs.n_req, tabs.n_val, tabs.steps = n_req, n_val, steps
    return tabs
