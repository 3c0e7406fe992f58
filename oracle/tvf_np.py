"""Oracle (test infrastructure): multi-horizon truncated-return (TVF) targets and horizon helpers.

NumPy restatement of ``rl/returns_truncated.py`` (dispatcher, cached n-step
estimator, scalar interpolation, slow triple-loop checker) and of the free
functions of ``rl/tvf.py`` (``horizon_interpolate``,
``get_rediscounted_value_estimate``, ``get_value_head_horizons``).

Closed form of one target::

    out[t,a,k] = 1/C * sum_c ( S_n(t,a) + D_n(t,a) * I(V[t+n, a, :], h_k - n) )
    n          = min(sample[k,c], h_k, N - t)
    S_n(t,a)   = sum_{i<n} r[t+i,a] * prod_{j<i} gamma * (1 - done[t+j,a])
    D_n(t,a)   = prod_{i<n} gamma * (1 - done[t+i,a])
    I(v, x)    = 0 if x <= 0, else linear interpolation of v over the value horizons

Rounding behaviour kept from the reference because parity is judged on it:
``S`` accumulates float32 products in forward order, ``D`` is a float32 array
multiplied in place by a float64 ``gamma * (1 - done)`` (so every step rounds a
float64 product to float32, ``rl/returns_truncated.py:668-669``), the lerp
factor is a float64 scalar (NumPy >= 2 promotes the lerp to float64 before it
is stored into a float32 buffer), samples are accumulated in sample order and
scaled by ``1 / C`` at the end (``rl/returns_truncated.py:616-618``).
"""
import collections
import math

import numpy as np


# --------------------------------------------------------------------------- interpolation

def interp_scalar(horizons, values, target_horizon):
    """Value at one target horizon; follows ``rl/returns_truncated.py:142-174``.

    ``values[..., j]`` belongs to ``horizons[j]`` (ascending).  A target beyond
    the last horizon raises IndexError exactly like the reference (:157).
    """
    if target_horizon <= 0:
        return values[..., 0] * 0  # :151-153
    j = horizons.searchsorted(target_horizon)  # side='left', :155
    if horizons[j] == target_horizon:  # :157-159 (IndexError when j == len)
        return values[..., j].copy()
    if j == 0:  # :161-162
        return values[..., 0].copy()
    lo = values[..., j - 1]
    hi = values[..., j]
    width = horizons[j] - horizons[j - 1]
    if width == 0:  # :168-170
        return lo.copy()
    frac = (target_horizon - horizons[j - 1]) / width  # :171
    return lo * (1 - frac) + hi * frac  # :174


def horizon_interpolate(horizons, values, target_horizons):
    """Per-row interpolation at per-row targets; follows ``rl/tvf.py:527-573``."""
    assert len(set(horizons)) == len(horizons), f"Horizons duplicates not supported {horizons}"
    assert np.all(np.diff(horizons) > 0), "Horizons must be sorted and unique"
    assert horizons[0] == 0, "first horizon must be 0"
    target_horizons = np.clip(target_horizons, horizons[0], horizons[-1])  # :550
    *lead, n_h = values.shape
    lead = tuple(lead)
    assert horizons.shape == (n_h,)
    assert target_horizons.shape == lead, f"{target_horizons.shape} != {lead}"
    rows = int(np.prod(lead))
    flat_v = values.reshape(rows, n_h)
    flat_t = target_horizons.reshape(rows)
    hi_i = np.searchsorted(horizons, flat_t, side='left')  # :562
    lo_i = np.maximum(hi_i - 1, 0)
    pick = np.arange(rows)
    v_lo = flat_v[pick, lo_i]
    v_hi = flat_v[pick, hi_i]
    width = (horizons[hi_i] - horizons[lo_i])
    width[width == 0] = 1.0  # :569-570
    frac = (flat_t - horizons[lo_i]) / width
    res = v_lo * (1 - frac) + v_hi * frac
    res[hi_i == 0] = 0  # :573
    return res.reshape(*lead)


# --------------------------------------------------------------------------- n-step estimator

def _n_step_tables(gamma, rewards, dones, wanted, sweep):
    """Forward sweep of ``sweep`` steps keeping ``S_n`` / ``D_n`` for every n in ``wanted``.

    Follows ``rl/returns_truncated.py:662-672``: row ``t`` stops updating once
    ``t + i >= N`` so ``S_n(t) = S_{N-t}(t)`` for ``n > N - t``.
    """
    steps, envs = rewards.shape
    run_s = np.zeros([steps, envs], dtype=np.float32)
    run_d = np.ones([steps, envs], dtype=np.float32)
    s_tab, d_tab = {}, {}
    for i in range(sweep):
        run_s[:steps - i] += rewards[i:] * run_d[:steps - i]
        run_d[:steps - i] *= gamma * (1 - dones[i:])
        if (i + 1) in wanted:
            s_tab[i + 1] = run_s.copy()
            d_tab[i + 1] = run_d.copy()
    return s_tab, d_tab


def _one_horizon(samples_k, h, rewards, vs_horizons, value_samples, s_tab, d_tab, use_log):
    """All samples of one required horizon; follows ``rl/returns_truncated.py:558-620``."""
    steps, envs = rewards.shape
    if use_log:
        axis = np.log10(10 + vs_horizons) - 1  # :565

        def lookup(v, x):
            return interp_scalar(axis, v, np.log10(10 + x) - 1)  # :573-574
    else:
        def lookup(v, x):
            return interp_scalar(vs_horizons, v, x)

    if h == 0:  # :580-583
        return np.zeros([steps, envs], dtype=np.float32)
    total = np.zeros_like(rewards, dtype=np.float32)
    boot = np.zeros([steps, envs], dtype=np.float32)
    for n in samples_k:
        n = h if n > h else n  # :590-591
        assert 1 <= n <= steps  # :597
        boot *= 0
        left = h - n
        if left > 0:  # :606-607 rows that bootstrap inside the rollout
            boot[:steps - n] = lookup(value_samples[n:-1], left)
        for i in range(n):  # :610-614 rows that run off the end bootstrap from the final row
            boot[steps - i - 1] = lookup(value_samples[-1], h - i - 1)
        total += s_tab[n] + boot * d_tab[n]  # :616
    total *= 1 / len(samples_k)  # :618
    return total


def sampled_return_multi(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples,
                         n_step_samples=None, n_step_list=None, use_log_interpolation=False):
    """Cached estimator ``[N, A, K]``; follows ``rl/returns_truncated.py:623-693``."""
    assert n_step_samples is not None or n_step_list is not None
    steps, envs = rewards.shape
    n_req = len(required_horizons)
    if n_step_list is not None:
        n_step_samples = np.asarray([list(n_step_list) for _ in range(n_req)])  # :645-646
    assert n_step_samples.shape == (n_req, n_step_samples.shape[1])
    wanted = set(int(x) for x in n_step_samples.ravel()) | set(int(h) for h in required_horizons)  # :648,661,670
    s_tab, d_tab = _n_step_tables(gamma, rewards, dones, wanted, int(n_step_samples.max()))  # sweep length :667
    out = np.zeros([steps, envs, n_req], dtype=np.float32)
    for k, h in enumerate(required_horizons):
        out[:, :, k] = _one_horizon(n_step_samples[k], h, rewards, value_sample_horizons, value_samples,
                                    s_tab, d_tab, use_log_interpolation)
    return out


def sampled_return_slow(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples,
                        n_step_samples=None, n_step_list=None):
    """Triple-loop checker (first moment only); follows ``rl/returns_truncated.py:430-555``.

    Duplicate n-steps are merged into weights (:407-427) and every (h, n, t) is
    evaluated from scratch.  Only meant for small cases.
    """
    assert (n_step_list is None) != (n_step_samples is None)
    steps, envs = rewards.shape
    n_req = len(required_horizons)
    if n_step_samples is not None:  # :468-484 one horizon at a time
        out = np.zeros([steps, envs, n_req], dtype=np.float32)
        for k in range(n_req):
            out[:, :, k:k + 1] = sampled_return_slow(
                gamma, rewards, dones, np.asarray([required_horizons[k]]), value_sample_horizons, value_samples,
                n_step_list=n_step_samples[k, :])
        return out
    clipped = list(np.clip(n_step_list, 1, None))  # :414
    weight_of = collections.defaultdict(float)
    for n in clipped:
        weight_of[n] += 1.0
    distinct = sorted(set(clipped))
    total_w = sum(weight_of.values())
    out = np.zeros([steps, envs, n_req], dtype=np.float32)
    for k, h in enumerate(required_horizons):
        for target_n in distinct:
            w = weight_of[target_n] / total_w
            for t in range(steps):
                acc = np.zeros([envs], dtype=np.float32)
                disc = np.ones([envs], dtype=np.float32)
                n = min(target_n, steps - t, h)  # :516
                for i in range(n):  # :519-522
                    acc += rewards[t + i] * disc
                    disc *= gamma
                    disc *= (1 - dones[t + i])
                boot = 0
                if h - n > 0:  # :529-530
                    boot = interp_scalar(value_sample_horizons, value_samples[t + n], h - n) * disc
                out[t, :, k] += (acc + boot) * w  # :542
    return out


# --------------------------------------------------------------------------- dispatcher

def draw_n_step_samples(distribution, mode, required_horizons, steps, n_step, max_samples, seed=None):
    """Host-side draw of the ``[K, C]`` n-step table; follows ``rl/returns_truncated.py:60-129``.

    Uses the same global ``np.random`` calls in the same order as the reference, so a given
    seed (or global RNG state) produces the same table.  ``full`` mode has no table and is
    handled by the caller.
    """
    n_req = len(required_horizons)
    lamb = 1 - (1 / n_step)  # :60
    if distribution == "fixed":  # :65-68
        return np.zeros([n_req, 1], dtype=np.int32) + n_step
    if distribution == "exponential":
        shape = [lamb ** n for n in range(1, steps + 1)]
    elif distribution == "uniform":
        shape = [1 for n in range(1, steps + 1)]
    elif distribution == "hyperbolic":
        shape = [1 / n for n in range(1, steps + 1)]
    elif distribution == "quadratic":
        shape = [1 / (steps + (n * n)) for n in range(1, steps + 1)]
    else:
        raise ValueError(f"Invalid distribution {distribution}")
    weights = np.asarray(shape, dtype=np.float32)
    weights /= np.sum(weights)  # :80
    if seed is not None:
        np.random.seed(seed)  # :82-83
    support = range(1, len(weights) + 1)
    if mode == "standard":  # :85-89
        row = np.random.choice(support, size=(1, max_samples), replace=True, p=weights)
        return np.repeat(row, n_req, axis=0)
    if mode == "advanced":  # :90-93
        return np.random.choice(support, size=(n_req, max_samples), replace=True, p=weights)
    if mode in ("clipped", "adaptive"):  # :94-117
        table = np.zeros([n_req, max_samples], dtype=np.int32)
        for k in range(n_req):
            h = required_horizons[k] if mode == "clipped" else required_horizons[k] // 2
            cut = max(h, 1)
            w = weights.copy()
            w[cut:] = 0
            w = w / w.sum()
            table[k, :] = np.random.choice(support, size=max_samples, replace=True, p=w)
        return table
    if mode == "mcx":  # :118-129
        table = np.zeros([n_req, max_samples], dtype=np.int32)
        for k in range(n_req):
            if required_horizons[k] <= 2 * n_step:
                table[k, :] = required_horizons[k]
            else:
                table[k, :] = np.random.choice(support, size=max_samples, replace=True, p=weights)
        return table
    if mode == "full":
        return weights
    raise ValueError(f"Invalid return mode {mode}")


def get_return_estimate(distribution, mode, gamma, rewards, dones, required_horizons, value_sample_horizons,
                        value_samples, n_step=40, max_samples=40, use_log_interpolation=False, seed=None):
    """Dispatcher ``[N, A, K]``; follows ``rl/returns_truncated.py:10-139``."""
    steps, envs = rewards.shape
    n_req = len(required_horizons)

    def run(table):
        return sampled_return_multi(gamma, rewards, dones, required_horizons, value_sample_horizons, value_samples,
                                    n_step_samples=table, use_log_interpolation=use_log_interpolation)

    drawn = draw_n_step_samples(distribution, mode, required_horizons, steps, n_step, max_samples, seed)
    if distribution != "fixed" and mode == "full":  # :130-137
        out = np.zeros([steps, envs, n_req], dtype=np.float32)
        for n, w in zip(range(1, steps + 1), drawn):
            out += run(np.zeros([n_req, 1], dtype=np.int32) + n) * w
        return out
    return run(drawn)


# --------------------------------------------------------------------------- rl/tvf.py free functions

def get_rediscounted_value_estimate(values, old_gamma, new_gamma, horizons, clipping=10):
    """Rediscount a ``[B, K]`` truncated value curve to a new gamma; follows ``rl/tvf.py:388-433``.

    NumPy float32 accumulation stands in for the reference's torch float32 tensor ops
    (same elementwise order: difference, times a python-float ratio, accumulate).
    """
    n_rows, n_h = values.shape
    if old_gamma == new_gamma:  # :405-406
        return values[:, -1]
    assert n_h == len(horizons), f"missmatch {n_h} {horizons}"
    assert horizons[0] == 0, 'first horizon must be 0'
    values = np.asarray(values, dtype=np.float32)
    acc = np.zeros([n_rows], dtype=np.float32)
    prev_col = values[:, 0]
    prev_h = 0
    for j in range(1, n_h):
        h = horizons[j]
        mid = ((prev_h + 1 + h) / 2) - 1  # :424
        step = values[:, j] - prev_col
        prev_col = values[:, j]
        prev_h = h
        ratio = min((new_gamma ** mid) / (old_gamma ** mid), clipping)  # :429
        acc += step * np.float32(ratio)
    return acc


def rediscount_horizons(values, old_gamma, new_gamma, horizons):
    """Prefix form of the above over k; follows ``rl/tvf.py:362-380``."""
    if old_gamma == new_gamma:
        return values
    out = values.copy()
    for k in range(values.shape[1]):
        out[:, k] = get_rediscounted_value_estimate(values[:, :k + 1], old_gamma, new_gamma, horizons[:k + 1])
    return out


def get_value_head_horizons(n_heads, max_horizon, spacing="geometric", include_weight=False):
    """Horizon table for the value heads; follows ``rl/tvf.py:576-610``."""
    if spacing == "linear":
        res = np.asarray(np.round(np.linspace(0, max_horizon, n_heads)), dtype=np.int32)
        return (res, np.ones([n_heads], dtype=np.float32)) if include_weight else res
    if spacing != "geometric":
        raise ValueError(f"Invalid spacing value {spacing}")

    def heads(x, unique=True):
        raw = np.asarray(np.round(np.geomspace(1, max_horizon + 1, x)) - 1, dtype=np.int32)
        return np.asarray(sorted(set(raw))) if unique else np.asarray(sorted(raw))

    trial = n_heads
    while len(heads(trial)) != n_heads:  # :594-598
        if len(heads(trial)) < n_heads:
            trial += int(math.sqrt(n_heads))
        else:
            trial -= 1
    multiplicity = collections.Counter()
    order = []
    for h in heads(trial, unique=False):
        if h not in multiplicity:
            order.append(h)
        multiplicity[h] += 1
    res = np.asarray(order, dtype=np.int32)
    counts = np.asarray([multiplicity[h] for h in order], dtype=np.float32)
    return (res, counts) if include_weight else res


def trim_horizons(horizons, tvf_value_estimates, time, method="timelimit", mode="interpolate", timeout=27000,
                  trim_clip=-1, episode_length_buffer=(), eta_percentile=90, eta_buffer=32, eta_minh=128,
                  randint=None):
    """Pull value estimates of horizons that extend past the time limit back to the limit; follows
    ``TVFRunnerModule.trim_horizons`` (``rl/tvf.py:91-208``) with the module state passed in explicitly.

    ``tvf_value_estimates [A, K, VH]``, ``time [A]`` -> ``(trimmed [A, K, VH] (``[A, K, 1]`` in substitute mode),
    final value estimate [A] float32 or 0, time till termination [A] or None)``.  ``randint(lo, hi)`` stands in for
    ``np.random.randint`` in the ``random`` mode.
    """
    horizons = np.asarray(horizons)
    old = tvf_value_estimates.copy()  # :106-107 two private copies
    new = tvf_value_estimates.copy()
    assert horizons[0] == 0, "First horizon must be zero"  # :110
    new[:, 0, :] = 0  # :111-112
    old[:, 0, :] = 0
    if method == "off":  # :116-117
        return new, 0, None
    elif method == "timelimit":  # :118-119
        ttt = np.maximum(timeout - time, 0)
    elif method == "est_term":  # :120-127
        est_len = np.percentile(list(episode_length_buffer) + list(time), eta_percentile).astype(int) + eta_buffer
        est_len += eta_minh / 4
        ttt = np.maximum(timeout - time, 0)
        est_ttt = np.maximum(est_len - time, eta_minh)
        ttt = np.minimum(ttt, est_ttt)
    else:
        raise ValueError(f"Invalid trimming method {method}")
    n_envs, n_h, n_vh = new.shape
    trimmed_ks = np.searchsorted(horizons, ttt)  # :137
    if mode == "interpolate":  # :140-155, interpolation on the log scale
        def log_scale(x):
            return np.log10(10 + x) - 1
        value = horizon_interpolate(log_scale(horizons), old[..., 0], log_scale(ttt))
        for a in range(n_envs):
            new[a, trimmed_ks[a]:, 0] = value[a]
    elif mode == "average":  # :156-171 running mean of the untrimmed estimates from the trimming point on
        for a, tk in enumerate(trimmed_ks):
            if tk >= n_h - 1:
                continue
            acc, cnt = 0, 0
            for k in range(tk, n_h):
                acc = acc + old[a, k, :]
                cnt += 1
                new[a, k, :] = acc / cnt
    elif mode == "substitute":  # :172-178 shortest admissible horizon, ext head only
        ks = np.minimum(trimmed_ks[:, None].copy(), np.arange(n_h)[None, :])
        new = np.take_along_axis(new[:, :, 0], ks, axis=1)[:, :, None]
    elif mode == "random":  # :179-189
        randint = randint or np.random.randint
        for a, tk in zip(range(n_envs), trimmed_ks):
            ks = np.arange(n_h)
            for k in range(tk, n_h):
                ks[k] = randint(tk, k + 1)
            new[a, range(n_h)] = old[a, ks]
    else:
        raise ValueError(f"Invalid trimming mode {mode}")
    final = np.zeros([n_envs], dtype=np.float32)  # :193-196 mean of the valid untrimmed ext estimates
    for a, tk in enumerate(trimmed_ks):
        final[a] = old[a, min(tk, n_h - 1):, 0].mean()
    if trim_clip >= 0:  # :199-201
        new = old + np.clip(new - old, -trim_clip, +trim_clip)
    return new, final, ttt
