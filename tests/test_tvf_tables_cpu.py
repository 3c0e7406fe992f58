"""CPU: the host index tables reproduce the reference's horizon indexing bit-exactly.

For every (k, n) the table's (lo, hi, w_lo, w_hi) must give exactly `_interpolate`'s answer
(oracle.interp_scalar, pinned to rl/returns_truncated.py:142-174) on a probe value vector.
"""
import numpy as np
import pytest

import oracle
from ppo_b200 import _tvf_tables as T


@pytest.mark.parametrize("use_log", [False, True])
@pytest.mark.parametrize("case", ["geometric", "test_tvf"])
def test_slots_match_interpolate(case, use_log):
    if case == "geometric":
        req = vsh = oracle.get_value_head_horizons(64, 3000)
    else:  # the reference's own test horizons, with duplicates (tests/test_tvf.py:19-20)
        req = np.geomspace(1, 1024, num=16).astype('int32')
        vsh = np.geomspace(1, 1024, num=32).astype('int32') - 1
    steps = 96
    rng = np.random.RandomState(0)
    table = rng.randint(1, steps + 1, size=[len(req), 6])
    tabs = T.build(0.999, steps, req, vsh, table, use_log)
    probe = rng.normal(size=[len(vsh)])  # float64 probe: weights are compared at float32 precision below
    axis = np.log10(10 + vsh) - 1 if use_log else vsh
    for k, h in enumerate(req):
        # multiplicities: counts sum to C and the n' are the distinct clamped samples
        want = np.unique(np.minimum(table[k], h)) if h > 0 else np.asarray([], dtype=np.int64)
        got_n = tabs.ent_n[tabs.ent_cnt[:, k] > 0, k]
        assert np.array_equal(np.sort(got_n), want)
        assert tabs.ent_cnt[:, k].sum() == (table.shape[1] if h > 0 else 0)
        for e in range(tabs.n_entries):
            if tabs.ent_cnt[e, k] == 0:
                continue
            n = int(tabs.ent_n[e, k])
            x = h - n
            lo, hi = tabs.ent_idx[e, k] & 0xffff, tabs.ent_idx[e, k] >> 16
            got = probe[lo] * np.float64(tabs.ent_wlo[e, k]) + probe[hi] * np.float64(tabs.ent_whi[e, k])
            target = (np.log10(10 + x) - 1) if use_log else x
            ref = oracle.interp_scalar(axis, probe, target) if x > 0 else 0.0
            assert abs(got - ref) <= 1e-6 * max(1.0, abs(ref)), (k, h, n)
            # slot identity (bit-exact): the table's slots are searchsorted's
            if x > 0:
                j = int(np.searchsorted(axis, target))
                assert hi == j or lo == j or (axis[j] - axis[j - 1] == 0 and lo == j - 1)
            assert tabs.ent_g[e, k] == tabs.gpow[n]


def test_discount_powers_follow_reference_rounding():
    g = T.discount_powers(0.99997, 200)
    d = np.ones([1], dtype=np.float32)
    for n in range(1, 201):
        d *= 0.99997 * (1 - np.asarray([False]))  # rl/returns_truncated.py:669
        assert g[n] == d[0]


def test_rejects_bad_tables():
    hz = oracle.get_value_head_horizons(16, 100)
    with pytest.raises(AssertionError):
        T.build(0.99, 10, hz, hz, np.zeros([16, 2], dtype=np.int32) + 11)  # n-step > N
    with pytest.raises(AssertionError):
        T.build(0.99, 10, hz, hz, np.zeros([16, 2], dtype=np.int32))       # n-step < 1
    with pytest.raises(IndexError):
        T.build(0.99, 10, np.asarray([500]), hz, np.ones([1, 1], dtype=np.int32))


@pytest.mark.parametrize("use_log", [False, True])
@pytest.mark.parametrize("case", ["geometric", "test_tvf", "production"])
def test_native_builder_matches_numpy_builder(built_library, case, use_log):
    """ppo_tvf_tables_build (C++, the product path) against the documented NumPy construction, bit for bit."""
    from ppo_b200 import returns_truncated as rt
    rng = np.random.RandomState(1)
    if case == "geometric":
        req = vsh = oracle.get_value_head_horizons(64, 3000)
        steps, table = 96, rng.randint(1, 97, size=[64, 6])
    elif case == "test_tvf":
        req = np.geomspace(1, 1024, num=16).astype('int32')
        vsh = np.geomspace(1, 1024, num=32).astype('int32') - 1
        steps, table = 48, rng.randint(1, 49, size=[16, 5])
    else:
        req = vsh = oracle.get_value_head_horizons(128, 30000)
        steps = 256
        table = oracle.draw_n_step_samples("exponential", "advanced", req, steps, 20, 8, seed=1)
    ref = T.build(0.99997, steps, req, vsh, table, use_log)
    nat = rt.build_tables(0.99997, steps, req, vsh, table, use_log)
    assert (nat.n_entries, nat.nmax) == (ref.n_entries, ref.nmax)
    assert np.float32(nat.inv_c) == ref.inv_c and nat.layout.gamma == 0.99997
    for name in ["ent_n", "ent_cnt", "ent_idx", "ent_wlo", "ent_whi", "ent_g", "tail_idx", "tail_wlo", "tail_whi",
                 "tail_cwlo", "tail_cwhi", "gpow"]:
        want = np.ascontiguousarray(getattr(ref, name)).reshape(-1)
        got = nat.table(name)[:want.size]
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), name
    # horizon-major tail tables are the transposes of the row-major ones
    pitch, k = nat.layout.tail_pitch, len(req)
    for name, src in [("tailT_idx", ref.tail_idx), ("tailT_cwlo", ref.tail_cwlo), ("tailT_cwhi", ref.tail_cwhi)]:
        got = nat.table(name)[:k * pitch].reshape(k, pitch)[:, :ref.nmax + 1]
        assert np.array_equal(got.view(np.int32), np.ascontiguousarray(src.T).view(np.int32)), name


def test_native_builder_errors(built_library):
    from ppo_b200 import returns_truncated as rt
    hz = oracle.get_value_head_horizons(16, 100)
    with pytest.raises(AssertionError):
        rt.build_tables(0.99, 10, hz, hz, np.zeros([16, 2], dtype=np.int32) + 11)
    with pytest.raises(AssertionError):
        rt.build_tables(0.99, 10, hz, hz, np.zeros([16, 2], dtype=np.int32))
    with pytest.raises(IndexError):
        rt.build_tables(0.99, 10, np.asarray([500]), hz, np.ones([1, 1], dtype=np.int32))


@pytest.mark.parametrize("seed", range(12))
def test_native_builder_fuzz(built_library, seed):
    """Random horizon sets (duplicates, zeros, dense and sparse), sample counts and rollout lengths: the C++ builder
    and the NumPy construction agree bit for bit, and every horizon's multiplicities sum to the sample count."""
    from ppo_b200 import returns_truncated as rt
    rng = np.random.RandomState(100 + seed)
    steps = int(rng.randint(1, 300))
    n_val = int(rng.choice([4, 8, 32, 64, 128]))
    vsh = np.sort(rng.randint(0, int(rng.choice([20, 300, 40000])), size=n_val))
    vsh[0] = 0
    if seed % 3 == 0:
        vsh = np.unique(vsh)  # strictly ascending variant
    n_req = int(rng.randint(1, 130))
    req = np.sort(rng.randint(0, vsh[-1] + 1, size=n_req))
    c = int(rng.randint(1, 12))
    table = rng.randint(1, steps + 1, size=[n_req, c])
    use_log = bool(seed % 2)
    ref = T.build(0.9997, steps, req, vsh, table, use_log)
    nat = rt.build_tables(0.9997, steps, req, vsh, table, use_log)
    assert (nat.n_entries, nat.nmax) == (ref.n_entries, ref.nmax)
    for name in ["ent_n", "ent_cnt", "ent_idx", "ent_wlo", "ent_whi", "ent_g", "tail_idx", "tail_wlo", "tail_whi",
                 "tail_cwlo", "tail_cwhi", "gpow"]:
        want = np.ascontiguousarray(getattr(ref, name)).reshape(-1)
        assert np.array_equal(nat.table(name)[:want.size].view(np.int32), want.view(np.int32)), (name, seed)
    cnt = nat.table("ent_cnt")[:ref.n_entries * n_req].reshape(ref.n_entries, n_req)
    assert np.array_equal(cnt.sum(axis=0), np.where(req > 0, c, 0))


def test_rediscount_weights_host_arithmetic(golden):
    """Host side of the rediscounted value estimate (no GPU): the per-horizon weights handed to the kernels give the
    reference's sum (rl/tvf.py:388-433) on the golden inputs, and the ratios are the reference's clipped ratios."""
    from ppo_b200 import tvf
    hz, rv = golden["prod_in_horizons"], golden["redis_in_values"].astype(np.float64)
    for old, new, key in ((0.9999, 0.999, "redis_down"), (0.99, 0.9999, "redis_up")):
        w = tvf.rediscount_weights(old, new, [int(h) for h in hz]).astype(np.float64)
        got = rv @ w
        ref = golden[key].astype(np.float64)
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
        rho = tvf.rediscount_ratios(old, new, [int(h) for h in hz])
        assert rho[0] == 0 and rho.max() <= 10 and np.all(rho[1:] > 0)
    # clipping engages when discounting up over long horizons (ratio capped at 10, rl/tvf.py:429)
    assert tvf.rediscount_ratios(0.99, 0.9999, [int(h) for h in hz]).max() == 10


@pytest.mark.parametrize("distribution", ["exponential", "uniform", "hyperbolic", "quadratic"])
@pytest.mark.parametrize("mode", ["standard", "advanced"])
def test_sample_draw_is_numpy_choice_on_the_same_stream(distribution, mode):
    """draw_n_step_samples skips np.random.choice's argument checks but must draw the very same table and leave the
    global generator in the very same state (the reference calls np.random.choice, rl/returns_truncated.py:85-129)."""
    from ppo_b200 import returns_truncated as rt
    hz = oracle.get_value_head_horizons(32, 3000)
    steps, n_step, c = 96, 12, 5
    lamb = 1 - 1 / n_step
    f = {"exponential": lambda x: lamb ** x, "uniform": lambda x: 1, "hyperbolic": lambda x: 1 / x,
         "quadratic": lambda x: 1 / (steps + x * x)}[distribution]
    w = np.asarray([f(n) for n in range(1, steps + 1)], dtype=np.float32)
    w /= np.sum(w)
    for seed in (0, 3, 12345):
        got = rt.draw_n_step_samples(distribution, mode, hz, steps, n_step, c, seed=seed)
        after = np.random.random_sample(4)
        np.random.seed(seed)
        if mode == "standard":
            want = np.repeat(np.random.choice(range(1, steps + 1), size=(1, c), replace=True, p=w), len(hz), axis=0)
        else:
            want = np.random.choice(range(1, steps + 1), size=(len(hz), c), replace=True, p=w)
        assert got.dtype == want.dtype and np.array_equal(got, want)
        assert np.array_equal(after, np.random.random_sample(4))
        # an unseeded draw continues the stream like the reference's call would
        np.random.seed(seed)
        a = rt.draw_n_step_samples(distribution, mode, hz, steps, n_step, c)
        b = rt.draw_n_step_samples(distribution, mode, hz, steps, n_step, c)
        np.random.seed(seed)
        shape = (1, c) if mode == "standard" else (len(hz), c)
        a2 = np.random.choice(range(1, steps + 1), size=shape, replace=True, p=w)
        b2 = np.random.choice(range(1, steps + 1), size=shape, replace=True, p=w)
        assert np.array_equal(a[:shape[0]], a2) and np.array_equal(b[:shape[0]], b2)


def test_native_builder_reuses_a_host_blob(built_library):
    """TableRing's reuse: building into a blob that held another table gives the same words as a fresh build."""
    from ppo_b200 import returns_truncated as rt
    hz = oracle.get_value_head_horizons(32, 3000)
    rng = np.random.RandomState(5)
    t1 = rt.build_tables(0.999, 64, hz, hz, rng.randint(1, 65, size=[32, 6]), False)
    table = rng.randint(1, 30, size=[32, 4])
    fresh = rt.build_tables(0.999, 64, hz, hz, table, False)
    again = rt.build_tables(0.999, 64, hz, hz, table, False, host=t1.host)
    assert again.host.data_ptr() == t1.host.data_ptr()
    n = int(fresh.layout.total_words)
    assert int(again.layout.total_words) == n
    assert np.array_equal(fresh.host.numpy()[:n], again.host.numpy()[:n])


def test_lean_sections_write_the_same_slab_constants(built_library):
    """ppo_tvf_tables_build_ex with the slab section only: same words where the slab kernel reads, the generic tables
    skipped -- unless the slab kernel could not serve the table, then the builder writes them regardless."""
    from ppo_b200 import returns_truncated as rt
    hz = oracle.get_value_head_horizons(128, 30000)
    table = rt.draw_n_step_samples("exponential", "advanced", hz, 256, 20, 8, seed=2)
    full = rt.build_tables(0.99997, 256, hz, hz, table, False)
    lean = rt.build_tables(0.99997, 256, hz, hz, table, False, sections=rt.TABLES_TSLAB)
    assert full.layout.sections == rt.TABLES_ALL and lean.layout.sections == rt.TABLES_TSLAB
    lo, hi = int(full.layout.offset_words[15]), int(full.layout.cslab_offset_words)
    assert (int(lean.layout.offset_words[15]), int(lean.layout.cslab_offset_words)) == (lo, hi)
    assert np.array_equal(full.host.numpy()[lo:hi], lean.host.numpy()[lo:hi])
    assert (lean.n_entries, lean.nmax, lean.inv_c) == (full.n_entries, full.nmax, full.inv_c)
    # the column slab header is invalidated, so that kernel refuses lean tables
    assert lean.host.numpy()[hi + 1] != full.host.numpy()[hi + 1] or full.host.numpy()[hi + 1] < 0
    # a fixed n-step table (one entry per horizon: the ring kernel's case) keeps its generic tables
    fixed = rt.draw_n_step_samples("fixed", "advanced", hz, 256, 20, 8)
    lean_fixed = rt.build_tables(0.99997, 256, hz, hz, fixed, False, sections=rt.TABLES_TSLAB)
    full_fixed = rt.build_tables(0.99997, 256, hz, hz, fixed, False)
    assert lean_fixed.layout.sections & rt.TABLES_GENERIC
    n_fixed = int(full_fixed.layout.cslab_offset_words)
    assert np.array_equal(lean_fixed.host.numpy()[:n_fixed], full_fixed.host.numpy()[:n_fixed])
    # so do shapes the slab kernel does not take (48 horizons)
    hz48 = oracle.get_value_head_horizons(48, 3000)
    t48 = rt.draw_n_step_samples("exponential", "advanced", hz48, 64, 12, 6, seed=1)
    assert rt.build_tables(0.999, 64, hz48, hz48, t48, False, sections=rt.TABLES_TSLAB).layout.sections & rt.TABLES_GENERIC


def test_column_slab_section_restates_the_entries(built_library):
    """The launch constants of the column slab kernel (csrc/tvf_common.cuh: CsHorizon, CsEntry) against the entry tables
    they are derived from: per horizon the same (n-step, multiplicity, slot) set, sorted by shift (n - 1) % 4 then n,
    byte offsets of the aligned row group, weights cnt * gamma^n * lerp / C, and job masks / arrival counts that agree."""
    from ppo_b200 import returns_truncated as rt
    steps, k_n, c_n = 256, 128, 8
    hz = oracle.get_value_head_horizons(k_n, 30000)
    table = rt.draw_n_step_samples("exponential", "advanced", hz, steps, 20, c_n, seed=3)
    tabs = rt.build_tables(0.99997, steps, hz, hz, table, False)
    words = tabs.host.numpy()
    base = int(tabs.layout.cslab_offset_words)
    hdr = words[base:base + 16]
    assert hdr[0] == 0x43534c33 and hdr[2] == steps and hdr[3] == k_n and hdr[4] == k_n // 4
    total = int(hdr[1])
    hzw = words[base + 16:base + 16 + 8 * k_n].view(np.uint32).reshape(k_n, 8)
    jobs = k_n // 4
    jobw = words[base + 16 + 8 * k_n:base + 16 + 8 * k_n + jobs].view(np.uint32)
    ent0 = base + 16 + 8 * k_n + (jobs + 3) // 4 * 4 + 128
    ent = words[ent0:ent0 + 4 * (total + 1)].reshape(total + 1, 4)
    e_n, e_cnt = tabs.table("ent_n").reshape(tabs.n_entries, k_n), tabs.table("ent_cnt").reshape(tabs.n_entries, k_n)
    e_idx = tabs.table("ent_idx").reshape(tabs.n_entries, k_n)
    e_wlo, e_whi = tabs.table("ent_wlo").reshape(tabs.n_entries, k_n), tabs.table("ent_whi").reshape(tabs.n_entries, k_n)
    seen = 0
    masks = np.zeros(jobs, dtype=np.uint32)
    for k in range(k_n):
        beg, maxn, live = int(hzw[k, 0] & 0x7ff), int(hzw[k, 0] >> 11 & 0x3ff), int(hzw[k, 0] >> 23 & 1)
        counts = [int(hzw[k, 1] >> (8 * s) & 0xff) for s in range(4)]
        rows = [e for e in range(tabs.n_entries) if e_cnt[e, k] > 0]
        assert live == int(hz[k] > 0) and sum(counts) == len(rows) and beg == seen
        want = sorted(rows, key=lambda e: ((int(e_n[e, k]) - 1) % 4, int(e_n[e, k])))
        assert maxn == (max(int(e_n[e, k]) for e in rows) if rows else 0)
        pos = beg
        for s in range(4):
            for _ in range(counts[s]):
                e = want[pos - beg]
                n = int(e_n[e, k])
                w = int(ent[pos, 0].view(np.uint32)) if hasattr(ent[pos, 0], "view") else int(ent[pos, 0]) & 0xffffffff
                assert (n - 1) % 4 == s and w >> 18 == n
                lo, hi = int(e_idx[e, k]) & 0xffff, int(e_idx[e, k]) >> 16
                wlo, whi = float(e_wlo[e, k]), float(e_whi[e, k])
                col = k_n if (wlo == 0 and whi == 0) else lo
                assert (w & 0x3fff0) == (col * steps + 4 * ((n - 1) // 4)) * 4
                scale = float(e_cnt[e, k]) * 0.99997 ** n / c_n
                a, b = (1.0, 0.0) if col == k_n else (wlo, whi if hi != lo else 0.0)
                got = ent[pos, 1:3].view(np.float32)
                np.testing.assert_allclose(got, [a * scale, b * scale], rtol=2e-7, atol=0)
                for cc, ww in ((col, got[0]), (col + 1, got[1])):
                    if ww != 0 and cc <= k_n:
                        masks[k // 4] |= np.uint32(1 << (cc // 32 if cc < k_n else 4))
                pos += 1
        seen = pos
    assert seen == total
    # the kernel's job masks contain every column group the entries touch (tails may add value-column groups)
    assert np.all((jobw & masks) == masks)
    arrivals = [int(np.sum(jobw >> g & 1)) for g in range(5)]
    assert list(hdr[5:10]) == arrivals
