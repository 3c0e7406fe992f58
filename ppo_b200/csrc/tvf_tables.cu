// Host-side builder of the TVF index tables (ppo_tvf_tables_build): integers and table arithmetic only,
// no rollout data.  Same construction as ppo_b200/_tvf_tables.py (which stays as the documented
// reference and is compared with this builder bit for bit in tests/test_tvf_tables_cpu.py):
//   * per horizon h_k the distinct clamped n-steps n' = min(sample, h_k) with multiplicities
//     (rl/returns_truncated.py:590-591, :616-618);
//   * per entry, and per tail length m = N - t, the searchsorted slot and the lerp weights of
//     rl/returns_truncated.py:142-174 (float64 arithmetic like NumPy, rounded to float32 once);
//   * discount powers as the reference's float32 running product times a float64 gamma (:669).
// It also writes horizon-major copies of the folded tail tables (what the transposed slab kernel reads with
// lanes along time).
#include <math.h>
#include <string.h>

#include <limits.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "tvf_common.cuh"

namespace ppo {
namespace {

struct Lerp {
    int32_t lo, hi;
    float wlo, whi;
};

// Vectorised `_interpolate` bookkeeping for one target (ppo_b200/_tvf_tables.py:_lerp_slots).
// Returns false when a positive target lies beyond the last horizon (the reference raises IndexError).
bool lerp_slot(const std::vector<double>& axis, double tg, bool positive, Lerp* out) {
    const int n_val = static_cast<int>(axis.size());
    int j = static_cast<int>(std::lower_bound(axis.begin(), axis.end(), tg) - axis.begin());
    if (positive && j >= n_val) return false;
    j = std::min(j, n_val - 1);
    const bool exact = axis[j] == tg;
    const int jm = std::max(j - 1, 0);
    const double width = axis[j] - axis[jm];
    const double safe = width == 0 ? 1.0 : width;
    const double frac = (tg - axis[jm]) / safe;
    const bool single = exact || j == 0;
    const bool left = !single && width == 0;
    int lo = single ? j : jm;
    int hi = (single || left) ? lo : j;
    double wlo = (single || left) ? 1.0 : 1.0 - frac;
    double whi = (single || left) ? 0.0 : frac;
    if (!positive) lo = 0, hi = 0, wlo = 0.0, whi = 0.0;
    out->lo = lo, out->hi = hi;
    out->wlo = static_cast<float>(wlo), out->whi = static_cast<float>(whi);
    return true;
}

struct Entry {
    int32_t n;
    float cnt;
    Lerp l;
};

// Sample-independent part of the tail tables, cached per horizon set (it depends on the horizons, the
// interpolation axis and the number of steps only).
struct TailCache {
    std::vector<int64_t> req, vsh;
    int64_t steps = -1;
    int use_log = -1;
    std::vector<Lerp> slots;  // [K][steps + 1]: slot and weights of horizon h_k - m
    std::vector<unsigned char> ok;
};
std::mutex g_cache_mutex;
TailCache g_cache;

// Scratch of one build (per thread, reused: a build per rollout must not pay for allocations).
struct Scratch {
    std::vector<double> axis, gp64;
    std::vector<Entry> ent;          // [K][C], ent_count[k] live entries in ascending n
    std::vector<int> ent_count, k_nmax;
    std::vector<float> gpow;
    std::vector<unsigned char> beyond;  // [K][nmax + 1] samples of the horizon with n > m (valid for C <= 255)
    std::vector<uint32_t> tail_mask;    // [K] column groups the tail bootstraps of the horizon read
    std::vector<float> park;
};

}  // namespace
}  // namespace ppo

using namespace ppo;

extern "C" int64_t ppo_tvf_tables_words(int64_t n_req, int64_t n_samples, int64_t steps) {
    if (n_req < 0 || n_samples < 1 || steps < 0) return -1;
    const int64_t e = std::min<int64_t>(n_samples, std::max<int64_t>(steps, 1));
    auto pad = [](int64_t w) { return (w + 3) / 4 * 4; };
    return 6 * pad(e * n_req) + 5 * pad((steps + 1) * n_req) + pad(steps + 1) + 3 * pad(n_req * pad(steps + 1)) +
           pad(4 + 4 * n_req + 4 * e * n_req) + 4 * n_req * pad(steps + 1) + 16 +
           kCsHeaderWords + 8 * n_req + pad((n_req + 3) / 4) + 128 + 4 * (e * n_req + 1) + pad(n_req * (steps + 8) / 8 + 8) + 16;
}

extern "C" int ppo_tvf_tables_build(double gamma, int64_t steps, const int64_t* required_horizons, int64_t n_req,
                                    const int64_t* value_horizons, int64_t n_val, const int64_t* samples,
                                    int64_t n_samples, int use_log_interpolation,
                                    int32_t* blob, int64_t blob_words, ppo_tvf_table_layout* layout) {
    return ppo_tvf_tables_build_ex(gamma, steps, required_horizons, n_req, value_horizons, n_val, samples, n_samples,
                                   use_log_interpolation, PPO_TVF_TABLES_ALL, blob, blob_words, layout);
}

extern "C" int ppo_tvf_tables_build_ex(double gamma, int64_t steps, const int64_t* required_horizons, int64_t n_req,
                                       const int64_t* value_horizons, int64_t n_val, const int64_t* samples,
                                       int64_t n_samples, int use_log_interpolation, int sections,
                                       int32_t* blob, int64_t blob_words, ppo_tvf_table_layout* layout) {
    PPO_REQUIRE(required_horizons && value_horizons && samples && blob && layout, "ppo_tvf_tables_build: null pointer");
    PPO_REQUIRE(sections > 0 && (sections & ~PPO_TVF_TABLES_ALL) == 0, "ppo_tvf_tables_build: unknown sections %d", sections);
    PPO_REQUIRE(n_req >= 1 && n_val >= 1 && n_samples >= 1 && steps >= 1, "ppo_tvf_tables_build: bad shape");
    const int K = static_cast<int>(n_req), V = static_cast<int>(n_val), C = static_cast<int>(n_samples);

    static thread_local Scratch sc;
    std::vector<double>& axis = sc.axis;
    axis.resize(V);
    for (int j = 0; j < V; ++j)
        axis[j] = use_log_interpolation ? log10(10.0 + static_cast<double>(value_horizons[j])) - 1.0
                                        : static_cast<double>(value_horizons[j]);
    auto to_axis = [&](int64_t x) {
        return use_log_interpolation ? log10(10.0 + static_cast<double>(std::max<int64_t>(x, -9))) - 1.0
                                     : static_cast<double>(x);
    };

    // ---- bootstrap slots of every (horizon, distance m): horizon h_k - m on the value axis.  They serve the entries
    // (m = n-step) and the tails (m = N - t) alike and depend on the horizon set only, so they are cached.
    std::lock_guard<std::mutex> guard(g_cache_mutex);
    TailCache& tc = g_cache;
    const size_t cpitch = static_cast<size_t>(steps) + 1;
    {
        const bool hit = tc.steps == steps && tc.use_log == use_log_interpolation &&
                         tc.req.size() == static_cast<size_t>(K) && tc.vsh.size() == static_cast<size_t>(V) &&
                         std::equal(tc.req.begin(), tc.req.end(), required_horizons) &&
                         std::equal(tc.vsh.begin(), tc.vsh.end(), value_horizons);
        if (!hit) {
            tc.req.assign(required_horizons, required_horizons + K);
            tc.vsh.assign(value_horizons, value_horizons + V);
            tc.steps = steps, tc.use_log = use_log_interpolation;
            tc.slots.assign(cpitch * K, Lerp{0, 0, 0.f, 0.f});
            tc.ok.assign(cpitch * K, 1);
            for (int k = 0; k < K; ++k)
                for (int64_t m = 1; m <= steps; ++m) {
                    const int64_t left = required_horizons[k] - m;
                    if (required_horizons[k] <= 0 || left <= 0) continue;  // no bootstrap: zero weights
                    if (!lerp_slot(axis, to_axis(left), true, &tc.slots[k * cpitch + m])) tc.ok[k * cpitch + m] = 0;
                }
        }
    }

    // ---- entries: distinct clamped n-steps per horizon, ascending, with multiplicities
    std::vector<Entry>& ent = sc.ent;
    std::vector<int>& ent_count = sc.ent_count;
    std::vector<int>& k_nmax = sc.k_nmax;
    ent.resize(static_cast<size_t>(K) * C);
    ent_count.assign(K, 0), k_nmax.assign(K, 0);
    int n_entries = 1, nmax = 1;
    for (int k = 0; k < K; ++k) {
        const int64_t h = required_horizons[k];
        if (h <= 0) continue;  // h == 0 -> zeros (rl/returns_truncated.py:580-583)
        Entry* row = ent.data() + static_cast<size_t>(k) * C;
        // insertion sort of the clamped samples into (n, multiplicity) pairs
        int cnt = 0;
        for (int c = 0; c < C; ++c) {
            const int64_t n = std::min(samples[static_cast<int64_t>(k) * C + c], h);
            if (n < 1 || n > steps) {
                set_error("n-step samples must satisfy 1 <= min(sample, horizon) <= N (horizon %lld, sample %lld, N %lld)",
                          (long long)h, (long long)samples[static_cast<int64_t>(k) * C + c], (long long)steps);
                return PPO_ERR_BAD_ARG;  // rl/returns_truncated.py:597 assert
            }
            int pos = cnt;
            while (pos > 0 && row[pos - 1].n > n) --pos;
            if (pos > 0 && row[pos - 1].n == n) {
                row[pos - 1].cnt += 1.f;
                continue;
            }
            for (int i = cnt; i > pos; --i) row[i] = row[i - 1];
            row[pos].n = static_cast<int32_t>(n), row[pos].cnt = 1.f;
            ++cnt;
        }
        for (int e = 0; e < cnt; ++e) {
            if (!tc.ok[k * cpitch + row[e].n]) {
                set_error("index %d is out of bounds for axis 0 with size %d", V, V);
                return PPO_ERR_UNSUPPORTED;  // the reference's IndexError (rl/returns_truncated.py:157)
            }
            row[e].l = tc.slots[k * cpitch + row[e].n];
        }
        ent_count[k] = cnt, k_nmax[k] = row[cnt - 1].n;
        n_entries = std::max(n_entries, cnt);
        nmax = std::max(nmax, k_nmax[k]);
    }
    const int E = n_entries;

    // ---- discount powers (float32 running product, float64 multiply); float64 powers for the slab kernels
    std::vector<float>& gpow = sc.gpow;
    std::vector<double>& gp64 = sc.gp64;
    gpow.resize(nmax + 1), gp64.resize(nmax + 1);
    gpow[0] = 1.f, gp64[0] = 1.0;
    for (int n = 1; n <= nmax; ++n) {
        gpow[n] = static_cast<float>(static_cast<double>(gpow[n - 1]) * gamma);
        gp64[n] = pow(gamma, static_cast<double>(n));
    }

    // ---- pack
    auto pad = [](int64_t w) { return (w + 3) / 4 * 4; };
    const int64_t ek = static_cast<int64_t>(E) * K, tk = static_cast<int64_t>(nmax + 1) * K;
    int64_t off = 0;
    int64_t* offs = layout->offset_words;
    for (int i = 0; i < 6; ++i) offs[i] = off, off += pad(ek);
    for (int i = 6; i < 11; ++i) offs[i] = off, off += pad(tk);
    offs[11] = off, off += pad(nmax + 1);
    // horizon-major copies of the folded tail tables, [K][tail_pitch], for the transposed slab kernel
    const int64_t tpitch = pad(nmax + 1);
    for (int i = 12; i < 15; ++i) offs[i] = off, off += pad(static_cast<int64_t>(K) * tpitch);
    // launch constants of the transposed slab kernel: header, K horizon records, the live entries
    // (+ the folded tail tables once more, interleaved {idx, cwlo, cwhi, 0} per (horizon, tail length): one 16-byte
    // load per row instead of three 4-byte loads)
    offs[15] = off, off += pad(4 + 4 * static_cast<int64_t>(K) + 4 * ek) + 4 * static_cast<int64_t>(K) * tpitch;
    // launch constants of the column slab kernel: header, K horizon words, K / 4 job words, the live entries + 1
    const int64_t cs_off = off;
    off += kCsHeaderWords + 8 * static_cast<int64_t>(K) + pad((K + 3) / 4) + 128 + 4 * (ek + 1) +
           pad(static_cast<int64_t>(K) * (steps + 8) / 8 + 8);
    if (off > blob_words) {
        set_error("ppo_tvf_tables_build: blob of %lld words required, got %lld", (long long)off, (long long)blob_words);
        return PPO_ERR_WORKSPACE;
    }
    // ---- which sections: the transposed slab constants first (cheap), because whether the generic tables can be
    // skipped depends on whether the slab kernel can serve this table at all
    bool want_generic = (sections & PPO_TVF_TABLES_GENERIC) != 0;
    const bool want_tslab = (sections & PPO_TVF_TABLES_TSLAB) != 0, want_cslab = (sections & PPO_TVF_TABLES_CSLAB) != 0;
    int32_t* ts_hdr = blob + offs[15];
    int32_t* t4 = ts_hdr + pad(4 + 4 * static_cast<int64_t>(K) + 4 * ek);  // [K][tail_pitch] x {idx, cwlo / C, cwhi / C, 0}
    bool tslab_ok = false;
    memset(ts_hdr, 0, static_cast<size_t>(cs_off - offs[15]) * 4);
    {
        // Per (horizon, n-step) constants of the transposed slab kernel: byte offset of W^T[col][n - 1] in a slab of
        // pitch N | 1, multiplicity * gamma^n * lerp weights (float64 product, rounded once), n.  Entries without a
        // bootstrap read the -R column (index V) with weight multiplicity * gamma^n; single-slot values read their
        // neighbour column with weight zero.
        int32_t* hdr = ts_hdr;
        const int pitch = static_cast<int>(steps) | 1;
        hdr[0] = kTsConstMagic, hdr[1] = -1, hdr[2] = pitch, hdr[3] = V;
        bool ok = want_tslab && K <= kTsMaxK && ek <= kTsMaxEntries;
        TsHorizon* hz = reinterpret_cast<TsHorizon*>(hdr + 4);
        TsEntry* te = reinterpret_cast<TsEntry*>(hdr + 4 + 4 * K);
        int total = 0;
        for (int k = 0; k < K && ok; ++k) {
            const Entry* row = ent.data() + static_cast<size_t>(k) * C;
            TsHorizon& h = hz[k];
            h.beg = total, h.cnt = 0, h.maxn = 0;
            double csum = 0.0;
            for (int e = 0; e < ent_count[k] && ok; ++e) {
                const Entry& x = row[e];
                double a = x.l.wlo, bw = x.l.whi;
                int col;
                if (a == 0.0 && bw == 0.0) col = V, a = 1.0;
                else if (x.l.hi == x.l.lo) col = x.l.lo, bw = 0.0;
                else if (x.l.hi == x.l.lo + 1) col = x.l.lo;
                else { ok = false; break; }
                // the 1 / C of the sample average is folded in here (and into the tail records below), so the kernel
                // finishes a target with one multiply-add
                const double cg = static_cast<double>(x.cnt) * gp64[x.n] / static_cast<double>(C);
                TsEntry& t = te[total++];
                t.off = static_cast<uint32_t>((static_cast<int64_t>(col) * pitch + x.n - 1) * 4);
                t.wlo = static_cast<float>(a * cg), t.whi = static_cast<float>(bw * cg);
                t.n = x.n;
                ++h.cnt;
                csum += x.cnt;
                h.maxn = std::max(h.maxn, x.n);
            }
            h.rscale = static_cast<float>(csum / static_cast<double>(C));
        }
        if (ok) hdr[1] = total;
        // would ppo_tvf_targets (PPO_TVF_AUTO) take the slab kernel for these tables?  (shape conditions of
        // launch_tvf_tslab; buffer alignment is the caller's business)
        tslab_ok = ok && E >= 3 && E <= 255 && nmax <= 1023 && (V == 32 || V == 64 || V == 128) && K % 32 == 0 &&
                   steps <= 512;
    }
    if (!tslab_ok) want_generic = true;  // some kernel must be able to run: never hand out tables nothing can use
    if (want_generic) memset(blob, 0, static_cast<size_t>(offs[15]) * 4);
    if (want_cslab) memset(blob + cs_off, 0, static_cast<size_t>(off - cs_off) * 4);
    else memset(blob + cs_off, 0, kCsHeaderWords * 4);
    int32_t* ent_n = blob + offs[0];
    float* ent_cnt = reinterpret_cast<float*>(blob + offs[1]);
    int32_t* ent_idx = blob + offs[2];
    float* ent_wlo = reinterpret_cast<float*>(blob + offs[3]);
    float* ent_whi = reinterpret_cast<float*>(blob + offs[4]);
    float* ent_g = reinterpret_cast<float*>(blob + offs[5]);
    int32_t* tail_idx = blob + offs[6];
    float* tail_wlo = reinterpret_cast<float*>(blob + offs[7]);
    float* tail_whi = reinterpret_cast<float*>(blob + offs[8]);
    float* tail_cwlo = reinterpret_cast<float*>(blob + offs[9]);
    float* tail_cwhi = reinterpret_cast<float*>(blob + offs[10]);
    float* gp = reinterpret_cast<float*>(blob + offs[11]);
    int32_t* tT_idx = blob + offs[12];
    float* tT_wlo = reinterpret_cast<float*>(blob + offs[13]);
    float* tT_whi = reinterpret_cast<float*>(blob + offs[14]);
    if (want_generic) {
        for (int k = 0; k < K; ++k) {
            const Entry* row = ent.data() + static_cast<size_t>(k) * C;
            for (int e = 0; e < E; ++e) {
                const int64_t i = static_cast<int64_t>(e) * K + k;
                if (e < ent_count[k]) {
                    const Entry& x = row[e];
                    ent_n[i] = x.n, ent_cnt[i] = x.cnt;
                    ent_idx[i] = x.l.lo | (x.l.hi << 16);
                    ent_wlo[i] = x.l.wlo, ent_whi[i] = x.l.whi;
                    ent_g[i] = gpow[x.n];
                } else {
                    ent_g[i] = gpow[0];  // padding slots: n = 0, weight 0, gamma^0 like the NumPy builder
                }
            }
        }
        for (int n = 0; n <= nmax; ++n) gp[n] = gpow[n];
    }

    // ---- tails: horizon h_k - m for m = N - t; used for m <= largest entry of the horizon.  One pass per horizon
    // writes the horizon-major folded tables and the interleaved records, a second one the length-major tables of the
    // generic kernel.
    std::vector<unsigned char>& beyond_tab = sc.beyond;
    std::vector<uint32_t>& tail_mask = sc.tail_mask;
    if (want_cslab) beyond_tab.assign(static_cast<size_t>(K) * (nmax + 1), 0);
    tail_mask.assign(K, 0);
    {
        const double inv_c = 1.0 / static_cast<double>(C);
        // pass 1, horizon-major (every stream sequential): multiplicities beyond m, folded weights, the horizon-major
        // tables and the interleaved records; the plain lerp weights are parked for pass 2
        std::vector<float>& park = sc.park;  // [K][tail_pitch] x {wlo, whi}
        if (want_generic) park.resize(static_cast<size_t>(K) * tpitch * 2);
        for (int k = 0; k < K; ++k) {
            const Entry* row = ent.data() + static_cast<size_t>(k) * C;
            const int ne = ent_count[k], maxn = k_nmax[k];
            double beyond = 0.0;  // multiplicity of the entries that run past m
            for (int i = 0; i < ne; ++i) beyond += row[i].cnt;
            int e = 0;
            uint32_t mask = 0;
            const Lerp* slots = tc.slots.data() + k * cpitch;
            const unsigned char* oks = tc.ok.data() + k * cpitch;
            const int64_t dst0 = static_cast<int64_t>(k) * tpitch;
            unsigned bad = 0;
            for (int m = 1; m <= maxn; ++m) {
                while (e < ne && row[e].n <= m) beyond -= row[e++].cnt;
                bad |= oks[m] ^ 1u;
                const Lerp l = slots[m];
                const int64_t dst = dst0 + m;
                const int32_t idx = l.lo | (l.hi << 16);
                const double scale = beyond * static_cast<double>(gpow[m]);
                const float cwlo = static_cast<float>(static_cast<double>(l.wlo) * scale);
                const float cwhi = static_cast<float>(static_cast<double>(l.whi) * scale);
                if (want_generic) {
                    tT_idx[dst] = idx, tT_wlo[dst] = cwlo, tT_whi[dst] = cwhi;
                    park[2 * dst] = l.wlo, park[2 * dst + 1] = l.whi;
                }
                const float wl = static_cast<float>(static_cast<double>(cwlo) * inv_c);
                const float wh = static_cast<float>(static_cast<double>(cwhi) * inv_c);
                t4[4 * dst] = idx;
                memcpy(t4 + 4 * dst + 1, &wl, 4), memcpy(t4 + 4 * dst + 2, &wh, 4);
                if (want_cslab) {
                    beyond_tab[static_cast<size_t>(k) * (nmax + 1) + m] = static_cast<unsigned char>(std::min(beyond, 255.0));
                    // value columns of the final row this tail bootstrap reads (column slab kernel: job masks)
                    mask |= (cwlo != 0.f && l.lo < V) ? 1u << (l.lo / 32 & 31) : 0u;
                    mask |= (cwhi != 0.f && l.hi < V) ? 1u << (l.hi / 32 & 31) : 0u;
                }
            }
            if (bad) {
                set_error("index %d is out of bounds for axis 0 with size %d", V, V);
                return PPO_ERR_UNSUPPORTED;
            }
            tail_mask[k] = mask;
        }
        // pass 2: the length-major tables of the generic kernel, transposed in blocks of 16 horizons (one cache line
        // per table and tail length)
        for (int kb = 0; kb < K && want_generic; kb += 16) {
            const int ke = std::min(kb + 16, K);
            int bmax = 0;
            for (int k = kb; k < ke; ++k) bmax = std::max(bmax, k_nmax[k]);
            for (int m = 1; m <= bmax; ++m)
                for (int k = kb; k < ke; ++k) {
                    // (rows beyond the horizon's largest n-step hold zeros in both layouts)
                    const int64_t i = static_cast<int64_t>(m) * K + k, src = static_cast<int64_t>(k) * tpitch + m;
                    tail_idx[i] = tT_idx[src], tail_cwlo[i] = tT_wlo[src], tail_cwhi[i] = tT_whi[src];
                    tail_wlo[i] = m <= k_nmax[k] ? park[2 * src] : 0.f;
                    tail_whi[i] = m <= k_nmax[k] ? park[2 * src + 1] : 0.f;
                }
        }
    }

    {
        // Column slab kernel (tvf_cslab.cu): per horizon the entries sorted by their shift against the thread's own
        // rows (then ascending n) with the aligned row group ("bin") they gather; per job (four consecutive horizons)
        // the column groups its gathers and its tail bootstraps touch, and per column group the number of jobs that
        // touch it (the kernel refills a column group of the slab with the next env as soon as that many jobs have
        // finished).
        int32_t* hdr = blob + cs_off;
        hdr[0] = kCsConstMagic, hdr[1] = -1, hdr[2] = static_cast<int32_t>(steps), hdr[3] = V;
        const int jobs = (K + 3) / 4;
        hdr[4] = jobs;
        bool ok = want_cslab && K <= kCsMaxK && K % 4 == 0 && ek <= kCsMaxEntries && V % 32 == 0 && V <= 128 &&
                  steps % 128 == 0 && steps <= 256;
        CsHorizon* hzw = reinterpret_cast<CsHorizon*>(hdr + kCsHeaderWords);
        uint32_t* jobw = reinterpret_cast<uint32_t*>(hzw + K);
        int32_t* vhw = reinterpret_cast<int32_t*>(jobw + pad(jobs));  // [128] value horizons (the kernel interpolates with them)
        CsEntry* ce = reinterpret_cast<CsEntry*>(vhw + 128);
        uint32_t* nib = reinterpret_cast<uint32_t*>(ce + ek + 1);  // multiplicities of the n-steps beyond m, 4 bits each
        for (int j = 0; j < V && j < 128 && want_cslab; ++j)
            vhw[j] = static_cast<int32_t>(std::min<int64_t>(std::max<int64_t>(value_horizons[j], 0), INT32_MAX));
        auto group_of = [&](int col) { return col < V ? col / 32 : (col == V ? kCsSpecial : -1); };
        int total = 0;
        int arrivals[kCsGroups] = {0, 0, 0, 0, 0};
        // Tail bootstraps (rows whose n-step window runs off the rollout) without table reads from global memory:
        //   tail[t, k] = cnt_k(m) / C * gamma^m * F(h_k - m),  m = N - t,  F = the final value curve interpolated
        // cnt_k(m) (samples of the horizon with n > m) comes from a 4-bit table; F either from a table over integer
        // horizons 1..xmax that the kernel interpolates once per env (mode 1), or, when every h_k - m of the horizon
        // lies in one interpolation slot, as a line in m (mode 2).  Anything else keeps the record tables (mode 3).
        const bool plain_axis = !use_log_interpolation && C <= 15 && nmax <= 128 && V <= 128;
        bool strictly_up = true;
        for (int j = 1; j < V; ++j) strictly_up = strictly_up && value_horizons[j] > value_horizons[j - 1];
        const int64_t xcap = 1535;  // table horizons the kernel has shared memory for
        int64_t xmax = 0, nibbles = 0;
        for (int k = 0; k < K && ok; ++k) {
            const Entry* row = ent.data() + static_cast<size_t>(k) * C;
            int cnt[4] = {0, 0, 0, 0}, maxn = 0;
            const int beg = total;
            uint32_t mask = 0;
            for (int shift = 0; shift < 4 && ok; ++shift)
                for (int e = 0; e < ent_count[k] && ok; ++e) {
                    const Entry& x = row[e];
                    if ((x.n - 1) % 4 != shift) continue;
                    double a = x.l.wlo, bw = x.l.whi;
                    int col;
                    if (a == 0.0 && bw == 0.0) col = V, a = 1.0;
                    else if (x.l.hi == x.l.lo) col = x.l.lo, bw = 0.0;
                    else if (x.l.hi == x.l.lo + 1) col = x.l.lo;
                    else { ok = false; break; }
                    if (x.n > 1023) { ok = false; break; }
                    const double cg = static_cast<double>(x.cnt) * gp64[x.n] / static_cast<double>(C);
                    const int bin = (x.n - 1) / 4;
                    CsEntry& t = ce[total++];
                    const uint32_t offb = static_cast<uint32_t>((static_cast<int64_t>(col) * steps + 4 * bin) * 4);
                    t.w = offb | (static_cast<uint32_t>(x.n) << 18);
                    t.wlo = static_cast<float>(a * cg), t.whi = static_cast<float>(bw * cg), t.pad = 0;
                    ++cnt[shift], maxn = std::max(maxn, x.n);
                    if (t.wlo != 0.f && group_of(col) >= 0) mask |= 1u << group_of(col);
                    if (t.whi != 0.f && group_of(col + 1) >= 0) mask |= 1u << group_of(col + 1);
                }
            for (int shift = 0; shift < 4; ++shift)
                if (cnt[shift] > 255) ok = false;
            // value columns of the final row the tail bootstraps of this horizon read (rows with N - t < maxn)
            mask |= tail_mask[k] & ((1u << kCsSpecial) - 1);
            // tail mode of the horizon
            const int64_t h = required_horizons[k];
            int mode = maxn > 0 ? kCsTailGeneric : kCsTailNone;
            int32_t w3 = 0;
            float w4 = 0.f;
            if (maxn > 0 && plain_axis && strictly_up) {
                // one slot for all m: vh[s-1] <= h - maxn and h - 1 <= vh[s]
                const int s_hi = static_cast<int>(std::lower_bound(value_horizons, value_horizons + V, h - 1) - value_horizons);
                if (s_hi >= 1 && s_hi < V && value_horizons[s_hi - 1] <= h - maxn && h - value_horizons[s_hi - 1] < (1 << 22)) {
                    mode = kCsTailLinear;
                    w3 = s_hi | static_cast<int32_t>((h - value_horizons[s_hi - 1]) << 8);
                    w4 = static_cast<float>(1.0 / static_cast<double>(value_horizons[s_hi] - value_horizons[s_hi - 1]));
                } else if (h - 1 <= xcap && h - 1 <= value_horizons[V - 1] && value_horizons[0] <= 1) {
                    mode = kCsTailTable;
                    w3 = static_cast<int32_t>(h - steps);
                    xmax = std::max<int64_t>(xmax, h - 1);
                }
            }
            CsHorizon& hz = hzw[k];
            hz.w0 = static_cast<uint32_t>(beg) | (static_cast<uint32_t>(maxn) << 11) | (static_cast<uint32_t>(mode) << 21) |
                    (h > 0 ? 1u << 23 : 0u);
            hz.w1 = static_cast<uint32_t>(cnt[0]) | (static_cast<uint32_t>(cnt[1]) << 8) | (static_cast<uint32_t>(cnt[2]) << 16) |
                    (static_cast<uint32_t>(cnt[3]) << 24);
            hz.w2 = static_cast<uint32_t>(nibbles), hz.w3 = w3, hz.w4 = w4;
            if (mode == kCsTailTable || mode == kCsTailLinear) {
                const unsigned char* bt = beyond_tab.data() + static_cast<size_t>(k) * (nmax + 1);
                for (int m = 1; m <= maxn; ++m) {
                    const int64_t ni = nibbles + (m - 1);
                    nib[ni >> 3] |= static_cast<uint32_t>(bt[m] & 15) << ((ni & 7) * 4);
                }
                nibbles += maxn;
            }
            jobw[k / 4] |= mask;
        }
        hdr[10] = static_cast<int32_t>(xmax);
        hdr[11] = static_cast<int32_t>((nibbles + 7) / 8 + 2);  // words of the multiplicity table the kernel copies
        if (ok) {
            for (int j = 0; j < jobs; ++j)
                for (int g = 0; g < kCsGroups; ++g)
                    if (jobw[j] >> g & 1u) ++arrivals[g];
            for (int g = 0; g < kCsGroups; ++g) hdr[5 + g] = arrivals[g];
            ce[total] = CsEntry{0u, 0.f, 0.f, 0u};  // the kernel fetches one entry ahead
            hdr[1] = total;
        }
    }
    layout->sections = (want_generic ? PPO_TVF_TABLES_GENERIC : 0) | (want_tslab ? PPO_TVF_TABLES_TSLAB : 0) |
                       (want_cslab ? PPO_TVF_TABLES_CSLAB : 0);
    layout->reserved = 0;
    layout->tail_pitch = static_cast<int32_t>(tpitch);
    layout->cslab_offset_words = cs_off;
    layout->gamma = gamma;
    layout->n_entries = E, layout->nmax = nmax;
    layout->total_words = off;
    layout->inv_c = static_cast<float>(1.0 / static_cast<double>(C));
    return PPO_OK;
}
