// TVF targets, TMA slab kernel (the fast path of ppo_tvf_targets).
//
// One persistent CTA per SM walks a contiguous range of envs.  For one env the kernel needs the
// value curve rows V[t+n, a, :] for every n-step n of every horizon: a window of up to nmax rows
// ahead of the output row t.  Shared memory holds `menv` env slabs of cpe * kRC value rows each; row r
// of an env always lives at the same offset of its slab (no modulo inside an env, so the four rows
// of a time group are immediate offsets of one address).  Rows are streamed ONCE from HBM by TMA
// in row-chunks of kRC rows, in the order they are consumed (env-major, time ascending); a chunk of
// the next env is fetched as soon as the time loop has passed the same chunk of the current env,
// so HBM loads overlap the evaluation of earlier rows.  Warp roles (issue priority on an SM
// sub-partition goes to the highest warp id, so the roles that feed everyone else get the top ids):
//   warps 0-15      consumers: lane <-> horizon k (entry constants live in registers for the whole
//                   kernel); a warp evaluates groups of 4 consecutive time rows; per (4 rows, k, entry)
//                   one 16-byte prefix read and eight slab reads; results stream to HBM as 128-byte rows
//   warps 16-19     prefix-sum ("scan") warps: PT[n][t] = S_n(t,a) and the bootstrap limit of row t for
//                   blocks of 32 rows, through a ring of 4 tables (warp w owns ring slot w), one lane per
//                   time row, float32 accumulation in reference order; rows are continued past a
//                   terminal / past the rollout end (constant), so the reward part of every entry is a
//                   plain table read.  The reward / done columns of an env are staged in shared memory
//                   two envs ahead with 4-byte cp.async copies (no registers, no stall on their latency)
//   warp 20         TMA producer (one elected lane)
// Every hand-off is an mbarrier ring, so no role ever waits for a whole-CTA barrier.
// HBM traffic = value rows read once + targets written once + the reward/done columns.
#include <stdlib.h>

#include <type_traits>

#include "tma.cuh"
#include "tvf_common.cuh"

namespace ppo {

constexpr int kRC = 64;            // value rows per TMA chunk
constexpr int kTB = 32;            // time rows per prefix-table block
constexpr int kPRing = 4;          // prefix-table ring slots (block g of the env-major sequence uses slot g % kPRing)
constexpr int kColRing = 4;        // reward/done column ring (envs)
constexpr int kTG = 4;             // consecutive time rows evaluated together by a consumer warp
constexpr int kConsumerWarps = 16;
constexpr int kRingScanWarps = 2;  // warp w fills blocks g = w, w + 2, ...
constexpr int kRingThreads = (1 + kRingScanWarps + kConsumerWarps) * 32;
constexpr int kEMax = 8;           // entries per horizon kept in registers
constexpr int kMaxSteps = 512;

struct RingParams {
    TvfParams p;
    int nslot, menv, cpe, nblk, npad, kb;
    int dbg;  // development switches (PPO_TVF_DEBUG): 1 no stores, 2 no TMA loads, 4 no scan loop, 8 no entry loop
    uint32_t slab_bytes;
    uint32_t pt_buf_bytes;      // one prefix table: (nmax + 1) rows of kTB + 4 floats
    int64_t env_base, env_rem;  // even split of the env axis over the grid
    const float* tail_wlo;      // [nmax + 1, K] weights of the bootstrap from the final row (see launch)
    const float* tail_whi;
    // shared-memory offsets (bytes)
    uint32_t off_p, off_lim, off_colr, off_cold, off_dm, off_gpow, off_bar, smem_total;
};

// explicit shared-window loads: 32-bit address register + immediate offset
template <int IMM>
__device__ __forceinline__ float lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ float lds0(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds_s32x4(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct ConsumerCtx {
    int cw, lane, nenv;
    int64_t env0;
    uint32_t ring_a, p_a, lim_a;
    uint64_t *full, *empty, *pfull, *pempty;
};

// Entry constants of one lane (= one horizon), registers for the whole kernel.
template <int E>
struct LaneEntries {
    int soff[E];    // n * (kTB + 4) * 4: byte offset of prefix row n; also the unit the row limits are stored in
    int voff[E];    // (n - 1) * row bytes + lo * 4: byte offset of the bootstrap pair relative to slab row t
    float cnt[E];   // multiplicity
    float wlo[E];   // multiplicity * gamma^n * lerp weights
    float whi[E];
    int nmin[E];    // smallest soff of the warp per slot (a slot past every row limit skips its slab reads)
};

// One group of kTG consecutive time rows for 32 horizons.  TAIL: some window reaches past the rollout.
template <int V, int E, bool TAIL>
__device__ __forceinline__ void eval_group(const RingParams& q, const LaneEntries<E>& L, uint32_t pt, uint32_t lim_addr,
                                           uint32_t vrow, uint32_t cap, uint32_t fin, int t, int rows_left, int k,
                                           bool k_ok, float* o) {
    constexpr int RB = V * 4;
    constexpr int TCP4 = (kTB + 4) * 4;
    const TvfParams& p = q.p;
    const int N = p.steps, K = p.n_req, nmax = p.tab.nmax;
    // lim[u] = min(N - t, first terminal) in prefix-row units: an entry bootstraps inside the rollout iff soff <= lim
    const int4 lim = lds_s32x4(lim_addr);
    const int lim_max = max(max(lim.x, lim.y), max(lim.z, lim.w));

    float tw0[kTG], tw1[kTG];
    int tix[kTG];
    if (TAIL) {
        // rows that run off the rollout bootstrap from the final row with host-folded weights
        // (sum of the multiplicities of the entries with n' > m, times gamma^m, times the lerp weight);
        // the table reads are issued first so their L2 latency hides behind the entry loop
#pragma unroll
        for (int u = 0; u < kTG; ++u) {
            const int m = N - (t + u);
            tw0[u] = 0.f, tw1[u] = 0.f, tix[u] = 0;
            if (m >= 1 && m < nmax && k_ok) {
                tw0[u] = __ldg(q.tail_wlo + m * K + k);
                tw1[u] = __ldg(q.tail_whi + m * K + k);
                tix[u] = __ldg(p.tab.tail_idx + m * K + k);
            }
        }
    }

    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (!(q.dbg & 8))
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const float4 s = lds_f32x4(pt + static_cast<uint32_t>(L.soff[i]));
        a0 = fmaf(L.cnt[i], s.x, a0), a1 = fmaf(L.cnt[i], s.y, a1);
        a2 = fmaf(L.cnt[i], s.z, a2), a3 = fmaf(L.cnt[i], s.w, a3);
        if (L.nmin[i] <= lim_max) {
            const uint32_t va = vrow + static_cast<uint32_t>(L.voff[i]);
            float l0, h0, l1, h1, l2, h2, l3, h3;
            if (!TAIL) {
                l0 = lds<0>(va), h0 = lds<4>(va), l1 = lds<RB>(va), h1 = lds<RB + 4>(va);
                l2 = lds<2 * RB>(va), h2 = lds<2 * RB + 4>(va), l3 = lds<3 * RB>(va), h3 = lds<3 * RB + 4>(va);
            } else {
                // clamp every row's address into the slab (the value is discarded by the limit test)
                const uint32_t v0 = min(va, cap), v1 = min(va + RB, cap);
                const uint32_t v2 = min(va + 2 * RB, cap), v3 = min(va + 3 * RB, cap);
                l0 = lds<0>(v0), h0 = lds<4>(v0), l1 = lds<0>(v1), h1 = lds<4>(v1);
                l2 = lds<0>(v2), h2 = lds<4>(v2), l3 = lds<0>(v3), h3 = lds<4>(v3);
            }
            if (L.soff[i] <= lim.x) a0 = fmaf(h0, L.whi[i], fmaf(l0, L.wlo[i], a0));
            if (L.soff[i] <= lim.y) a1 = fmaf(h1, L.whi[i], fmaf(l1, L.wlo[i], a1));
            if (L.soff[i] <= lim.z) a2 = fmaf(h2, L.whi[i], fmaf(l2, L.wlo[i], a2));
            if (L.soff[i] <= lim.w) a3 = fmaf(h3, L.whi[i], fmaf(l3, L.wlo[i], a3));
        }
    }
    if (TAIL) {
        float* accs[4] = {&a0, &a1, &a2, &a3};
        const int lims[4] = {lim.x, lim.y, lim.z, lim.w};
#pragma unroll
        for (int u = 0; u < kTG; ++u) {
            const int m = N - (t + u);
            // alive to the end of the rollout <=> the row limit equals the rollout limit
            if ((tw0[u] != 0.f || tw1[u] != 0.f) && lims[u] == m * TCP4) {
                *accs[u] = fmaf(lds0(fin + (tix[u] & 0xffff) * 4u), tw0[u], *accs[u]);
                *accs[u] = fmaf(lds0(fin + (tix[u] >> 16) * 4u), tw1[u], *accs[u]);
            }
        }
    }
    if (k_ok && !(q.dbg & 1)) {
        const float inv_c = p.tab.inv_c;
        const int64_t ostride = p.envs * static_cast<int64_t>(K);
        __stcs(o, a0 * inv_c);
        if (rows_left > 1) __stcs(o + ostride, a1 * inv_c);
        if (rows_left > 2) __stcs(o + 2 * ostride, a2 * inv_c);
        if (rows_left > 3) __stcs(o + 3 * ostride, a3 * inv_c);
    }
}

// Consumer warp body, E = number of live entry slots of this warp's 32 horizons (compile time, so
// the evaluation of one 4-row group is straight-line code the scheduler can overlap freely).
template <int V, int E>
__device__ __forceinline__ void consume(const RingParams& q, const ConsumerCtx& c) {
    constexpr int RB = V * 4;
    constexpr int TCP4 = (kTB + 4) * 4;
    const TvfParams& p = q.p;
    const int N = p.steps, K = p.n_req, nmax = p.tab.nmax;
    const int lane = c.lane;
    const int kb = c.cw % q.kb, tg = c.cw / q.kb, ntg = kConsumerWarps / q.kb;
    const int k = kb * 32 + lane;
    const bool k_ok = k < K;

    // entry constants; dead slots keep zero weights and read prefix row 0 (zeros)
    LaneEntries<E> L;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        L.soff[e] = 0, L.voff[e] = lane * 4, L.cnt[e] = 0.f, L.wlo[e] = 0.f, L.whi[e] = 0.f;
        int mine = INT32_MAX;
        if (k_ok && e < p.tab.n_entries) {
            const float cc = __ldg(p.tab.ent_cnt + e * K + k);
            if (cc > 0.f) {
                const int n = __ldg(p.tab.ent_n + e * K + k);
                const int idx = __ldg(p.tab.ent_idx + e * K + k);
                int lo = idx & 0xffff;
                const int hi = idx >> 16;
                float a = __ldg(p.tab.ent_wlo + e * K + k), bw = __ldg(p.tab.ent_whi + e * K + k);
                if (hi == lo) {  // single-slot value: read (lo, lo + 1) with a zero weight on the unused one
                    if (lo + 1 < V) bw = 0.f;
                    else lo -= 1, bw = a, a = 0.f;
                }
                const float cg = cc * __ldg(p.tab.ent_g + e * K + k);  // multiplicity * gamma^n
                L.soff[e] = n * TCP4;
                L.cnt[e] = cc, L.wlo[e] = a * cg, L.whi[e] = bw * cg;
                // entries without a bootstrap (both weights zero) never read the slab
                if (a != 0.f || bw != 0.f) L.voff[e] = (n - 1) * RB + lo * 4, mine = n * TCP4;
            }
        }
        L.nmin[e] = __reduce_min_sync(0xffffffffu, mine);
    }

    const int64_t ostride = p.envs * static_cast<int64_t>(K);
    uint32_t blk = 0;  // global prefix-block counter (env-major): ring slot blk % kPRing, use blk / kPRing

    for (int e = 0; e < c.nenv; ++e) {
        const int64_t env = c.env0 + e;
        const int slab = e % q.menv;
        const uint32_t fpar = static_cast<uint32_t>((e / q.menv) & 1);
        const uint32_t slab_a = c.ring_a + static_cast<uint32_t>(slab) * q.slab_bytes;
        const uint32_t cap = slab_a + static_cast<uint32_t>(N - 1) * RB + (V - 2) * 4u;  // last pair of the final row
        const uint32_t fin = slab_a + static_cast<uint32_t>(N - 1) * RB;
        const int slot0 = slab * q.cpe;
        float* out_env = p.out + env * K + k;
        int waited = 0, released = 0;  // chunks of this env
        for (int j = 0; j < q.nblk; ++j, ++blk) {
            const uint32_t ps = blk % kPRing;
            mbar_wait(&c.pfull[ps], (blk / kPRing) & 1);

            const int t_hi = min(N, j * kTB + kTB - 1 + nmax);
            const int cneed = (t_hi - 1) / kRC;
            while (waited <= cneed) {
                mbar_wait(&c.full[slot0 + waited], fpar);
                ++waited;
            }

            const int rows = min(kTB, N - j * kTB);
            const uint32_t pt_buf = c.p_a + ps * q.pt_buf_bytes;
            const uint32_t lim_buf = c.lim_a + ps * (kTB * 4u);
            for (int tl = tg * kTG; tl < rows; tl += ntg * kTG) {
                const int t = j * kTB + tl;
                const uint32_t pt = pt_buf + static_cast<uint32_t>(tl) * 4u;
                const uint32_t lim_addr = lim_buf + static_cast<uint32_t>(tl) * 4u;
                const uint32_t vrow = slab_a + static_cast<uint32_t>(t) * RB;
                float* o = out_env + static_cast<int64_t>(t) * ostride;
                if (t + (kTG - 1) + nmax > N)  // warp-uniform
                    eval_group<V, E, true>(q, L, pt, lim_addr, vrow, cap, fin, t, rows - tl, k, k_ok, o);
                else
                    eval_group<V, E, false>(q, L, pt, lim_addr, vrow, cap, fin, t, rows - tl, k, k_ok, o);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&c.pempty[ps]);
            const int rel_end = (j == q.nblk - 1) ? q.cpe : min(q.cpe, ((j + 1) * kTB) / kRC);
            while (released < rel_end) {
                if (lane == 0) mbar_arrive(&c.empty[slot0 + released]);
                ++released;
            }
        }
    }
}

// V = value horizons (row of V floats)
template <int V>
__global__ void __launch_bounds__(kRingThreads, 1)
tvf_ring_kernel(const RingParams q, const __grid_constant__ CUtensorMap vmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RB = V * 4;               // bytes per value row
    constexpr uint32_t CHUNK = kRC * RB;    // bytes per TMA chunk
    constexpr int TCP = kTB + 4;            // pitch (floats) of one prefix-table row
    const TvfParams& p = q.p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.steps, K = p.n_req, nmax = p.tab.nmax;

    unsigned char* ring = smem;
    float* PT = reinterpret_cast<float*>(smem + q.off_p);          // [kPRing][nmax + 1][TCP]
    int* lim_s = reinterpret_cast<int*>(smem + q.off_lim);         // [kPRing][kTB]
    float* col_r = reinterpret_cast<float*>(smem + q.off_colr);    // [kColRing][npad] rewards, zero past N
    uint32_t* col_d = reinterpret_cast<uint32_t*>(smem + q.off_cold);  // [kColRing][npad] raw done words
    uint32_t* dm_s = reinterpret_cast<uint32_t*>(smem + q.off_dm);     // [scan warp][npad / 32] done bitmask
    float* gpow_s = reinterpret_cast<float*>(smem + q.off_gpow);   // [nmax + 8]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + q.off_bar);
    uint64_t* empty = full + q.nslot;
    uint64_t* pfull = empty + q.nslot;
    uint64_t* pempty = pfull + kPRing;
    uint64_t* cfull = pempty + kPRing;
    uint64_t* cempty = cfull + kColRing;

    const int64_t b = blockIdx.x;
    const int nenv = static_cast<int>(q.env_base + (b < q.env_rem ? 1 : 0));
    const int64_t env0 = b * q.env_base + (b < q.env_rem ? b : q.env_rem);

    if (threadIdx.x == 0) {
        for (int i = 0; i < q.nslot; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kConsumerWarps);
        }
        for (int i = 0; i < kPRing; ++i) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pempty[i], kConsumerWarps);
        }
        for (int i = 0; i < kColRing; ++i) {
            mbar_init(&cfull[i], 32);              // one cp.async-completion arrival per copying lane
            mbar_init(&cempty[i], kRingScanWarps);
        }
        mbar_fence_init();
        tma_prefetch_desc(&vmap);
    }
    // Slab, prefix and column memory start finite / zero: entries whose bootstrap is switched off still read
    // the slab, prefix row 0 (S_0 = 0) is never written again, and reward rows past N stay zero.
    for (uint32_t i = threadIdx.x * 16u; i < q.off_bar; i += blockDim.x * 16u)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = threadIdx.x; i <= nmax; i += blockDim.x) gpow_s[i] = __ldg(p.tab.gpow + i);
    // make the generic-proxy zero fill visible to the async proxy before any TMA write lands
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == kConsumerWarps + kRingScanWarps) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int slot = 0;
            uint32_t par = 0;
            for (int e = 0; e < nenv; ++e) {
                const int env = static_cast<int>(env0 + e);
                for (int c = 0; c < q.cpe; ++c) {
                    mbar_wait(&empty[slot], par ^ 1);
                    if (q.dbg & 2) {
                        mbar_arrive(&full[slot]);
                    } else {
                        // a box that reaches past row N is zero-filled by the TMA unit and still counts in full
                        mbar_arrive_expect_tx(&full[slot], CHUNK);
                        tma_load_3d(ring + static_cast<size_t>(slot) * CHUNK, &vmap, &full[slot], 0, env, c * kRC + 1);
                    }
                    if (++slot == q.nslot) slot = 0, par ^= 1;
                }
            }
        }
    } else if (warp >= kConsumerWarps) {
        // ------------------------------------------------------------------ prefix-sum warps
        const int w = warp - kConsumerWarps;
        const int dm_words = q.npad / 32;
        uint32_t* my_dm = dm_s + static_cast<size_t>(w) * dm_words;  // bit r % 32 of word r / 32
        const int total = nenv * q.nblk;
        const int ncols = (N + 31) / 32;  // 32-row groups of a column

        // Stage the reward / done columns of local env `ec` (this warp's duty when ec % scan warps == w).
        auto stage_cols = [&](int ec) {
            const int cs = ec % kColRing;
            mbar_wait(&cempty[cs], ((ec / kColRing) & 1) ^ 1);
            const int64_t env = env0 + ec;
            float* dr = col_r + static_cast<size_t>(cs) * q.npad;
            uint32_t* dd = col_d + static_cast<size_t>(cs) * q.npad;
            for (int i = 0; i < ncols; ++i) {
                const int row = i * 32 + lane;
                if (row < N) {
                    cp_async_4(dr + row, p.rewards + static_cast<int64_t>(row) * p.r_rs + env);
                    if (p.done_kind == PPO_DONE_U8) {
                        const uintptr_t a = reinterpret_cast<uintptr_t>(p.dones) + static_cast<uint64_t>(row) * p.d_rs + env;
                        cp_async_4(dd + row, reinterpret_cast<const void*>(a & ~static_cast<uintptr_t>(3)));
                    } else if (p.done_kind == PPO_DONE_F32) {
                        cp_async_4(dd + row, static_cast<const float*>(p.dones) + static_cast<int64_t>(row) * p.d_rs + env);
                    }
                }
            }
            cp_async_mbar_arrive(&cfull[cs]);
        };
        for (int ec = 0; ec < 2 && ec < nenv; ++ec)
            if (ec % kRingScanWarps == w) stage_cols(ec);

        int e_cur = -1;
        const float* r_env = col_r;
        for (int g = w; g < total; g += kRingScanWarps) {
            const int e = g / q.nblk, j = g - e * q.nblk;
            // Every scan warp visits every env in order (even one in which it owns no block), so the column
            // ring's full / empty phases advance in lock step for all four warps.
            while (e_cur < e) {
                if (e_cur >= 0 && lane == 0) mbar_arrive(&cempty[e_cur % kColRing]);
                ++e_cur;
                const int ec = e_cur + 2;  // staged two envs ahead
                if (ec < nenv && ec % kRingScanWarps == w) stage_cols(ec);
                mbar_wait(&cfull[e_cur % kColRing], (e_cur / kColRing) & 1);
                if (e_cur == e) {
                    // done bitmask of the env (private copy)
                    const int cs = e % kColRing;
                    const uint32_t* dd = col_d + static_cast<size_t>(cs) * q.npad;
                    for (int i = 0; i < ncols; ++i) {
                        const int row = i * 32 + lane;
                        bool dn = false;
                        if (row < N) {
                            const uint32_t wd = dd[row];
                            if (p.done_kind == PPO_DONE_U8) {
                                const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p.dones) +
                                                                         static_cast<uint64_t>(row) * p.d_rs + (env0 + e));
                                dn = ((wd >> (8 * (a & 3u))) & 0xffu) != 0;
                            } else if (p.done_kind == PPO_DONE_F32) {
                                dn = __uint_as_float(wd) != 0.f;
                            }
                        }
                        const uint32_t word = __ballot_sync(0xffffffffu, dn);
                        if (lane == 0) my_dm[i] = word;
                    }
                    __syncwarp();
                    r_env = col_r + static_cast<size_t>(cs) * q.npad;
                }
            }

            const int ps = g % kPRing;
            float* myPT = PT + static_cast<size_t>(ps) * (q.pt_buf_bytes / 4);
            int* my_lim = lim_s + ps * kTB;
            mbar_wait(&pempty[ps], ((g / kPRing) & 1) ^ 1);
            {
                const int t = j * kTB + lane;  // rows past the rollout read zero rewards
                const float* r_col = r_env + t;
                float* col = myPT + lane;
                // distance to the first terminal at or after t (only terminals inside the n-step window matter)
                int first_done = kNoDone;
                {
                    const uint32_t* dm = my_dm + j;
                    uint32_t m = dm[0] & (0xffffffffu << lane);
                    const int words = (nmax + 31 + 31) / 32;
                    int base = -lane;
                    for (int wi = 0; wi < words; ++wi) {
                        if (first_done == kNoDone && m != 0u) first_done = base + __ffs(m) - 1;
                        base += 32;
                        m = (j + wi + 1 < ncols) ? dm[wi + 1] : 0u;
                    }
                }
                // S_n(t) = sum_{i<n} float32(r[t+i] * D_i), D_i = gamma^i while no terminal lies in t..t+i-1, else 0:
                // float32 accumulation in forward order like rl/returns_truncated.py:668; adding r * 0 is a no-op,
                // so the sum simply stops growing after the terminal step and the row continues as a constant.
                float acc = 0.f;
                for (int i0 = 0; i0 < ((q.dbg & 4) ? 1 : nmax); i0 += 8) {
                    float rr[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) rr[u] = r_col[i0 + u];
                    const float4 g0 = *reinterpret_cast<const float4*>(gpow_s + i0);
                    const float4 g1 = *reinterpret_cast<const float4*>(gpow_s + i0 + 4);
                    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (i0 + u < nmax) {
                            const float x = __fmul_rn(rr[u], gg[u]);
                            if (i0 + u <= first_done) acc = __fadd_rn(acc, x);
                            col[(i0 + u + 1) * TCP] = acc;
                        }
                    }
                }
                my_lim[lane] = min(N - t, first_done) * (TCP * 4);  // bootstrap limit of the row, in prefix-row bytes
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[ps]);
        }
        // drain the column ring protocol for envs after this warp's last block (other warps may still wait on phases)
        while (e_cur < nenv - 1) {
            if (e_cur >= 0 && lane == 0) mbar_arrive(&cempty[e_cur % kColRing]);
            ++e_cur;
            const int ec = e_cur + 2;
            if (ec < nenv && ec % kRingScanWarps == w) stage_cols(ec);
            mbar_wait(&cfull[e_cur % kColRing], (e_cur / kColRing) & 1);
        }
    } else {
        // ------------------------------------------------------------------ consumers
        ConsumerCtx c;
        c.cw = warp;
        c.lane = lane;
        c.ring_a = smem_u32(ring), c.p_a = smem_u32(PT), c.lim_a = smem_u32(lim_s);
        c.full = full, c.empty = empty, c.pfull = pfull, c.pempty = pempty;
        c.nenv = nenv, c.env0 = env0;
        // number of live entry slots of this warp's horizons (warp-uniform for the whole kernel)
        const int kq = (c.cw % q.kb) * 32 + lane;
        int mine = 0;
        if (kq < K)
            for (int e = 0; e < p.tab.n_entries && e < kEMax; ++e)
                if (__ldg(p.tab.ent_cnt + e * K + kq) > 0.f) mine = e + 1;
        const int ecnt = __reduce_max_sync(0xffffffffu, mine);
        switch (ecnt) {
            case 0:
            case 1: consume<V, 1>(q, c); break;
            case 2: consume<V, 2>(q, c); break;
            case 3: consume<V, 3>(q, c); break;
            case 4: consume<V, 4>(q, c); break;
            case 5: consume<V, 5>(q, c); break;
            case 6: consume<V, 6>(q, c); break;
            case 7: consume<V, 7>(q, c); break;
            default: consume<V, 8>(q, c); break;
        }
    }
}

static uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

template <int V>
static int launch_variant(RingParams& q, const CUtensorMap& vmap, int64_t grid, cudaStream_t stream) {
    PPO_CUDA(cudaFuncSetAttribute(tvf_ring_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(q.smem_total)));
    tvf_ring_kernel<V><<<static_cast<unsigned>(grid), kRingThreads, q.smem_total, stream>>>(q, vmap);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// Shared-memory plan; false when one env slab plus the tables do not fit.
static bool plan_smem(RingParams& q, int N, int V, int nmax, int max_smem) {
    const uint32_t row_bytes = static_cast<uint32_t>(V) * 4;
    q.cpe = (N + kRC - 1) / kRC;
    q.nblk = (N + kTB - 1) / kTB;
    // reward rows are read as r_col[t + i], t < nblk * kTB, i < nmax + 7, zero past N
    q.npad = ((q.nblk * kTB + nmax + 8 + 31) / 32 + 1) * 32;
    const uint32_t chunk_bytes = kRC * row_bytes;
    q.slab_bytes = static_cast<uint32_t>(q.cpe) * chunk_bytes;
    q.pt_buf_bytes = static_cast<uint32_t>(nmax + 1) * (kTB + 4) * 4u;
    const uint32_t p_bytes = kPRing * q.pt_buf_bytes;
    const uint32_t lim_bytes = kPRing * kTB * 4u;
    const uint32_t col_bytes = kColRing * static_cast<uint32_t>(q.npad) * 4u;  // each of rewards / done words
    const uint32_t dm_bytes = align_up(kRingScanWarps * (q.npad / 32) * 4u, 16);
    const uint32_t gp_bytes = align_up((nmax + 8) * 4u, 16);
    const uint32_t other = p_bytes + lim_bytes + 2 * col_bytes + dm_bytes + gp_bytes + 256;
    int menv = 0;
    while (menv < 4) {
        const uint64_t need = static_cast<uint64_t>(menv + 1) * q.slab_bytes + other +
                              (2ull * (menv + 1) * q.cpe + 2 * kPRing + 2 * kColRing) * 8;
        if (need > static_cast<uint64_t>(max_smem)) break;
        ++menv;
    }
    if (menv < 1) return false;
    q.menv = menv;
    q.nslot = menv * q.cpe;
    q.off_p = align_up(static_cast<uint32_t>(menv) * q.slab_bytes, 128);
    q.off_lim = q.off_p + p_bytes;
    q.off_colr = q.off_lim + lim_bytes;
    q.off_cold = q.off_colr + col_bytes;
    q.off_dm = q.off_cold + col_bytes;
    q.off_gpow = align_up(q.off_dm + dm_bytes, 16);
    q.off_bar = align_up(q.off_gpow + gp_bytes, 16);
    q.smem_total = q.off_bar + (2 * q.nslot + 2 * kPRing + 2 * kColRing) * 8;
    return q.smem_total <= static_cast<uint32_t>(max_smem);
}

int launch_tvf_ring(const TvfParams& p, const float* tail_wlo, const float* tail_whi, cudaStream_t stream,
                    bool probe_only) {
    const int N = p.steps, K = p.n_req, V = p.n_val, nmax = p.tab.nmax;
    auto unsupported = [&](const char* why) {
        set_error("ppo_tvf_targets: TMA slab kernel unavailable: %s", why);
        return PPO_ERR_UNSUPPORTED;
    };
    if (V != 32 && V != 64 && V != 128) return unsupported("value horizons must be 32, 64 or 128");
    if (K > 32 * kConsumerWarps) return unsupported("more than 512 required horizons");
    if (p.tab.n_entries > kEMax) return unsupported("more than 8 distinct n-steps per horizon");
    if (N > kMaxSteps) return unsupported("more than 512 steps");
    if (nmax > 256) return unsupported("n-steps above 256");
    if ((reinterpret_cast<uintptr_t>(p.vs) & 15) || (p.vs_es * 4) % 16 || (p.vs_rs * 4) % 16 || p.vs_es < V)
        return unsupported("value samples are not 16-byte aligned / strided");
    if (p.envs > INT32_MAX) return unsupported("too many envs");
    if (!tail_wlo || !tail_whi) return unsupported("folded tail weights missing");

    int dev = 0, max_smem = 0;
    PPO_CUDA(cudaGetDevice(&dev));
    PPO_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    RingParams q{};
    q.p = p;
    {
        const char* d = getenv("PPO_TVF_DEBUG");
        q.dbg = d ? atoi(d) : 0;
    }
    q.tail_wlo = tail_wlo, q.tail_whi = tail_whi;
    int kb = 1;
    while (kb * 32 < K) kb <<= 1;
    q.kb = kb;
    if (!plan_smem(q, N, V, nmax, max_smem))
        return unsupported("one env slab of value rows plus the prefix tables do not fit in shared memory");
    if (probe_only) return PPO_OK;

    CUtensorMap vmap;
    const uint64_t dims[3] = {static_cast<uint64_t>(V), static_cast<uint64_t>(p.envs), static_cast<uint64_t>(N) + 1};
    const uint64_t strides[2] = {static_cast<uint64_t>(p.vs_es) * 4, static_cast<uint64_t>(p.vs_rs) * 4};
    const uint32_t box[3] = {static_cast<uint32_t>(V), 1u, static_cast<uint32_t>(kRC)};
    int rc = encode_tensor_map(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.vs, dims, strides, box);
    if (rc != PPO_OK) return rc;

    int64_t grid = sm_count();
    if (grid > p.envs) grid = p.envs;
    q.env_base = p.envs / grid;
    q.env_rem = p.envs % grid;

    if (V == 128) return launch_variant<128>(q, vmap, grid, stream);
    if (V == 64) return launch_variant<64>(q, vmap, grid, stream);
    if (V == 32) return launch_variant<32>(q, vmap, grid, stream);
    return unsupported("no kernel variant");
}

}  // namespace ppo
