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
//   warps 16-17     prefix-sum ("scan") warps: PT[n][t] = S_n(t,a) and the bootstrap limit of row t for
//                   TC rows at a time, double buffered, one lane per time row, float32 accumulation in
//                   reference order; rows are continued past a terminal / past the rollout end
//                   (constant), so the reward part of every entry is a plain table read
//   warp 18         TMA producer (one elected lane)
// HBM traffic = value rows read once + targets written once + the reward/done columns.
#include <stdlib.h>

#include <type_traits>

#include "tma.cuh"
#include "tvf_common.cuh"

namespace ppo {

constexpr int kRC = 16;            // value rows per TMA chunk
constexpr int kTG = 4;             // consecutive time rows evaluated together by a consumer warp
constexpr int kConsumerWarps = 16;
constexpr int kRingScanWarps = 2;
constexpr int kRingThreads = (1 + kRingScanWarps + kConsumerWarps) * 32;
constexpr int kEMax = 8;           // entries per horizon kept in registers
constexpr int kMaxSteps = 512;     // reward/done columns are prefetched through registers
constexpr int kColRegs = kMaxSteps / 32;

struct RingParams {
    TvfParams p;
    int nslot, menv, cpe, ncht, npad, kb;
    int dbg;  // development switches (PPO_TVF_DEBUG): 1 no stores, 2 no TMA loads, 4 no scan loop, 8 no entry loop
    uint32_t slab_bytes;
    uint32_t pt_buf_bytes;      // one prefix table: (nmax + 1) rows of TC + 4 floats
    int64_t env_base, env_rem;  // even split of the env axis over the grid
    const float* tail_wlo;      // [nmax + 1, K] weights of the bootstrap from the final row (see launch)
    const float* tail_whi;
    // shared-memory offsets (bytes)
    uint32_t off_p, off_nd, off_rs, off_ds, off_gpow, off_bar, smem_total;
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
    int soff[E];    // n * (TC + 4) * 4: byte offset of prefix row n; also the unit the row limits are stored in
    int voff[E];    // (n - 1) * row bytes + lo * 4: byte offset of the bootstrap pair relative to slab row t
    float cnt[E];   // multiplicity
    float wlo[E];   // multiplicity * gamma^n * lerp weights
    float whi[E];
    int nmin[E];    // smallest soff of the warp per slot (a slot past every row limit skips its slab reads)
};

// One group of kTG consecutive time rows for 32 horizons.  TAIL: some window reaches past the rollout.
template <int V, int TC, int E, bool TAIL>
__device__ __forceinline__ void eval_group(const RingParams& q, const LaneEntries<E>& L, uint32_t pt, uint32_t lim_addr,
                                           uint32_t vrow, uint32_t cap, uint32_t fin, int t, int rows_left, int k,
                                           bool k_ok, float* o) {
    constexpr int RB = V * 4;
    constexpr int TCP4 = (TC + 4) * 4;
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
template <int V, int TC, int E>
__device__ __forceinline__ void consume(const RingParams& q, const ConsumerCtx& c) {
    constexpr int RB = V * 4;
    constexpr int TCP4 = (TC + 4) * 4;
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
    uint32_t pf_par0 = 0, pf_par1 = 0;

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
        for (int j = 0; j < q.ncht; ++j) {
            const int gsel = (e * q.ncht + j) & 1;
            if (gsel == 0) mbar_wait(&c.pfull[0], pf_par0), pf_par0 ^= 1;
            else mbar_wait(&c.pfull[1], pf_par1), pf_par1 ^= 1;

            const int t_hi = min(N, j * TC + TC - 1 + nmax);
            const int cneed = (t_hi - 1) / kRC;
            while (waited <= cneed) {
                mbar_wait(&c.full[slot0 + waited], fpar);
                ++waited;
            }

            const int rows = min(TC, N - j * TC);
            const uint32_t pt_buf = c.p_a + static_cast<uint32_t>(gsel) * q.pt_buf_bytes;
            const uint32_t lim_buf = c.lim_a + static_cast<uint32_t>(gsel * TC) * 4u;
            for (int tl = tg * kTG; tl < rows; tl += ntg * kTG) {
                const int t = j * TC + tl;
                const uint32_t pt = pt_buf + static_cast<uint32_t>(tl) * 4u;
                const uint32_t lim_addr = lim_buf + static_cast<uint32_t>(tl) * 4u;
                const uint32_t vrow = slab_a + static_cast<uint32_t>(t) * RB;
                float* o = out_env + static_cast<int64_t>(t) * ostride;
                if (t + (kTG - 1) + nmax > N)  // warp-uniform
                    eval_group<V, TC, E, true>(q, L, pt, lim_addr, vrow, cap, fin, t, rows - tl, k, k_ok, o);
                else
                    eval_group<V, TC, E, false>(q, L, pt, lim_addr, vrow, cap, fin, t, rows - tl, k, k_ok, o);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&c.pempty[gsel]);
            const int rel_end = (j == q.ncht - 1) ? q.cpe : min(q.cpe, ((j + 1) * TC) / kRC);
            while (released < rel_end) {
                if (lane == 0) mbar_arrive(&c.empty[slot0 + released]);
                ++released;
            }
        }
    }
}

// V = value horizons (row of V floats), TC = time rows per prefix chunk
template <int V, int TC>
__global__ void __launch_bounds__(kRingThreads, 1)
tvf_ring_kernel(const RingParams q, const __grid_constant__ CUtensorMap vmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int RB = V * 4;               // bytes per value row
    constexpr uint32_t CHUNK = kRC * RB;    // bytes per TMA chunk
    constexpr int TCP = TC + 4;             // pitch (floats) of one prefix-table row
    const TvfParams& p = q.p;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.steps, K = p.n_req, nmax = p.tab.nmax;

    unsigned char* ring = smem;
    float* PT = reinterpret_cast<float*>(smem + q.off_p);       // [2][nmax + 1][TCP]
    int* lim_s = reinterpret_cast<int*>(smem + q.off_nd);       // [2][TC]
    float* rs = reinterpret_cast<float*>(smem + q.off_rs);      // [scan warp][2][npad]
    unsigned char* ds = smem + q.off_ds;                        // [scan warp][2][npad / 32] done bitmask words
    float* gpow_s = reinterpret_cast<float*>(smem + q.off_gpow);  // [nmax + 8]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + q.off_bar);
    uint64_t* empty = full + q.nslot;
    uint64_t* pfull = empty + q.nslot;
    uint64_t* pempty = pfull + 2;

    const int64_t b = blockIdx.x;
    const int nenv = static_cast<int>(q.env_base + (b < q.env_rem ? 1 : 0));
    const int64_t env0 = b * q.env_base + (b < q.env_rem ? b : q.env_rem);

    if (threadIdx.x == 0) {
        for (int i = 0; i < q.nslot; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kConsumerWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&pfull[i], 1);
            mbar_init(&pempty[i], kConsumerWarps);
        }
        mbar_fence_init();
        tma_prefetch_desc(&vmap);
    }
    // Slab and prefix memory start finite: entries whose bootstrap is switched off still read them,
    // and prefix row 0 (S_0 = 0) is never written again.
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
                        if (++slot == q.nslot) slot = 0, par ^= 1;
                        continue;
                    }
                    mbar_arrive_expect_tx(&full[slot], CHUNK);
                    tma_load_3d(ring + static_cast<size_t>(slot) * CHUNK, &vmap, &full[slot], 0, env, c * kRC + 1);
                    if (++slot == q.nslot) slot = 0, par ^= 1;
                }
            }
        }
    } else if (warp >= kConsumerWarps) {
        // ------------------------------------------------------------------ prefix-sum warps
        const int s = warp - kConsumerWarps;
        float* my_rs = rs + static_cast<size_t>(s) * 2 * q.npad;
        const int dm_words = q.npad / 32;
        uint32_t* my_dm = reinterpret_cast<uint32_t*>(ds) + static_cast<size_t>(s) * 2 * dm_words;  // bit r%32 of word r/32
        float* myPT = PT + static_cast<size_t>(s) * (q.pt_buf_bytes / 4);
        int* my_lim = lim_s + s * TC;
        const int total = nenv * q.ncht;
        int cur = 0;       // which private column buffer holds the current env
        uint32_t use = 0;  // how many times this warp has filled its prefix buffer

        float cr[kColRegs];
        uint32_t cd[kColRegs / 4];
        auto load_cols = [&](int e_local) {
            const int64_t env = env0 + e_local;
#pragma unroll
            for (int i = 0; i < kColRegs / 4; ++i) cd[i] = 0;
#pragma unroll
            for (int i = 0; i < kColRegs; ++i) {
                const int row = i * 32 + lane;
                cr[i] = 0.f;
                if (row < N) {
                    cr[i] = __ldg(p.rewards + static_cast<int64_t>(row) * p.r_rs + env);
                    const uint32_t dn = done_at(p.dones, p.done_kind, static_cast<int64_t>(row) * p.d_rs + env) ? 1u : 0u;
                    cd[i / 4] |= dn << (8 * (i % 4));
                }
            }
        };
        auto store_cols = [&](int buf) {
#pragma unroll
            for (int i = 0; i < kColRegs; ++i) {
                const int row = i * 32 + lane;
                const uint32_t dn = (cd[i / 4] >> (8 * (i % 4))) & 0xffu;
                const uint32_t word = __ballot_sync(0xffffffffu, dn != 0);
                if (i * 32 < N) {  // rows >= N stay zero: the prefix rows continue as constants past the rollout
                    if (row < N) my_rs[buf * q.npad + row] = cr[i];
                    if (lane == 0) my_dm[buf * dm_words + i] = word;
                }
            }
            __syncwarp();
        };

        if (s < total) {
            load_cols(s / q.ncht);
            store_cols(cur);
        }
        for (int g = s; g < total; g += kRingScanWarps) {
            const int e = g / q.ncht, j = g - e * q.ncht;
            const int g_next = g + kRingScanWarps;
            const int e_next = g_next < total ? g_next / q.ncht : e;
            const bool switch_env = e_next != e;
            if (switch_env) load_cols(e_next);  // in flight while this chunk is computed

            mbar_wait(&pempty[s], (use & 1) ^ 1);
#pragma unroll 1
            for (int sp = 0; sp < TC / 32; ++sp) {
                const int tl = sp * 32 + lane;
                const int t = j * TC + tl;  // rows past the rollout read zero rewards
                const float* r_col = my_rs + cur * q.npad + t;
                float* col = myPT + tl;
                // distance to the first terminal at or after t (only terminals inside the n-step window matter)
                int first_done = kNoDone;
                {
                    const int w0 = t >> 5;  // == j * TC / 32 + sp
                    const uint32_t* dm = my_dm + cur * dm_words + w0;
                    uint32_t m = dm[0] & (0xffffffffu << lane);
                    const int words = (nmax + 31 + 31) / 32;
                    int base = -lane;
                    for (int w = 0; w < words; ++w) {
                        if (first_done == kNoDone && m != 0u) first_done = base + __ffs(m) - 1;
                        base += 32;
                        m = (w0 + w + 1 < dm_words) ? dm[w + 1] : 0u;
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
                my_lim[tl] = min(N - t, first_done) * (TCP * 4);  // bootstrap limit of the row, in prefix-row bytes
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[s]);
            ++use;

            if (switch_env) {
                cur ^= 1;
                store_cols(cur);
            }
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
            case 1: consume<V, TC, 1>(q, c); break;
            case 2: consume<V, TC, 2>(q, c); break;
            case 3: consume<V, TC, 3>(q, c); break;
            case 4: consume<V, TC, 4>(q, c); break;
            case 5: consume<V, TC, 5>(q, c); break;
            case 6: consume<V, TC, 6>(q, c); break;
            case 7: consume<V, TC, 7>(q, c); break;
            default: consume<V, TC, 8>(q, c); break;
        }
    }
}

static uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

template <int V, int TC>
static int launch_variant(RingParams& q, const CUtensorMap& vmap, int64_t grid, cudaStream_t stream) {
    PPO_CUDA(cudaFuncSetAttribute(tvf_ring_kernel<V, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(q.smem_total)));
    tvf_ring_kernel<V, TC><<<static_cast<unsigned>(grid), kRingThreads, q.smem_total, stream>>>(q, vmap);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// Shared-memory plan for a time-chunk length; false when it does not fit.
static bool plan_smem(RingParams& q, int N, int V, int nmax, int tc, int max_smem) {
    const uint32_t row_bytes = static_cast<uint32_t>(V) * 4;
    q.cpe = (N + kRC - 1) / kRC;
    q.ncht = (N + tc - 1) / tc;
    // reward rows are read as r_col[t + i], t < ncht * tc, i < nmax + 7, zero past N
    q.npad = ((q.ncht * tc + nmax + 8 + 31) / 32 + 1) * 32;
    const uint32_t chunk_bytes = kRC * row_bytes;
    q.slab_bytes = static_cast<uint32_t>(q.cpe) * chunk_bytes;
    q.pt_buf_bytes = static_cast<uint32_t>(nmax + 1) * (tc + 4) * 4u;
    const uint32_t p_bytes = 2u * q.pt_buf_bytes;
    const uint32_t nd_bytes = 2u * tc * 4u;
    const uint32_t rs_bytes = kRingScanWarps * 2u * q.npad * 4u;
    const uint32_t ds_bytes = align_up(kRingScanWarps * 2u * (q.npad / 32) * 4u, 16);
    const uint32_t gp_bytes = align_up((nmax + 8) * 4u, 16);
    const uint32_t other = p_bytes + nd_bytes + rs_bytes + ds_bytes + gp_bytes + 256;
    int menv = 0;
    while (menv < 4) {
        const uint64_t need = static_cast<uint64_t>(menv + 1) * q.slab_bytes + other +
                              (2ull * (menv + 1) * q.cpe + 4) * 8;
        if (need > static_cast<uint64_t>(max_smem)) break;
        ++menv;
    }
    if (menv < 1) return false;
    q.menv = menv;
    q.nslot = menv * q.cpe;
    q.off_p = align_up(static_cast<uint32_t>(menv) * q.slab_bytes, 128);
    q.off_nd = q.off_p + p_bytes;
    q.off_rs = q.off_nd + nd_bytes;
    q.off_ds = q.off_rs + rs_bytes;
    q.off_gpow = align_up(q.off_ds + ds_bytes, 16);
    q.off_bar = align_up(q.off_gpow + gp_bytes, 16);
    q.smem_total = q.off_bar + (2 * q.nslot + 4) * 8;
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
    int tc = 64;  // longer chunks halve the producer/consumer handshakes; fall back when the tables do not fit
    if (!plan_smem(q, N, V, nmax, tc, max_smem)) {
        tc = 32;
        if (!plan_smem(q, N, V, nmax, tc, max_smem))
            return unsupported("one env slab of value rows plus the prefix tables do not fit in shared memory");
    }
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

#define PPO_RING_CASE(VV)                                                          \
    if (V == VV) return tc == 64 ? launch_variant<VV, 64>(q, vmap, grid, stream)   \
                                 : launch_variant<VV, 32>(q, vmap, grid, stream);
    PPO_RING_CASE(128)
    PPO_RING_CASE(64)
    PPO_RING_CASE(32)
#undef PPO_RING_CASE
    return unsupported("no kernel variant");
}

}  // namespace ppo
