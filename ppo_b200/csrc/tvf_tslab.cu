// TVF targets, transposed slab kernel (PPO_TVF_TSLAB): the path for several n-step samples per horizon.
//
// With E distinct n-steps per horizon every target needs 2 E gathered value samples, so the pass is bound by
// shared-memory gather bandwidth (one 128-byte wavefront per clock per SM), not by HBM.  This kernel is built
// around that limit: every gather is one conflict-free wavefront and nothing else is gathered.
//
//   * Return form.  With R(t) = sum_i r[t+i] * prod_{j<i} gamma * nt[t+j] (the discounted reward-to-go inside
//     the rollout, zero from row N on),  S_n(t) = R(t) - gamma^n * R(t+n)  whenever no terminal lies in
//     t..t+n-1, and S_n(t) = R(t) otherwise.  Hence
//         out[t,k] = R(t) + 1/C * sum_e cnt_e * gamma^n_e * [n_e <= lim(t)] * lerp_e(W[t+n_e, :]),
//         W[r, j] = V[r, j] - R(r),   lim(t) = min(N - t, steps to the first terminal at or after t),
//     so the reward part rides on the value gathers: no prefix tables, no extra reads per entry.  Entries
//     without a bootstrap (h_k - n <= 0) read an extra slab column that holds -R(r).  R is carried in float64
//     and stored as float32; the error model is in DESIGN.md (a few ulp of max|R|).
//   * Transposed slab.  Shared memory holds W^T[j][r] for ONE env: column j of the value curve along time,
//     pitch P odd.  A consumer warp owns 32 consecutive time rows (lane <-> t) and walks horizons; n_e, the
//     column and the lerp weights are warp-uniform (they come from the kernel parameter bank), so a gather is
//     32 consecutive words W^T[lo][t0 + n - 1 + lane]: one wavefront whatever the sample table looks like.
//   * Flow per SM (one persistent CTA, envs in order): TMA loads 32-row chunks of V[:, env, :] into a staging
//     ring; transform warps subtract R(r) and write the chunk transposed into the slab (odd pitch: conflict
//     free) as soon as the consumers of the PREVIOUS env have finished the same 32-row block (the slab is
//     replaced in place, row by row); consumers evaluate blocks of 32 rows x K horizons; results go through a
//     row-major staging tile (16-byte stores, pitch chosen conflict free) and leave with bulk async copies
//     (cp.async.bulk shared -> global), so neither the output nor the input path uses gather bandwidth.
//   * Rows that run off the rollout bootstrap from the final row with host-folded weights (horizon-major
//     tables so that lanes read consecutive words).
// Warp roles: 0-15 consumers, 16-19 transform, 20 scan (R and lim tables, float64), 21 store, 22 TMA producer.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tma.cuh"
#include "tvf_common.cuh"

namespace ppo {

constexpr int kTsConsumers = 16;
constexpr int kTsXform = 4;
constexpr int kTsThreads = (kTsConsumers + kTsXform + 3) * 32;
constexpr int kTsRows = 32;        // value rows per TMA chunk = time rows per lane group
constexpr int kTsStep = 64;        // time rows per consumer step: a thread owns rows t and t + 32 of the step
constexpr int kTsMaxChunks = 16;   // N <= 512
constexpr int kTsMaxSteps = 8;
constexpr unsigned kTsPollNs = 400;  // poll period of the helper warps (producer, transform, store, scan)

struct TsParams {
    TvfParams p;
    double gamma;
    int nc, ns, pitch, nstage, kq, rt_words, tail_pitch, want_stats, n_ent;
    int nslab;            // env slabs in shared memory: 2 when they fit (the next env is staged while this one is evaluated)
    uint32_t slab_bytes;  // bytes of one slab (value columns + -R column + zero column + padding)
    int dbg;  // development switches (PPO_TVF_DEBUG): 1 no stores, 2 no TMA loads, 4 no transform, 8 no entry loop, 16 no tail term
    int64_t env_base, env_rem;
    uint32_t off_stage, off_ot, off_ent, off_hz, off_rtab, off_lim, off_fr, off_bar, smem_total;
    float* value_out;   // [N + 1, A] value estimate sum_j vw[j] * V[r, a, j] (want_value), row stride vo_rs
    int64_t vo_rs;
    int want_value;     // 0 none, 1 the last value horizon (no arithmetic), 2 weighted with vw
    float vw[kTsMaxK];
    double* partials;  // [grid][8] per-CTA statistics of the longest horizon (want_stats)
    unsigned int* ticket;
    double* stats;
    // fused GAE + TD(lambda) over the value estimate of the same launch (rl/rollout.py:1207-1223, rl/returns.py:7-29,
    // 58-67): the scan warp runs the two float64 recurrences of an env once its value column is complete
    int want_gae, gae_same_lambda;
    float gae_gamma32;       // gamma as the float32 scalar of `gamma * value_next_t` (rl/returns.py:26)
    double gae_gl_adv, gae_gl_ret;  // gamma * lambda (float64 products, rl/returns.py:27)
    float* gae_adv;          // [N, A] contiguous
    float* gae_ret;
    double* gae_partials;    // [grid][PPO_STATS_LEN]
    unsigned int* gae_ticket;
    double* gae_stats;       // PPO_STATS_LEN doubles
    uint32_t off_val;        // three slots of {value column, rewards + alive bits} in shared memory (env e uses slot e % 3)
    int val_pitch;           // words per array
    const TsHorizon* hz;  // [K] device copies of the launch constants (inside the packed table blob)
    const TsEntry* ent;   // [n_ent]
    const uint4* tail4;   // [K][tail_pitch] folded tail records {idx lo | hi << 16, cwlo, cwhi, 0}
};

// development: globaltimer stamps of CTA 0 at a few hand-off points (PPO_TVF_DEBUG bit 64), read by ppo_tvf_debug_times
__device__ unsigned long long g_ts_times[16];
__device__ __forceinline__ void ts_stamp(const TsParams& q, int slot) {
    if ((q.dbg & 64) && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_ts_times[slot] = t;
    }
}

template <int IMM>
__device__ __forceinline__ float ts_ldsi(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ float ts_lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int ts_lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 ts_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void ts_sts(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void ts_sts_s32(uint32_t addr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void ts_sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory form of the launch constants: an entry is 12 bytes {off | n << 18, wlo, whi} and every horizon's list
// starts at a multiple of four entries (48 bytes), so four entries are three 16-byte broadcast reads -- 1.5 wavefronts
// per entry instead of the 2 of a 16-byte entry (a broadcast LDS.128 costs two wavefronts, tools/micro_lsu.cu); a
// horizon record is 8 bytes {beg | cnt << 12 | maxn << 20, rscale}.
constexpr uint32_t kTsOffMask = 0x3ffffu;  // byte offset of W^T[col][n - 1] in the slab (< 256 KB)
constexpr int kTsOffBits = 18;
constexpr int kTsEntBytes = 12;

__device__ __forceinline__ uint2 ts_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t ts_lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// acc += [n <= lim] * (lo * wlo + hi * whi) with the predicate taken on the packed word: w = n << 18 | off and
// lim_key = (lim + 1) << 18, so n <= lim  <=>  w < lim_key (no shift per entry).
__device__ __forceinline__ void ts_fma_ifw(float& acc, float lo, float hi, float wlo, float whi, uint32_t w, uint32_t lim_key) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %5, %6;\n\t"
        "@p fma.rn.f32 %0, %1, %3, %0;\n\t"
        "@p fma.rn.f32 %0, %2, %4, %0;\n\t}"
        : "+f"(acc)
        : "f"(lo), "f"(hi), "f"(wlo), "f"(whi), "r"(w), "r"(lim_key));
}

// G entries of one horizon for the thread's two time rows (t, t + 32); `cent` is 16-byte aligned for G = 4 and G = 2
// and 8-byte aligned for G = 1.  All gathers are issued before the first use.  PITCH4 = slab pitch in bytes when it is
// known at compile time (immediate offsets), 0 = runtime pitch in `p4`.
template <int G, int PITCH4>
__device__ __forceinline__ void ts_eval(uint32_t cent, uint32_t base, uint32_t p4, uint32_t lim0, uint32_t lim1,
                                        float& acc0, float& acc1) {
    uint32_t w[3 * G];
    if (G == 4) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint4 v = ts_lds128(cent + i * 16);
            w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
        }
    } else if (G == 2) {
        const uint4 v = ts_lds128(cent);
        const uint2 u = ts_lds64(cent + 16);
        w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w, w[4] = u.x, w[5] = u.y;
    } else {
        const uint2 u = ts_lds64(cent);
        w[0] = u.x, w[1] = u.y, w[2] = ts_lds_u32(cent + 8);
    }
    float lo0[G], hi0[G], lo1[G], hi1[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const uint32_t a = base + (w[3 * i] & kTsOffMask);
        if (PITCH4) {
            lo0[i] = ts_ldsi<0>(a), hi0[i] = ts_ldsi<PITCH4>(a);
            lo1[i] = ts_ldsi<kTsRows * 4>(a), hi1[i] = ts_ldsi<PITCH4 + kTsRows * 4>(a);
        } else {
            const uint32_t ah = a + p4;
            lo0[i] = ts_ldsi<0>(a), hi0[i] = ts_ldsi<0>(ah);
            lo1[i] = ts_ldsi<kTsRows * 4>(a), hi1[i] = ts_ldsi<kTsRows * 4>(ah);
        }
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
        const float wlo = __uint_as_float(w[3 * i + 1]), whi = __uint_as_float(w[3 * i + 2]);
        ts_fma_ifw(acc0, lo0[i], hi0[i], wlo, whi, w[3 * i], lim0);
        ts_fma_ifw(acc1, lo1[i], hi1[i], wlo, whi, w[3 * i], lim1);
    }
}

// GAE: the optional GAE + TD(lambda) stage of the scan warp is compiled in (a separate instantiation: the default
// kernel is sensitive to code size)
template <int V, int PITCH, bool GAE>
__global__ void __launch_bounds__(kTsThreads, 1)
tvf_tslab_kernel(const __grid_constant__ TsParams q, const __grid_constant__ CUtensorMap vmap,
                 const __grid_constant__ CUtensorMap omap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t CHUNK = kTsRows * V * 4;  // bytes per TMA chunk
    constexpr int PITCH4 = PITCH * 4;
    const TvfParams& p = q.p;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
    const int N = p.steps, K = p.n_req, nmax = p.tab.nmax, NC = q.nc, NS = q.ns;
    const int P = PITCH ? PITCH : q.pitch;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + q.off_bar);
    uint64_t* sfull = bars;                     // [4] TMA chunk landed
    uint64_t* sempty = sfull + 4;               // [4] chunk transposed, slot reusable
    uint64_t* wfull = sempty + 4;                   // [2][NC] chunk c of an env is in its slab
    uint64_t* bdone = wfull + 2 * kTsMaxChunks;     // [2][NS] consumers finished step s (slab rows of its chunks reusable)
    uint64_t* rfull = bdone + 2 * kTsMaxSteps;      // [2] R / lim tables of env parity ready
    uint64_t* rempty = rfull + 2;               // [2]
    uint64_t* ofull = rempty + 2;               // output tile written
    uint64_t* oempty = ofull + 1;               // output tile copied out

    const uint32_t wt_a = smem_u32(smem);
    const uint32_t stage_a = smem_u32(smem + q.off_stage);
    const uint32_t ot_a = smem_u32(smem + q.off_ot);
    const uint32_t ent_a = smem_u32(smem + q.off_ent);
    const uint32_t hz_a = smem_u32(smem + q.off_hz);
    const uint32_t rt_a = smem_u32(smem + q.off_rtab);
    const uint32_t lim_a = smem_u32(smem + q.off_lim);
    const uint32_t fr_a = smem_u32(smem + q.off_fr);
    const uint32_t rt_bytes = static_cast<uint32_t>(q.rt_words) * 4u;
    // fused GAE stage: three slots (env e uses slot e % 3) of {value column [N + 1], rewards [N], alive mask [N]}; row r
    // lives at word r + (r >> gsh), which spreads the rows of a lane (stride rows-per-lane) over the banks
    const uint32_t val_a = smem_u32(smem + q.off_val);
    const int val_pitch = q.val_pitch;  // words per array: N + 1 rows, padded, + the alive bits (one word per 32 rows)
    const int gsh = (NC & (NC - 1)) == 0 && NC > 1 ? 31 - __clz(NC) : 31;
    auto gpad = [&](int r) { return static_cast<uint32_t>(r + (r >> gsh)) * 4u; };
    auto gslot = [&](int e, int which) { return val_a + static_cast<uint32_t>(((e % 3) * 2 + which) * val_pitch) * 4u; };
    const uint32_t gbits = gpad(N + 2);  // byte offset of the terminal bits inside array 1

    const int64_t bidx = blockIdx.x;
    const int nenv = static_cast<int>(q.env_base + (bidx < q.env_rem ? 1 : 0));
    const int64_t env0 = bidx * q.env_base + (bidx < q.env_rem ? bidx : q.env_rem);

    if (threadIdx.x == 0) ts_stamp(q, 0);  // kernel start
    // The reward / terminal columns of the first env are requested before anything else: the first R / lim tables
    // gate the whole pipeline fill, and these strided loads are pure latency.
    float rr[kTsMaxChunks];
    uint32_t dw[kTsMaxChunks];  // raw terminal words (nothing depends on the loads until the scan itself)
    auto load_columns = [&](int e) {
        const int64_t env = env0 + e;
#pragma unroll
        for (int i = 0; i < kTsMaxChunks; ++i) {
            const int row = lane * NC + i;
            rr[i] = 0.f, dw[i] = 1u;  // rows past the rollout: reward 0, "terminal" (R = 0, lim = 0)
            if (i < NC && row < N) {
                rr[i] = __ldg(p.rewards + static_cast<int64_t>(row) * p.r_rs + env);
                const int64_t di = static_cast<int64_t>(row) * p.d_rs + env;
                if (p.done_kind == PPO_DONE_U8) dw[i] = __ldg(static_cast<const unsigned char*>(p.dones) + di);
                else if (p.done_kind == PPO_DONE_F32) dw[i] = __float_as_uint(__ldg(static_cast<const float*>(p.dones) + di));
                else dw[i] = 0u;
            }
        }
    };
    if (warp == kTsConsumers + kTsXform && nenv > 0) load_columns(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&sfull[i], 1), mbar_init(&sempty[i], kTsXform);
        for (int i = 0; i < 2 * kTsMaxChunks; ++i) mbar_init(&wfull[i], kTsXform);
        for (int i = 0; i < 2 * kTsMaxSteps; ++i) mbar_init(&bdone[i], kTsConsumers);
        for (int i = 0; i < 2; ++i) mbar_init(&rfull[i], 1), mbar_init(&rempty[i], kTsConsumers + kTsXform);
        mbar_init(ofull, kTsConsumers), mbar_init(oempty, 1);
        mbar_fence_init();
        tma_prefetch_desc(&vmap);
        tma_prefetch_desc(&omap);
    }
    // Everything starts zero: the zero column behind the -R column, the padding behind the slab, R(N) = 0.
    for (uint32_t i = threadIdx.x * 16u; i < q.off_bar; i += blockDim.x * 16u)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // per-horizon launch constants: table blob -> shared memory in the packed form above (read by the consumers as
    // broadcasts).  Thread k owns horizon k: its list starts at the padded prefix sum of the earlier counts.
    {
        __shared__ uint32_t warp_total[kTsMaxK / 32];
        const int k = threadIdx.x;  // K <= 128 < blockDim.x
        uint4 h = make_uint4(0, 0, 0, 0);
        if (k < K) h = __ldg(reinterpret_cast<const uint4*>(q.hz) + k);  // beg, cnt, rscale, maxn
        const uint32_t padded = (h.y + 3u) & ~3u;
        uint32_t incl = padded;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31 && warp < kTsMaxK / 32) warp_total[warp] = incl;
        __syncthreads();
        if (k < K) {
            uint32_t pbeg = incl - padded;
            for (int j = 0; j < warp; ++j) pbeg += warp_total[j];
            *reinterpret_cast<uint2*>(smem + q.off_hz + k * 8) = make_uint2(pbeg | (h.y << 12) | (h.w << 20), h.z);
            for (uint32_t e = 0; e < h.y; ++e) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(q.ent) + h.x + e);  // off, wlo, whi, n
                uint32_t* dst = reinterpret_cast<uint32_t*>(smem + q.off_ent + (pbeg + e) * kTsEntBytes);
                dst[0] = t.x | (t.w << kTsOffBits), dst[1] = t.y, dst[2] = t.z;
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) ts_stamp(q, 1);  // shared memory initialised

    if (warp == kTsConsumers + kTsXform + 2) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int slot = 0;
            uint32_t par = 0;
            for (int e = 0; e < nenv; ++e) {
                const int env = static_cast<int>(env0 + e);
                for (int c = 0; c < NC; ++c) {
                    mbar_wait_sleep(&sempty[slot], par ^ 1, kTsPollNs);
                    if (q.dbg & 2) {
                        mbar_arrive(&sfull[slot]);
                    } else {
                        // rows past N are zero-filled by the TMA unit and still count in the transaction bytes
                        mbar_arrive_expect_tx(&sfull[slot], CHUNK);
                        tma_load_3d(smem + q.off_stage + static_cast<size_t>(slot) * CHUNK, &vmap, &sfull[slot], 0, env,
                                    c * kTsRows + 1);
                    }
                    if (++slot == q.nstage) slot = 0, par ^= 1;
                }
            }
        }
    } else if (warp == kTsConsumers + kTsXform + 1) {
        // ------------------------------------------------------------------ store warp
        uint32_t it = 0;
        for (int e = 0; e < nenv; ++e) {
            const int64_t env = env0 + e;
            for (int s = 0; s < NS; ++s, ++it) {
                mbar_wait_sleep(ofull, it & 1u, kTsPollNs);
                // one tensor store per (32 rows x 32 horizons) sub-tile; rows past N are clipped by the TMA unit
                const int nsub = 2 * (K / 32);
                if (lane < nsub && !(q.dbg & 1)) {
                    const int half = lane / (K / 32), kg = lane % (K / 32);
                    const int row0 = s * kTsStep + half * kTsRows;
                    if (row0 < N)
                        tma_store_3d(&omap, smem + q.off_ot + static_cast<size_t>(lane) * 4096, kg * 32,
                                     static_cast<int>(env), row0);
                }
                bulk_commit();
                bulk_wait_read();
                __syncwarp();
                if (lane == 0) mbar_arrive(oempty);
            }
        }
        bulk_wait_all();
        if (lane == 0) ts_stamp(q, 9);  // last store complete
    } else if (warp == kTsConsumers + kTsXform) {
        // ------------------------------------------------------------------ scan warp: R(t) and lim(t) per env
        const int RPL = NC;  // rows per lane, lane l owns rows [l * RPL, (l + 1) * RPL)
        const double gamma = q.gamma;
        // statistics of the fused GAE stage (this warp sees every env of the CTA)
        double gs_n = 0, gs_sa = 0, gs_qa = 0, gs_sr = 0, gs_qr = 0, gs_sv = 0, gs_qv = 0, gs_nv = 0, gs_za = 0, gs_zr = 0;
        float gs_mna = INFINITY, gs_mxa = -INFINITY, gs_mnr = INFINITY, gs_mxr = -INFINITY;
        // GAE + TD(lambda) of one env: x_t = delta_t + gamma * lambda * nt_t * x_{t+1} with float64 carries and the
        // reference's operation order (rl/returns.py:26-28); lane l owns rows [l * RPL, (l + 1) * RPL), the carries
        // cross the lanes as a suffix scan of affine maps.  Rewards / terminals are read again (they left the
        // registers two envs ago; the lines are still in L2).
        auto gae_env = [&](int e2) {
            if (q.dbg & 1024) return;
            const uint32_t va = gslot(e2, 0), ra = gslot(e2, 1);
            // terminal bits of this lane's rows, bit i = row lane * RPL + i
            uint32_t dbits;
            if (RPL <= 32 && (RPL & (RPL - 1)) == 0)
                dbits = static_cast<uint32_t>(ts_lds_s32(ra + gbits + static_cast<uint32_t>((lane * RPL) >> 5) * 4u)) >> ((lane * RPL) & 31);
            else
                dbits = static_cast<uint32_t>(ts_lds_s32(ra + gbits + static_cast<uint32_t>(lane) * 4u));
            const int64_t env = env0 + e2;
            auto step = [&](float rw, float vv, float vn, bool dn_, double& xa, double& xr) {
                const double boot = dn_ ? 0.0 : static_cast<double>(__fmul_rn(q.gae_gamma32, vn));
                const double delta = __dsub_rn(__dadd_rn(static_cast<double>(rw), boot), static_cast<double>(vv));
                xa = __dadd_rn(delta, __dmul_rn(dn_ ? 0.0 : q.gae_gl_adv, xa));
                xr = __dadd_rn(delta, __dmul_rn(dn_ ? 0.0 : q.gae_gl_ret, xr));
            };
            const int r0 = lane * RPL, r1 = min(r0 + RPL, N);  // this lane's rows [r0, r1)
            double ba = 0, ca = 1, br = 0, cr = 1;
#pragma unroll 1  // compact code: this runs once per env beside the consumers' hot loop (instruction cache)
            for (int row = r1 - 1; row >= r0; --row) {
                const bool dn_ = (dbits >> (row - r0)) & 1u;
                step(ts_lds(ra + gpad(row)), ts_lds(va + gpad(row)), ts_lds(va + gpad(row + 1)), dn_, ba, br);
                ca *= dn_ ? 0.0 : q.gae_gl_adv;
                cr *= dn_ ? 0.0 : q.gae_gl_ret;
            }
#pragma unroll 1
            for (int d = 1; d < 32; d <<= 1) {
                const double ba2 = shfl_down_f64(ba, d), ca2 = shfl_down_f64(ca, d);
                const double br2 = shfl_down_f64(br, d), cr2 = shfl_down_f64(cr, d);
                if (lane + d < 32) {
                    ba = ba + ca * ba2, ca = ca * ca2;
                    br = br + cr * br2, cr = cr * cr2;
                }
            }
            double xa = shfl_down_f64(ba, 1), xr = shfl_down_f64(br, 1);  // carry into this lane's last row
            if (lane == 31) xa = 0.0, xr = 0.0;
#pragma unroll 1
            for (int row = r1 - 1; row >= r0; --row) {
                const bool dn_ = (dbits >> (row - r0)) & 1u;
                const float vv = ts_lds(va + gpad(row)), vn = ts_lds(va + gpad(row + 1));
                step(ts_lds(ra + gpad(row)), vv, vn, dn_, xa, xr);
                const float a = static_cast<float>(xa);
                const float t = __fadd_rn(static_cast<float>(xr), vv);  // rl/returns.py:66-67
                q.gae_adv[static_cast<int64_t>(row) * p.envs + env] = a;
                q.gae_ret[static_cast<int64_t>(row) * p.envs + env] = t;
                gs_n += 1.0, gs_sa += a, gs_qa += static_cast<double>(a) * a, gs_sr += t, gs_qr += static_cast<double>(t) * t;
                gs_za += a == 0.f ? 1.0 : 0.0, gs_zr += t == 0.f ? 1.0 : 0.0;
                gs_mna = fminf(gs_mna, a), gs_mxa = fmaxf(gs_mxa, a), gs_mnr = fminf(gs_mnr, t), gs_mxr = fmaxf(gs_mxr, t);
                gs_sv += vv, gs_qv += static_cast<double>(vv) * vv, gs_nv += 1.0;
                if (row == N - 1) gs_sv += vn, gs_qv += static_cast<double>(vn) * vn, gs_nv += 1.0;  // the final estimate, row N
            }
        };
        for (int e = 0; e < (GAE ? nenv + 2 : nenv); ++e) {
            const uint32_t par = static_cast<uint32_t>(e & 1);
            if (e >= nenv) {
                // the GAE stage of the last two envs: wait until they have left the pipeline
                mbar_wait_sleep(&rempty[par], ((static_cast<uint32_t>(e) >> 1) & 1u) ^ 1u, 4 * kTsPollNs);
                gae_env(e - 2);
                continue;
            }
            if (e > 0) load_columns(e);
            bool dn[kTsMaxChunks];
#pragma unroll
            for (int i = 0; i < kTsMaxChunks; ++i)
                dn[i] = p.done_kind == PPO_DONE_F32 ? (__uint_as_float(dw[i]) != 0.f) : (dw[i] != 0u);
            // segment maps: R_start = a + bc * R_next;  lim_start = hd ? c : c + lim_next
            double a = 0.0, bc = 1.0;
            int c = 0;
            bool hd = false;
#pragma unroll
            for (int i = kTsMaxChunks - 1; i >= 0; --i) {
                if (i < RPL) {
                    const double g = dn[i] ? 0.0 : gamma;
                    a = static_cast<double>(rr[i]) + g * a;
                    bc = g * bc;
                    c = dn[i] ? 0 : c + 1;
                    hd = hd || dn[i];
                }
            }
            // suffix scan over lanes (lane l composes with lanes l+1.. ; beyond lane 31 both are 0)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double a2 = shfl_down_f64(a, d), b2 = shfl_down_f64(bc, d);
                const int c2 = __shfl_down_sync(0xffffffffu, c, d);
                const int hd2 = __shfl_down_sync(0xffffffffu, static_cast<int>(hd), d);
                if (lane + d < 32) {
                    a = a + bc * a2;
                    bc = bc * b2;
                    if (!hd) c += c2, hd = hd2 != 0;
                }
            }
            double x = shfl_down_f64(a, 1);
            int l = __shfl_down_sync(0xffffffffu, c, 1);
            if (lane == 31) x = 0.0, l = 0;
            mbar_wait_sleep(&rempty[par], ((static_cast<uint32_t>(e) >> 1) & 1u) ^ 1u, 4 * kTsPollNs);
            const uint32_t rt = rt_a + par * rt_bytes, lt = lim_a + par * rt_bytes;
#pragma unroll
            for (int i = kTsMaxChunks - 1; i >= 0; --i) {
                if (i < RPL) {
                    const int row = lane * RPL + i;
                    const double g = dn[i] ? 0.0 : gamma;
                    x = static_cast<double>(rr[i]) + g * x;
                    l = dn[i] ? 0 : l + 1;
                    ts_sts(rt + static_cast<uint32_t>(row) * 4u, static_cast<float>(x));
                    ts_sts_s32(lt + static_cast<uint32_t>(row) * 4u, l);
                    if (GAE && row < N) ts_sts(gslot(e, 1) + gpad(row), rr[i]);  // for the GAE stage, two envs later
                }
            }
            if (GAE) {
                // terminal bits of the env, one word per 32 rows (lane l owns rows [l * RPL, (l + 1) * RPL))
                uint32_t mine = 0;
#pragma unroll
                for (int i = 0; i < kTsMaxChunks; ++i)
                    if (i < RPL && dn[i]) mine |= 1u << ((lane * RPL + i) & 31);
                const int per_word = 32 / RPL > 0 ? 32 / RPL : 1;  // lanes that share a word (RPL divides 32 when it is a power of two)
                uint32_t word = mine;
                if (RPL < 32 && (RPL & (RPL - 1)) == 0)
                    for (int d = 1; d < per_word; d <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, d);
                if (RPL <= 32 && (RPL & (RPL - 1)) == 0) {
                    if (lane % per_word == 0) ts_sts_s32(gslot(e, 1) + gbits + static_cast<uint32_t>((lane * RPL) >> 5) * 4u, static_cast<int>(word));
                } else {
                    // rows per lane that do not divide 32: one word per lane (bit i = row lane * RPL + i)
                    uint32_t w2 = 0;
#pragma unroll
                    for (int i = 0; i < kTsMaxChunks; ++i)
                        if (i < RPL && dn[i]) w2 |= 1u << i;
                    ts_sts_s32(gslot(e, 1) + gbits + static_cast<uint32_t>(lane) * 4u, static_cast<int>(w2));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rfull[par]);
            if (lane == 0 && e == 0) ts_stamp(q, 2);  // first R / lim tables ready
            // env e - 2 has left the pipeline (its tables were released above): its value column is complete
            if (GAE && e >= 2) gae_env(e - 2);
        }
        if (GAE) {
            // per-CTA record, then the last CTA folds all records in index order (deterministic, no float atomics)
            double rec[PPO_STATS_LEN];
#pragma unroll
            for (int j = 0; j < PPO_STATS_LEN; ++j) rec[j] = 0.0;
            rec[PPO_STAT_COUNT] = warp_sum_f64(gs_n), rec[PPO_STAT_ADV_SUM] = warp_sum_f64(gs_sa);
            rec[PPO_STAT_ADV_SUMSQ] = warp_sum_f64(gs_qa), rec[PPO_STAT_RET_SUM] = warp_sum_f64(gs_sr);
            rec[PPO_STAT_RET_SUMSQ] = warp_sum_f64(gs_qr), rec[PPO_STAT_VAL_SUM] = warp_sum_f64(gs_sv);
            rec[PPO_STAT_VAL_SUMSQ] = warp_sum_f64(gs_qv), rec[PPO_STAT_VAL_COUNT] = warp_sum_f64(gs_nv);
            rec[PPO_STAT_ADV_ZEROS] = warp_sum_f64(gs_za), rec[PPO_STAT_RET_ZEROS] = warp_sum_f64(gs_zr);
            rec[PPO_STAT_ADV_MIN] = warp_min_f32(gs_mna), rec[PPO_STAT_ADV_MAX] = warp_max_f32(gs_mxa);
            rec[PPO_STAT_RET_MIN] = warp_min_f32(gs_mnr), rec[PPO_STAT_RET_MAX] = warp_max_f32(gs_mxr);
            int last = 0;
            if (lane == 0) {
                double* out = q.gae_partials + static_cast<int64_t>(blockIdx.x) * PPO_STATS_LEN;
#pragma unroll
                for (int j = 0; j < PPO_STATS_LEN; ++j) out[j] = rec[j];
                __threadfence();
                last = atomicAdd(q.gae_ticket, 1u) == gridDim.x - 1;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                if (lane < PPO_STATS_LEN) {
                    const int j = lane;
                    const bool is_min = j == PPO_STAT_ADV_MIN || j == PPO_STAT_RET_MIN;
                    const bool is_max = j == PPO_STAT_ADV_MAX || j == PPO_STAT_RET_MAX;
                    double acc = is_min ? INFINITY : is_max ? -INFINITY : 0.0;
                    for (unsigned int b = 0; b < gridDim.x; ++b) {
                        const double x = __ldcg(q.gae_partials + static_cast<int64_t>(b) * PPO_STATS_LEN + j);
                        acc = is_min ? fmin(acc, x) : is_max ? fmax(acc, x) : acc + x;
                    }
                    q.gae_stats[j] = j <= PPO_STAT_RET_ZEROS ? acc : 0.0;
                }
                __syncwarp();
                if (lane == 0) *q.gae_ticket = 0u;
            }
        }
    } else if (warp >= kTsConsumers) {
        // ------------------------------------------------------------------ transform warps
        const int xw = warp - kTsConsumers;
        int slot = 0;
        uint32_t spar = 0;
        // value estimate for the GAE stage off the rows that pass through here anyway (rl/tvf.py:303-360: the last
        // horizon, or the rediscounted sum over horizons as one dot product per row)
        float vw[V / 32];
#pragma unroll
        for (int cg = 0; cg < V / 32; ++cg) vw[cg] = q.want_value == 2 ? q.vw[cg * 32 + lane] : 0.f;
        for (int e = 0; e < nenv; ++e) {
            const uint32_t par = static_cast<uint32_t>(e & 1);
            float row0_val = 0.f;
            if (q.want_value && xw == 0) {
                // row 0 is not part of the slab (no target bootstraps from it): straight from global memory
                const float* row0 = p.vs + (env0 + e) * p.vs_es;
                float acc = 0.f;
#pragma unroll
                for (int cg = 0; cg < V / 32; ++cg) {
                    const float v = __ldg(row0 + cg * 32 + lane);
                    acc = q.want_value == 2 ? fmaf(v, vw[cg], acc) : (cg == V / 32 - 1 ? v : 0.f);
                }
                if (q.want_value == 2) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                }
                if (lane == 31) q.value_out[env0 + e] = acc;
                row0_val = acc;
            }
            mbar_wait_sleep(&rfull[par], (static_cast<uint32_t>(e) >> 1) & 1u, kTsPollNs);
            // (the value column slot e % 3 was read by the GAE stage of env e - 3, which ran before these tables were released)
            if (GAE && xw == 0 && lane == 31) ts_sts(gslot(e, 0) + gpad(0), row0_val);
            const uint32_t rt = rt_a + par * rt_bytes;
            // slab of this env; it was last used by env e - nslab (the j-th use of a slab waits for use j - 1)
            const int sl = e % q.nslab;
            const uint32_t use = static_cast<uint32_t>(e / q.nslab);
            const uint32_t wt_e = wt_a + static_cast<uint32_t>(sl) * q.slab_bytes;
            const uint32_t fr_e = fr_a + static_cast<uint32_t>(sl * V) * 4u;
            for (int c = 0; c < NC; ++c) {
                // slab rows of chunk c were last needed by step c / 2 of the previous env in this slab
                if (use > 0) mbar_wait_sleep(&bdone[sl * kTsMaxSteps + (c >> 1)], (use - 1) & 1u, kTsPollNs);
                mbar_wait(&sfull[slot], spar);
                const uint32_t src = stage_a + static_cast<uint32_t>(slot) * CHUNK + static_cast<uint32_t>(lane) * 4u;
                float part[kTsRows / kTsXform];  // value-estimate partial sums of this warp's eight rows
#pragma unroll
                for (int rl = 0; rl < kTsRows / kTsXform; ++rl) part[rl] = 0.f;
                if (!(q.dbg & 4))
#pragma unroll
                for (int rl = 0; rl < kTsRows / kTsXform; ++rl) {
                    const int lr = xw + rl * kTsXform;  // row inside the chunk
                    const int i = c * kTsRows + lr;     // slab index: value row r = i + 1
                    if (i < N) {
                        const float Rr = ts_lds(rt + static_cast<uint32_t>(i + 1) * 4u);
                        float v[V / 32];
#pragma unroll
                        for (int cg = 0; cg < V / 32; ++cg) v[cg] = ts_lds(src + static_cast<uint32_t>(lr * V + cg * 32) * 4u);
#pragma unroll
                        for (int cg = 0; cg < V / 32; ++cg) {
                            ts_sts(wt_e + static_cast<uint32_t>((cg * 32 + lane) * P + i) * 4u, v[cg] - Rr);
                            if (i == N - 1) ts_sts(fr_e + static_cast<uint32_t>(cg * 32 + lane) * 4u, v[cg]);
                        }
                        if (q.want_value == 2) {
#pragma unroll
                            for (int cg = 0; cg < V / 32; ++cg) part[rl] = fmaf(v[cg], vw[cg], part[rl]);
                        } else if (q.want_value == 1) {
                            part[rl] = v[V / 32 - 1];  // lane 31 holds the last value horizon
                        }
                    }
                }
                if (q.want_value == 2) {
                    // eight row sums at once: each exchange halves the rows a lane still carries (4 + 2 + 1 shuffles),
                    // two more fold the remaining four lanes -- nine shuffles per chunk instead of forty
                    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
                    float k4[4], k2[2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float got = __shfl_xor_sync(0xffffffffu, b4 ? part[j] : part[j + 4], 16);
                        k4[j] = (b4 ? part[j + 4] : part[j]) + got;
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float got = __shfl_xor_sync(0xffffffffu, b3 ? k4[j] : k4[j + 2], 8);
                        k2[j] = (b3 ? k4[j + 2] : k4[j]) + got;
                    }
                    float k1 = (b2 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? k2[0] : k2[1], 4);
                    k1 += __shfl_xor_sync(0xffffffffu, k1, 2);
                    k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
                    const int rl = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
                    const int i = c * kTsRows + xw + rl * kTsXform;
                    if ((lane & 3) == 0 && i < N) {
                        q.value_out[static_cast<int64_t>(i + 1) * q.vo_rs + env0 + e] = k1;
                        if (GAE) ts_sts(gslot(e, 0) + gpad(i + 1), k1);
                    }
                } else if (q.want_value == 1 && lane == 31) {
#pragma unroll
                    for (int rl = 0; rl < kTsRows / kTsXform; ++rl) {
                        const int i = c * kTsRows + xw + rl * kTsXform;
                        if (i < N) {
                            q.value_out[static_cast<int64_t>(i + 1) * q.vo_rs + env0 + e] = part[rl];
                            if (GAE) ts_sts(gslot(e, 0) + gpad(i + 1), part[rl]);
                        }
                    }
                }
                if (xw == 0) {  // the -R column (index V); column V + 1 stays zero
                    const int i = c * kTsRows + lane;
                    if (i < N) ts_sts(wt_e + static_cast<uint32_t>(V * P + i) * 4u, -ts_lds(rt + static_cast<uint32_t>(i + 1) * 4u));
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&sempty[slot]);
                    mbar_arrive(&wfull[sl * kTsMaxChunks + c]);
                }
                if (lane == 0 && xw == 0 && e == 0 && c == 0) ts_stamp(q, 3);       // first chunk transposed
                if (lane == 0 && xw == 0 && e == 0 && c == NC - 1) ts_stamp(q, 4);  // first env transposed
                if (++slot == q.nstage) slot = 0, spar ^= 1;
            }
            if (lane == 0) mbar_arrive(&rempty[par]);
        }
    } else {
        // ------------------------------------------------------------------ consumers
        const int cw = warp;
        const uint32_t P4 = static_cast<uint32_t>(P) * 4u;
        // statistics of the longest horizon (want_stats): float64 partial sums of this thread
        double st_n = 0.0, st_s = 0.0, st_q = 0.0, st_z = 0.0;
        float st_mn = INFINITY, st_mx = -INFINITY;
        uint32_t it = 0;
        for (int e = 0; e < nenv; ++e) {
            const uint32_t par = static_cast<uint32_t>(e & 1);
            mbar_wait(&rfull[par], (static_cast<uint32_t>(e) >> 1) & 1u);
            const uint32_t rt = rt_a + par * rt_bytes, lt = lim_a + par * rt_bytes;
            int waited = 0;
            const int sl = e % q.nslab;
            const uint32_t use_par = static_cast<uint32_t>(e / q.nslab) & 1u;
            const uint32_t wt_e = wt_a + static_cast<uint32_t>(sl) * q.slab_bytes;
            const uint32_t fr_e = fr_a + static_cast<uint32_t>(sl * V) * 4u;
            for (int s = 0; s < NS; ++s, ++it) {
                const int t0 = s * kTsStep + lane, t1 = t0 + kTsRows;
                const int last_i = min(N - 1, s * kTsStep + kTsStep - 1 + nmax - 1);  // last slab index a live gather may touch
                const int cneed = last_i / kTsRows;
                while (waited <= cneed) {
                    mbar_wait(&wfull[sl * kTsMaxChunks + waited], use_par);
                    ++waited;
                }
                if (q.dbg & 64) {  // (the switch first: one uniform branch per step when it is off)
                    if (lane == 0 && cw == 0 && e == 0 && s == 0) ts_stamp(q, 5);  // consumers start the first step
                    if (lane == 0 && cw == 0 && e == nenv - 1 && s == NS - 1) ts_stamp(q, 7);  // ... the last step
                }
                const int lim0 = ts_lds_s32(lt + static_cast<uint32_t>(t0) * 4u);
                const int lim1 = ts_lds_s32(lt + static_cast<uint32_t>(t1) * 4u);
                const float R0 = ts_lds(rt + static_cast<uint32_t>(t0) * 4u);
                const float R1 = ts_lds(rt + static_cast<uint32_t>(t1) * 4u);
                // n <= lim as one unsigned compare on the packed entry word (ts_fma_ifw)
                const uint32_t key0 = static_cast<uint32_t>(lim0 + 1) << kTsOffBits, key1 = static_cast<uint32_t>(lim1 + 1) << kTsOffBits;
                const int m0 = N - t0, m1 = N - t1;              // rows left
                const int m_min = N - (s * kTsStep + kTsStep - 1);  // smallest m of the step (may be <= 0)
                const bool end0 = m0 >= 1 && m0 <= nmax && lim0 == m0;  // alive until the rollout ends, window may run off
                const bool end1 = m1 >= 1 && m1 <= nmax && lim1 == m1;
                // warp votes: the tail work of a 32-row half is skipped when none of its rows is alive to the end
                const bool any_end0 = __any_sync(0xffffffffu, end0), any_end1 = __any_sync(0xffffffffu, end1);
                const bool any_end = any_end0 || any_end1;
                // this thread's two columns of the folded tail records (only dereferenced for rows alive to the end)
                const uint4* tail0 = q.tail4 + (end0 ? m0 : 0);
                const uint4* tail1 = q.tail4 + (end1 ? m1 : 0);
                const uint32_t base = wt_e + static_cast<uint32_t>(t0) * 4u;
                // output tile: [half][32-horizon group][32 rows][128 B], 16-byte chunks swizzled with the row (TMA 128B
                // swizzle): the 8 lanes of a quarter-warp store to 8 different bank groups
                const uint32_t orow = ot_a + static_cast<uint32_t>(lane) * 128u;
                const uint32_t ohalf = static_cast<uint32_t>(K / 32) * 4096u;
                bool tile_free = false;
                for (int quad = (cw + s) % kTsConsumers; quad < q.kq; quad += kTsConsumers) {
                    float o0[4], o1[4];
                    uint2 hz[4];  // {beg | cnt << 12 | maxn << 20, rscale} of the quad's four horizons: two 16-byte reads
                    {
                        const uint4 a = ts_lds128(hz_a + static_cast<uint32_t>(quad) * 32u);
                        const uint4 b = ts_lds128(hz_a + static_cast<uint32_t>(quad) * 32u + 16u);
                        hz[0] = make_uint2(a.x, a.y), hz[1] = make_uint2(a.z, a.w);
                        hz[2] = make_uint2(b.x, b.y), hz[3] = make_uint2(b.z, b.w);
                    }
                    // rows whose n-step windows run off the rollout: one folded bootstrap from the final row, host
                    // weights from the horizon-major tables (issued for the whole quad before the entry loops)
                    float tacc0[4] = {0.f, 0.f, 0.f, 0.f}, tacc1[4] = {0.f, 0.f, 0.f, 0.f};
                    if (any_end && !(q.dbg & 16) &&
                        m_min < static_cast<int>(max(max(hz[0].x >> 20, hz[1].x >> 20), max(hz[2].x >> 20, hz[3].x >> 20)))) {
                        uint4 r0[4], r1[4];  // one 16-byte record per (horizon, row): slots, folded weights
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int tb = (quad * 4 + u) * q.tail_pitch;  // one 32 x 32 + 64 multiply-add per load
                            r0[u] = make_uint4(0u, 0u, 0u, 0u), r1[u] = r0[u];
                            if (any_end0 && end0) r0[u] = __ldg(tail0 + tb);
                            if (any_end1 && end1) r1[u] = __ldg(tail1 + tb);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (any_end0 && end0) {
                                const float v0 = ts_lds(fr_e + (r0[u].x & 0xffffu) * 4u);
                                const float v1 = ts_lds(fr_e + (r0[u].x >> 16) * 4u);
                                tacc0[u] = fmaf(v1, __uint_as_float(r0[u].z), v0 * __uint_as_float(r0[u].y));
                            }
                            if (any_end1 && end1) {
                                const float v0 = ts_lds(fr_e + (r1[u].x & 0xffffu) * 4u);
                                const float v1 = ts_lds(fr_e + (r1[u].x >> 16) * 4u);
                                tacc1[u] = fmaf(v1, __uint_as_float(r1[u].z), v0 * __uint_as_float(r1[u].y));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float acc0 = tacc0[u], acc1 = tacc1[u];
                        uint32_t ce = ent_a + (hz[u].x & 0xfffu) * kTsEntBytes;
                        int left = (q.dbg & 8) ? 0 : static_cast<int>((hz[u].x >> 12) & 0xffu);
                        for (; left >= 4; left -= 4, ce += 4 * kTsEntBytes) ts_eval<4, PITCH4>(ce, base, P4, key0, key1, acc0, acc1);
                        if (left >= 2) ts_eval<2, PITCH4>(ce, base, P4, key0, key1, acc0, acc1), ce += 2 * kTsEntBytes, left -= 2;
                        if (left >= 1) ts_eval<1, PITCH4>(ce, base, P4, key0, key1, acc0, acc1);
                        const float rs = __uint_as_float(hz[u].y);
                        o0[u] = fmaf(rs, R0, acc0);  // the entry and tail weights carry the 1 / C of the average
                        o1[u] = fmaf(rs, R1, acc1);
                    }
                    if (!tile_free) {  // the previous step's tile has been copied out (single output tile)
                        mbar_wait(oempty, (it & 1u) ^ 1u);
                        tile_free = true;
                    }
                    const uint32_t oq = orow + static_cast<uint32_t>(quad >> 3) * 4096u +
                                        (static_cast<uint32_t>((quad & 7) ^ (lane & 7)) << 4);
                    ts_sts128(oq, o0[0], o0[1], o0[2], o0[3]);
                    ts_sts128(oq + ohalf, o1[0], o1[1], o1[2], o1[3]);
                    if (q.want_stats && quad == q.kq - 1) {
                        if (t0 < N) {
                            const float x = o0[3];
                            st_n += 1.0, st_s += x, st_q += static_cast<double>(x) * x, st_z += x == 0.f ? 1.0 : 0.0;
                            st_mn = fminf(st_mn, x), st_mx = fmaxf(st_mx, x);
                        }
                        if (t1 < N) {
                            const float x = o1[3];
                            st_n += 1.0, st_s += x, st_q += static_cast<double>(x) * x, st_z += x == 0.f ? 1.0 : 0.0;
                            st_mn = fminf(st_mn, x), st_mx = fmaxf(st_mx, x);
                        }
                    }
                }
                if (!tile_free) mbar_wait(oempty, (it & 1u) ^ 1u);  // warps without a quad still keep the phases in step
                fence_async_smem();  // tile writes -> visible to the bulk-copy (async) proxy
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(ofull);
                    mbar_arrive(&bdone[sl * kTsMaxSteps + s]);
                }
            }
            if (lane == 0) mbar_arrive(&rempty[par]);
            if (lane == 0 && cw == 0 && e == 0) ts_stamp(q, 6);  // first env evaluated
        }
        if (lane == 0 && cw == 0) ts_stamp(q, 8);  // consumers done
        if (q.want_stats) {
            // per-CTA record: the staging ring is idle by now (every chunk has been transposed and consumed)
            asm volatile("bar.sync 1, %0;" ::"n"(kTsConsumers * 32) : "memory");
            double* red = reinterpret_cast<double*>(smem + q.off_stage);
            const double n = warp_sum_f64(st_n), sm = warp_sum_f64(st_s), sq = warp_sum_f64(st_q), sz = warp_sum_f64(st_z);
            const float mn = warp_min_f32(st_mn), mx = warp_max_f32(st_mx);
            if (lane == 0) {
                red[cw * 8 + 0] = n, red[cw * 8 + 1] = sm, red[cw * 8 + 2] = sq;
                red[cw * 8 + 3] = static_cast<double>(mn), red[cw * 8 + 4] = static_cast<double>(mx);
                red[cw * 8 + 5] = sz;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTsConsumers * 32) : "memory");
            if (cw == 0) {
                double an = 0, as = 0, aq = 0, az = 0, amn = INFINITY, amx = -INFINITY;
                if (lane < kTsConsumers) {
                    an = red[lane * 8 + 0], as = red[lane * 8 + 1], aq = red[lane * 8 + 2], az = red[lane * 8 + 5];
                    amn = red[lane * 8 + 3], amx = red[lane * 8 + 4];
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    an += shfl_down_f64(an, d), as += shfl_down_f64(as, d), aq += shfl_down_f64(aq, d);
                    az += shfl_down_f64(az, d);
                    amn = fmin(amn, shfl_down_f64(amn, d)), amx = fmax(amx, shfl_down_f64(amx, d));
                }
                int last = 0;
                if (lane == 0) {
                    double* rec = q.partials + static_cast<int64_t>(blockIdx.x) * 8;
                    rec[0] = an, rec[1] = as, rec[2] = aq, rec[3] = amn, rec[4] = amx, rec[5] = az;
                    __threadfence();
                    last = atomicAdd(q.ticket, 1u) == gridDim.x - 1;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    // the last CTA folds all records in index order: deterministic, no float atomics
                    __threadfence();
                    an = 0, as = 0, aq = 0, az = 0, amn = INFINITY, amx = -INFINITY;
                    for (unsigned int g = lane; g < gridDim.x; g += 32) {
                        const double* rec = q.partials + static_cast<int64_t>(g) * 8;
                        an += __ldcg(rec), as += __ldcg(rec + 1), aq += __ldcg(rec + 2), az += __ldcg(rec + 5);
                        amn = fmin(amn, __ldcg(rec + 3)), amx = fmax(amx, __ldcg(rec + 4));
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        an += shfl_down_f64(an, d), as += shfl_down_f64(as, d), aq += shfl_down_f64(aq, d);
                        az += shfl_down_f64(az, d);
                        amn = fmin(amn, shfl_down_f64(amn, d)), amx = fmax(amx, shfl_down_f64(amx, d));
                    }
                    if (lane == 0) {
                        for (int j = 0; j < PPO_STATS_LEN; ++j) q.stats[j] = 0.0;
                        q.stats[PPO_STAT_COUNT] = an;
                        q.stats[PPO_STAT_ADV_SUM] = as;
                        q.stats[PPO_STAT_ADV_SUMSQ] = aq;
                        q.stats[PPO_STAT_ADV_MIN] = amn;
                        q.stats[PPO_STAT_ADV_MAX] = amx;
                        q.stats[PPO_STAT_ADV_ZEROS] = az;
                        *q.ticket = 0u;
                    }
                }
            }
        }
    }
}

static uint32_t ts_align(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

template <int V, int PITCH, bool GAE>
static int ts_launch_one(const TsParams& q, const CUtensorMap& vmap, const CUtensorMap& omap, int64_t grid,
                         cudaStream_t stream) {
    PPO_CUDA(cudaFuncSetAttribute(tvf_tslab_kernel<V, PITCH, GAE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(q.smem_total)));
    tvf_tslab_kernel<V, PITCH, GAE><<<static_cast<unsigned>(grid), kTsThreads, q.smem_total, stream>>>(q, vmap, omap);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

template <int V, int PITCH>
static int ts_launch_variant(const TsParams& q, const CUtensorMap& vmap, const CUtensorMap& omap, int64_t grid,
                             cudaStream_t stream) {
    return q.want_gae ? ts_launch_one<V, PITCH, true>(q, vmap, omap, grid, stream)
                      : ts_launch_one<V, PITCH, false>(q, vmap, omap, grid, stream);
}

// stats / partials / ticket: optional fused statistics of the longest horizon (workspace layout of tvf.cu).
int launch_tvf_tslab(const TvfParams& p, double* stats, double* partials, unsigned int* ticket, int64_t max_partials,
                     const float* host_value_weights, float* value_out, int64_t value_row_stride, cudaStream_t stream,
                     bool* stats_fused, const TsGae* gae, bool* gae_fused) {
    *stats_fused = false;
    if (gae_fused) *gae_fused = false;
    const int N = p.steps, K = p.n_req, V = p.n_val, nmax = p.tab.nmax, E = p.tab.n_entries;
    auto unsupported = [&](const char* why) {
        set_error("ppo_tvf_targets: transposed slab kernel unavailable: %s", why);
        return PPO_ERR_UNSUPPORTED;
    };
    if (V != 32 && V != 64 && V != 128) return unsupported("value horizons must be 32, 64 or 128");
    if (K > kTsMaxK || K % 32 != 0) return unsupported("required horizons must be a multiple of 32, at most 128");
    if (N > kTsMaxChunks * kTsRows) return unsupported("more than 512 steps");
    if (!p.tab.host_blob || !p.tab.dev_blob) return unsupported("tables carry no host mirror (ppo_tvf_tables_build)");
    if (p.tab.tail_pitch <= nmax) return unsupported("tail records missing (tail_pitch)");
    if (!(p.tab.gamma > 0.0)) return unsupported("tables carry no gamma");
    if ((reinterpret_cast<uintptr_t>(p.vs) & 15) || (p.vs_es * 4) % 16 || (p.vs_rs * 4) % 16 || p.vs_es < V)
        return unsupported("value samples are not 16-byte aligned / strided");
    if (reinterpret_cast<uintptr_t>(p.out) & 15) return unsupported("targets are not 16-byte aligned");
    if (p.envs > INT32_MAX) return unsupported("too many envs");
    if (static_cast<int64_t>(K) * E > kTsMaxEntries) return unsupported("more than 1024 (horizon, n-step) entries");
    if (E > 255 || nmax > 1023) return unsupported("more than 255 n-steps per horizon or an n-step above 1023");

    TsParams q{};
    q.p = p;
    q.gamma = p.tab.gamma;
    q.nc = (N + kTsRows - 1) / kTsRows;
    q.ns = (N + kTsStep - 1) / kTsStep;
    q.pitch = N | 1;
    q.kq = K / 4;
    q.rt_words = q.ns * kTsStep + kTsStep;
    q.tail_pitch = p.tab.tail_pitch;
    {
        const char* d = getenv("PPO_TVF_DEBUG");
        q.dbg = d ? atoi(d) : 0;
    }

    // ---- per-horizon launch constants: resolved once by ppo_tvf_tables_build; the header is checked on the host
    // mirror, the kernel reads the device copy
    if (p.tab.tslab_offset_words <= 0) return unsupported("tables carry no launch constants for this kernel");
    const int32_t* hdr = static_cast<const int32_t*>(p.tab.host_blob) + p.tab.tslab_offset_words;
    if (hdr[0] != kTsConstMagic || hdr[1] < 0)
        return unsupported("more than 128 horizons / 1024 entries, or interpolation slots that are not adjacent");
    if (hdr[2] != q.pitch || hdr[3] != V) return unsupported("launch constants were built for another shape");
    const int total = hdr[1];
    const int32_t* dhdr = static_cast<const int32_t*>(p.tab.dev_blob) + p.tab.tslab_offset_words;
    q.hz = reinterpret_cast<const TsHorizon*>(dhdr + 4);
    q.ent = reinterpret_cast<const TsEntry*>(dhdr + 4 + 4 * K);
    q.tail4 = reinterpret_cast<const uint4*>(dhdr + (4 + 4 * K + 4 * static_cast<int64_t>(E) * K + 3) / 4 * 4);
    q.n_ent = total;

    int dev = 0, max_smem = 0;
    PPO_CUDA(cudaGetDevice(&dev));
    PPO_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    max_smem -= 1024;  // the kernel's static shared memory
    const uint32_t chunk = kTsRows * static_cast<uint32_t>(V) * 4u;
    const uint32_t wt_bytes = ts_align((static_cast<uint32_t>(V + 2) * q.pitch + nmax + 2 * kTsStep) * 4u, 1024);
    const uint32_t ot_bytes = static_cast<uint32_t>(kTsStep) * K * 4u;  // dense swizzled sub-tiles of 4096 bytes
    // packed entries, every horizon's list padded to a multiple of four
    const uint32_t ent_bytes = ts_align(static_cast<uint32_t>(total + 3 * K + 4) * kTsEntBytes, 128);
    const uint32_t hz_bytes = ts_align(static_cast<uint32_t>(K) * 8u, 128);
    const uint32_t rt_bytes = ts_align(2u * q.rt_words * 4u, 128);
    const uint32_t fr_bytes = ts_align(2u * static_cast<uint32_t>(V) * 4u, 128);
    const uint32_t bar_bytes = (8u + 2 * kTsMaxChunks + 2 * kTsMaxSteps + 4 + 2) * 8u;
    const bool with_gae = gae && gae->adv && gae->ret && gae->stats && gae->partials && gae->ticket && value_out &&
                          sm_count() <= gae->max_partials;
    {
        const int ncl = q.nc, sh = (ncl & (ncl - 1)) == 0 && ncl > 1 ? 31 - __builtin_clz(ncl) : 31;
        q.val_pitch = N + 2 + ((N + 2) >> sh) + 34;  // rows (padded index) + up to 32 words of terminal bits
    }
    const uint32_t val_bytes = with_gae ? ts_align(6u * q.val_pitch * 4u, 128) : 0u;
    const uint32_t rest = ot_bytes + ent_bytes + hz_bytes + 2 * rt_bytes + fr_bytes + val_bytes + bar_bytes;
    // two slabs when they fit beside two staging chunks: the next env is transposed while this one is evaluated
    // (with one slab the replacement is in place, which stalls when the n-step window is as long as the rollout)
    q.nslab = (2 * wt_bytes + rest + 2 * chunk <= static_cast<uint32_t>(max_smem)) ? 2 : 1;
    q.slab_bytes = wt_bytes;
    const uint32_t fixed = q.nslab * wt_bytes + rest;
    int nstage = 0;
    while (nstage < 4 && fixed + static_cast<uint32_t>(nstage + 1) * chunk <= static_cast<uint32_t>(max_smem)) ++nstage;
    if (nstage < 2) return unsupported("the transposed env slab does not fit in shared memory");
    q.nstage = nstage;
    q.off_stage = q.nslab * wt_bytes;
    q.off_ot = q.off_stage + static_cast<uint32_t>(nstage) * chunk;
    q.off_ent = q.off_ot + ot_bytes;
    q.off_hz = q.off_ent + ent_bytes;
    q.off_rtab = q.off_hz + hz_bytes;
    q.off_lim = q.off_rtab + rt_bytes;
    q.off_fr = q.off_lim + rt_bytes;
    q.off_val = q.off_fr + fr_bytes;
    q.off_bar = q.off_val + val_bytes;
    q.smem_total = q.off_bar + bar_bytes;
    q.want_gae = with_gae ? 1 : 0;
    if (with_gae) {
        q.gae_gamma32 = static_cast<float>(gae->gamma);
        q.gae_gl_adv = gae->gamma * gae->lambda_adv, q.gae_gl_ret = gae->gamma * gae->lambda_ret;
        q.gae_same_lambda = q.gae_gl_adv == q.gae_gl_ret;
        q.gae_adv = gae->adv, q.gae_ret = gae->ret, q.gae_stats = gae->stats;
        q.gae_partials = gae->partials, q.gae_ticket = gae->ticket;
    }

    CUtensorMap vmap;
    const uint64_t dims[3] = {static_cast<uint64_t>(V), static_cast<uint64_t>(p.envs), static_cast<uint64_t>(N) + 1};
    const uint64_t strides[2] = {static_cast<uint64_t>(p.vs_es) * 4, static_cast<uint64_t>(p.vs_rs) * 4};
    const uint32_t box[3] = {static_cast<uint32_t>(V), 1u, static_cast<uint32_t>(kTsRows)};
    int rc = encode_tensor_map(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.vs, dims, strides, box);
    if (rc != PPO_OK) return rc;

    CUtensorMap omap;  // targets [N, A, K]: 32 rows x 32 horizons per store, 128-byte swizzle
    {
        const uint64_t odims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(p.envs), static_cast<uint64_t>(N)};
        const uint64_t ostrides[2] = {static_cast<uint64_t>(K) * 4, static_cast<uint64_t>(p.envs) * K * 4};
        const uint32_t obox[3] = {32u, 1u, static_cast<uint32_t>(kTsRows)};
        rc = encode_tensor_map(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.out, odims, ostrides, obox, true);
        if (rc != PPO_OK) return rc;
    }

    int64_t grid = sm_count();
    if (grid > p.envs) grid = p.envs;
    q.env_base = p.envs / grid;
    q.env_rem = p.envs % grid;
    q.want_value = value_out ? (host_value_weights ? 2 : 1) : 0;
    q.value_out = value_out, q.vo_rs = value_row_stride;
    if (host_value_weights)
        for (int j = 0; j < V; ++j) q.vw[j] = host_value_weights[j];
    q.want_stats = (stats && partials && ticket && grid <= max_partials) ? 1 : 0;
    q.stats = stats, q.partials = partials, q.ticket = ticket;

    // the production shapes get the slab pitch as a compile-time constant (immediate gather offsets)
    if (V == 128 && q.pitch == 257) rc = ts_launch_variant<128, 257>(q, vmap, omap, grid, stream);
    else if (V == 128 && q.pitch == 129) rc = ts_launch_variant<128, 129>(q, vmap, omap, grid, stream);
    else if (V == 128) rc = ts_launch_variant<128, 0>(q, vmap, omap, grid, stream);
    else if (V == 64) rc = ts_launch_variant<64, 0>(q, vmap, omap, grid, stream);
    else rc = ts_launch_variant<32, 0>(q, vmap, omap, grid, stream);
    if (rc != PPO_OK) return rc;
    *stats_fused = q.want_stats != 0;
    if (gae_fused) *gae_fused = q.want_gae != 0;
    return PPO_OK;
}

}  // namespace ppo

// development aid: the stamps of the last launch with PPO_TVF_DEBUG bit 64 (nanoseconds, CTA 0)
extern "C" int ppo_tvf_debug_times(unsigned long long* out16) {
    PPO_CUDA(cudaDeviceSynchronize());
    PPO_CUDA(cudaMemcpyFromSymbol(out16, ppo::g_ts_times, sizeof(unsigned long long) * 16));
    return PPO_OK;
}
