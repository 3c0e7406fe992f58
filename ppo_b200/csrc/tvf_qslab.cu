// TVF targets, quad slab kernel (PPO_TVF_QSLAB): several n-step samples per horizon at 128 / 256 steps x 128 horizons.
//
//   out[t,k] = R(t) + 1/C * sum_e cnt_e * gamma^n_e * [n_e <= lim(t)] * lerp_e(W[t+n_e, :])  (+ tail bootstraps),
//   W[r, j] = V[r, j] - R(r),  R = discounted reward-to-go inside the rollout,  lim(t) = steps to the first terminal
// (the return form of tvf_tslab.cu; rl/returns_truncated.py:558-693), on the same transposed slab W^T[col][row] of ONE
// env, with the same launch constants (the transposed slab section of the table blob) and the same conflict-free
// 4-byte gathers (lane <-> time row, 32 consecutive words per gather) -- and bit-identical targets.  What differs is
// who owns what:
//   * A consumer warp owns eight consecutive horizons (two quads, one after the other) for HALF of the env's rows -- a
//     thread carries rows half * N / 2 + lane + 32 i, i < N / 64 -- so the constants of an entry are fetched twice per
//     env instead of once per 64-row step, an entry costs 23 instructions for 8 gathers instead of 27.5, and there are
//     no per-step hand-offs between the consumer warps at all.
//   * Units (env, eight horizons, half) are drawn from a shared counter in env order, highest horizons first.  Horizon
//     k reads value columns near k, so the slab is refilled in column groups of 32, highest group first: a group is
//     released when the last unit that reads it has finished (arrival counts from the tables), and the next env's
//     columns stream in (TMA, 128-byte swizzle -> transposing writes, lane <-> row, the value estimate for GAE as a
//     thread-local dot product) behind the units that still work on the low horizons.  Half of an env's work separates
//     the release of a group from its next use.
//   * Results leave as one tensor store per unit (N / 2 rows x 32 bytes: whole sectors) from a small pool of output
//     tiles; a store warp issues the stores in the order the consumers publish the tiles and frees a tile as soon as
//     its store has read it, so no consumer ever waits for another consumer.
// Measured on B200 at 256 x 4096 x 128, eight exponential samples per horizon: 447-449 us against
// 453-455 us of the transposed slab kernel for the targets alone, 0.492 against 0.511 ms per bench pass (statistics and
// value estimate fused, a fresh table per pass; same box), with 94 M shared-memory wavefronts and 339 M instructions
// against 105 M / 352 M, and 363-390 us without the output path; at 128 steps x 1024 envs it loses (93 against 76 us).
// What the output path taught: 16-byte global stores straight from the registers reach L2 as partial-sector writes
// (734 us), 32-byte stores (st.global.v8.f32, one sector per lane) 470 us, tiles + tensor stores 460-475 us (before the
// tail records were fetched ahead).  PPO_TVF_AUTO selects this kernel at 256 steps (PPO_TVF_QSLAB=0 / 1 forces it).
// Warp roles: 0-15 consumers, 16-19 transform, 20 scan (R and lim tables, float64), 21 TMA producer, 22 store warp,
// 23 idle (setmaxnreg works on whole warpgroups).  Registers move from the helper warps to the consumers: 80 per
// thread at launch -> 96 / 48 (the CTA pool is 768 x 80 registers: 16 x 32 x 96 + 8 x 32 x 48).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tma.cuh"
#include "tvf_common.cuh"

namespace ppo {

constexpr int kQsConsumers = 16;
constexpr int kQsXform = 4;
constexpr int kQsThreads = (kQsConsumers + kQsXform + 4) * 32;  // two helper warpgroups: transform x 4 | scan, producer, 2 idle
constexpr int kQsChunkRows = 128;                           // value rows per TMA chunk (32 columns wide)
constexpr uint32_t kQsChunkBytes = kQsChunkRows * 128;
constexpr int kQsMaxStages = 4;
constexpr int kQsGroups = 5;                                // four column groups of 32 value columns + the -R column
constexpr int kQsSpecial = kQsGroups - 1;
constexpr int kQsK = 128;                                   // horizons (= 32 quads: see the phase argument below)
constexpr int kQsQuads = kQsK / 4;
constexpr unsigned kQsPollNs = 200;
constexpr uint32_t kQsOffMask = 0x3ffffu;
constexpr int kQsOffBits = 18;
constexpr int kQsEntBytes = 12;
constexpr int kQsMaxTiles = 8;
constexpr unsigned int kQsRing = 32;  // job ring of the store warp (more slots than tiles)

struct QsParams {
    TvfParams p;
    double gamma;
    int tail_pitch, want_stats, want_value, nstage, n_ent;
    int dbg;  // development switches (PPO_TVF_DEBUG): 1 no stores, 2 no TMA loads, 4 no transform
    int arrivals[kQsGroups];     // units (oct, half of the rows) of an env that read the column group
    unsigned char need[kQsQuads / 2];  // column groups an oct (eight consecutive horizons) reads (bit g)
    int64_t env_base, env_rem;
    uint32_t off_stage, off_tile, off_ent, off_hz, off_rtab, off_lim, off_fr, off_misc, off_stat, off_bar, smem_total;
    int ntile;  // output tiles of (N / 2 rows x 32 bytes) in shared memory, handed out through a bit mask
    float* value_out;  // [N + 1, A] value estimate sum_j vw[j] * V[r, a, j] (want_value), row stride vo_rs
    int64_t vo_rs;
    float vw[kQsK];
    double* partials;  // [grid][8] per-CTA statistics of the longest horizon (want_stats)
    unsigned int* ticket;
    double* stats;
    const TsHorizon* hz;  // device copies of the launch constants (transposed slab section of the table blob)
    const TsEntry* ent;
    const uint4* tail4;   // [K][tail_pitch] folded tail records {idx lo | hi << 16, cwlo, cwhi, 0}
};

template <int IMM>
__device__ __forceinline__ float qs_ldsi(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ float qs_lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int qs_lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t qs_lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 qs_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 qs_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 qs_lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void qs_sts(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void qs_sts_s32(uint32_t addr, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ void qs_sts128(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void qs_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// the three live words of a 16-byte tail record {slots, weight lo, weight hi, 0}
__device__ __forceinline__ void qs_ldg_rec(const uint4* p, uint32_t& x, uint32_t& y, uint32_t& z) {
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "l"(p));
    asm volatile("ld.global.nc.u32 %0, [%1+8];" : "=r"(z) : "l"(p));
}
__device__ __forceinline__ void qs_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void qs_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void qs_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void qs_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// acc += [n <= lim] * (lo * wlo + hi * whi), the predicate taken on the packed entry word: w = n << 18 | off and
// key = (lim + 1) << 18, so n <= lim  <=>  w < key.
__device__ __forceinline__ void qs_fma_if(float& acc, float lo, float hi, float wlo, float whi, uint32_t w, uint32_t key) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %5, %6;\n\t"
        "@p fma.rn.f32 %0, %1, %3, %0;\n\t"
        "@p fma.rn.f32 %0, %2, %4, %0;\n\t}"
        : "+f"(acc)
        : "f"(lo), "f"(hi), "f"(wlo), "f"(whi), "r"(w), "r"(key));
}

template <int I, int RG, int PITCH4>
__device__ __forceinline__ void qs_gather(uint32_t a, float (&lo)[RG], float (&hi)[RG]) {
    if constexpr (I < RG) {
        lo[I] = qs_ldsi<I * 128>(a);
        hi[I] = qs_ldsi<PITCH4 + I * 128>(a);
        qs_gather<I + 1, RG, PITCH4>(a, lo, hi);
    }
}

// One entry for the thread's RG rows: 2 RG gathers, all issued before the first use, then the multiply-adds row by row
// within each slot (the four row chains are independent: no multiply-add waits for the one before it).
template <int RG, int PITCH4>
__device__ __forceinline__ void qs_entry(uint32_t w, float wlo, float whi, uint32_t base, const uint32_t (&key)[RG],
                                         float (&acc)[RG]) {
    const uint32_t a = base + (w & kQsOffMask);
    float lo[RG], hi[RG];
    qs_gather<0, RG, PITCH4>(a, lo, hi);
    if constexpr (RG == 4) {
        asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
            "setp.lt.u32 p0, %14, %15;\n\t"
            "setp.lt.u32 p1, %14, %16;\n\t"
            "setp.lt.u32 p2, %14, %17;\n\t"
            "setp.lt.u32 p3, %14, %18;\n\t"
            "@p0 fma.rn.f32 %0, %4, %12, %0;\n\t"
            "@p1 fma.rn.f32 %1, %5, %12, %1;\n\t"
            "@p2 fma.rn.f32 %2, %6, %12, %2;\n\t"
            "@p3 fma.rn.f32 %3, %7, %12, %3;\n\t"
            "@p0 fma.rn.f32 %0, %8, %13, %0;\n\t"
            "@p1 fma.rn.f32 %1, %9, %13, %1;\n\t"
            "@p2 fma.rn.f32 %2, %10, %13, %2;\n\t"
            "@p3 fma.rn.f32 %3, %11, %13, %3;\n\t}"
            : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
            : "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]), "f"(wlo),
              "f"(whi), "r"(w), "r"(key[0]), "r"(key[1]), "r"(key[2]), "r"(key[3]));
    } else {
#pragma unroll
        for (int i = 0; i < RG; ++i) qs_fma_if(acc[i], lo[i], hi[i], wlo, whi, w, key[i]);
    }
}

// G = 2 or 1 entries of one horizon (the packed 12-byte entries of tvf_tslab.cu; `cent` is 8-byte aligned): a pair is
// three broadcast 8-byte reads, one wavefront each -- 1.5 wavefronts per entry like the 16-byte reads of four entries
// there, at half the registers (measured 461 us against 474 us with groups of four).
template <int G, int RG, int PITCH4>
__device__ __forceinline__ void qs_eval(uint32_t cent, uint32_t base, const uint32_t (&key)[RG], float (&acc)[RG]) {
    uint32_t w[3 * G];
    if (G == 2) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint2 u = qs_lds64(cent + i * 8);
            w[2 * i] = u.x, w[2 * i + 1] = u.y;
        }
    } else {
        const uint2 u = qs_lds64(cent);
        w[0] = u.x, w[1] = u.y, w[2] = qs_lds_u32(cent + 8);
    }
#pragma unroll
    for (int i = 0; i < G; ++i)
        qs_entry<RG, PITCH4>(w[3 * i], __uint_as_float(w[3 * i + 1]), __uint_as_float(w[3 * i + 2]), base, key, acc);
}

// RG = N / 32 rows per consumer thread (4 or 8); slab pitch N + 1.
template <int RG>
__global__ void __launch_bounds__(kQsThreads, 1)
tvf_qslab_kernel(const __grid_constant__ QsParams q, const __grid_constant__ CUtensorMap vmap,
                 const __grid_constant__ CUtensorMap omap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int N = 32 * RG;
    constexpr int NCH = N / kQsChunkRows;  // TMA chunks per column group
    constexpr int P = N + 1;               // slab pitch in words (the table blob's pitch, N | 1)
    constexpr int PITCH4 = P * 4;
    constexpr int V = kQsK, NG = V / 32, KQ = kQsQuads;
    constexpr uint32_t RT_BYTES = (N + 32) * 4;
    const TvfParams& p = q.p;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
    const int nmax = p.tab.nmax;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + q.off_bar);
    uint64_t* sfull = bars;                      // [nstage] TMA chunk landed
    uint64_t* sempty = sfull + kQsMaxStages;     // chunk transposed, slot reusable
    uint64_t* gfull = sempty + kQsMaxStages;     // [5] column group of env (phase) is in the slab
    uint64_t* gfree = gfull + kQsGroups;         // [5] every quad of env (phase) that reads the group has finished
    uint64_t* rfull = gfree + kQsGroups;         // [2] R / lim tables of env parity ready
    uint64_t* rempty = rfull + 2;                // [2]

    const uint32_t wt_a = smem_u32(smem);
    const uint32_t stage_a = smem_u32(smem + q.off_stage);
    const uint32_t ent_a = smem_u32(smem + q.off_ent);
    const uint32_t hz_a = smem_u32(smem + q.off_hz);
    const uint32_t rt_a = smem_u32(smem + q.off_rtab);
    const uint32_t lim_a = smem_u32(smem + q.off_lim);
    const uint32_t fr_a = smem_u32(smem + q.off_fr);
    int* unit_counter = reinterpret_cast<int*>(smem + q.off_misc);
    unsigned int* tile_free = reinterpret_cast<unsigned int*>(smem + q.off_misc + 16);  // bit b: output tile b is free
    unsigned int* ring_tail = reinterpret_cast<unsigned int*>(smem + q.off_misc + 32);  // jobs published so far
    unsigned int* ring = reinterpret_cast<unsigned int*>(smem + q.off_misc + 64);       // [kQsRing] job words of filled tiles
    unsigned int* gcount = reinterpret_cast<unsigned int*>(smem + q.off_misc + 192);    // [5] transform-warp arrivals per column group, ever
    const uint32_t tile_a = smem_u32(smem + q.off_tile);

    const int64_t bidx = blockIdx.x;
    const int nenv = static_cast<int>(q.env_base + (bidx < q.env_rem ? 1 : 0));
    const int64_t env0 = bidx * q.env_base + (bidx < q.env_rem ? bidx : q.env_rem);

    // The reward / terminal columns of the first env are requested before anything else: the first R / lim tables
    // gate the pipeline fill, and these strided loads are pure latency.  Lane l owns rows [l * RG, (l + 1) * RG).
    float rr[RG];
    uint32_t dw[RG];
    auto load_columns = [&](int e) {
        const int64_t env = env0 + e;
#pragma unroll
        for (int i = 0; i < RG; ++i) {
            const int row = lane * RG + i;
            rr[i] = __ldg(p.rewards + static_cast<int64_t>(row) * p.r_rs + env);
            const int64_t di = static_cast<int64_t>(row) * p.d_rs + env;
            if (p.done_kind == PPO_DONE_U8) dw[i] = __ldg(static_cast<const unsigned char*>(p.dones) + di);
            else if (p.done_kind == PPO_DONE_F32) dw[i] = __float_as_uint(__ldg(static_cast<const float*>(p.dones) + di));
            else dw[i] = 0u;
        }
    };
    if (warp == kQsConsumers + kQsXform && nenv > 0) load_columns(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kQsMaxStages; ++i) mbar_init(&sfull[i], 1), mbar_init(&sempty[i], kQsXform);
        for (int g = 0; g < kQsGroups; ++g) {
            mbar_init(&gfull[g], kQsXform);
            mbar_init(&gfree[g], q.arrivals[g] > 0 ? q.arrivals[g] : 1);
        }
        for (int i = 0; i < 2; ++i) mbar_init(&rfull[i], 1), mbar_init(&rempty[i], KQ + kQsXform);
        mbar_fence_init();
        tma_prefetch_desc(&vmap);
        tma_prefetch_desc(&omap);
    }
    // Everything starts zero: the zero column behind the -R column, the padding behind the slab, R(N) = 0, the unit
    // counter.
    for (uint32_t i = threadIdx.x * 16u; i < q.off_bar; i += blockDim.x * 16u)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // per-horizon launch constants: table blob -> shared memory in the packed form of tvf_tslab.cu (12-byte entries
    // {off | n << 18, wlo, whi}, every horizon's list padded to a multiple of four; 8-byte horizon records
    // {beg | cnt << 12 | maxn << 20, rscale}).  Thread k owns horizon k.
    {
        __shared__ uint32_t warp_total[kQsK / 32];
        const int k = threadIdx.x;
        uint4 h = make_uint4(0, 0, 0, 0);
        if (k < kQsK) h = __ldg(reinterpret_cast<const uint4*>(q.hz) + k);  // beg, cnt, rscale, maxn
        const uint32_t padded = (h.y + 3u) & ~3u;
        uint32_t incl = padded;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31 && warp < kQsK / 32) warp_total[warp] = incl;
        __syncthreads();
        if (k < kQsK) {
            uint32_t pbeg = incl - padded;
            for (int j = 0; j < warp; ++j) pbeg += warp_total[j];
            *reinterpret_cast<uint2*>(smem + q.off_hz + k * 8) = make_uint2(pbeg | (h.y << 12) | (h.w << 20), h.z);
            for (uint32_t e = 0; e < h.y; ++e) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(q.ent) + h.x + e);  // off, wlo, whi, n
                uint32_t* dst = reinterpret_cast<uint32_t*>(smem + q.off_ent + (pbeg + e) * kQsEntBytes);
                dst[0] = t.x | (t.w << kQsOffBits), dst[1] = t.y, dst[2] = t.z;
            }
        }
    }
    if (threadIdx.x == 0) *tile_free = (1u << q.ntile) - 1u;
    if (threadIdx.x < kQsConsumers) {
        double* rec = reinterpret_cast<double*>(smem + q.off_stat) + threadIdx.x * 8;
        rec[3] = INFINITY, rec[4] = -INFINITY;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= kQsConsumers + kQsXform) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");  // second helper warpgroup
    if (warp > kQsConsumers + kQsXform + 2) {
        // idle warp of the second helper warpgroup
    } else if (warp == kQsConsumers + kQsXform + 2) {
        // ------------------------------------------------------------------ store warp
        // Takes the filled output tiles in the order the consumers publish them, issues one tensor store per tile
        // and frees a tile as soon as its store has read it (stores complete in order; two may be pending).
        if (lane == 0 && !(q.dbg & 1)) {
            const int total = nenv * KQ;
            int open = -1;  // tile of the youngest store, not yet known to have been read
            for (int n = 0; n < total; ++n) {
                volatile unsigned int* slot = ring + (static_cast<unsigned int>(n) % kQsRing);
                unsigned int job;
                while (((job = *slot) & 0x80000000u) == 0u) {
                    if (open >= 0) {  // nothing to do: give the last tile back now
                        qs_bulk_wait_read();
                        atomicOr(tile_free, 1u << open);
                        open = -1;
                    } else {
                        __nanosleep(100);
                    }
                }
                *slot = 0u;
                const unsigned int tb = (job >> 24) & 0x7fu, oct = (job >> 20) & 0xfu, half = (job >> 19) & 1u, e = job & 0x7ffffu;
                qs_tma_store_3d(&omap, tile_a + tb * ((N / 2) * 32u), static_cast<int>(oct * 8),
                                static_cast<int>(env0 + e), static_cast<int>(half * (N / 2)));
                qs_bulk_commit();
                if (open >= 0) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store before this one has read its tile
                    atomicOr(tile_free, 1u << open);
                }
                open = static_cast<int>(tb);
            }
            qs_bulk_wait_all();
        }
    } else if (warp == kQsConsumers + kQsXform + 1) {
        // ------------------------------------------------------------------ TMA producer: groups highest first
        if (lane == 0) {
            int slot = 0;
            uint32_t par = 0;
            for (int e = 0; e < nenv; ++e) {
                const int env = static_cast<int>(env0 + e);
                for (int g = NG - 1; g >= 0; --g)
                    for (int c = 0; c < NCH; ++c) {
                        mbar_wait_sleep(&sempty[slot], par ^ 1, 2 * kQsPollNs);
                        if (q.dbg & 2) {
                            mbar_arrive(&sfull[slot]);
                        } else {
                            mbar_arrive_expect_tx(&sfull[slot], kQsChunkBytes);
                            tma_load_3d(smem + q.off_stage + static_cast<size_t>(slot) * kQsChunkBytes, &vmap, &sfull[slot],
                                        g * 32, env, c * kQsChunkRows + 1);
                        }
                        if (++slot == q.nstage) slot = 0, par ^= 1;
                    }
            }
        }
    } else if (warp == kQsConsumers + kQsXform) {
        // ------------------------------------------------------------------ scan warp: R(t) and lim(t) per env
        const double gamma = q.gamma;
        for (int e = 0; e < nenv; ++e) {
            const uint32_t par = static_cast<uint32_t>(e & 1);
            if (e > 0) load_columns(e);
            bool dn[RG];
#pragma unroll
            for (int i = 0; i < RG; ++i)
                dn[i] = p.done_kind == PPO_DONE_F32 ? (__uint_as_float(dw[i]) != 0.f) : (dw[i] != 0u);
            // segment maps: R_start = a + bc * R_next;  lim_start = hd ? c : c + lim_next
            double a = 0.0, bc = 1.0;
            int c = 0;
            bool hd = false;
#pragma unroll
            for (int i = RG - 1; i >= 0; --i) {
                const double g = dn[i] ? 0.0 : gamma;
                a = static_cast<double>(rr[i]) + g * a;
                bc = g * bc;
                c = dn[i] ? 0 : c + 1;
                hd = hd || dn[i];
            }
            // suffix scan over lanes (lane l composes with lanes l + 1 ..; beyond lane 31 both are 0)
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
            mbar_wait_sleep(&rempty[par], ((static_cast<uint32_t>(e) >> 1) & 1u) ^ 1u, 4 * kQsPollNs);
            const uint32_t rt = rt_a + par * RT_BYTES, lt = lim_a + par * RT_BYTES;
#pragma unroll
            for (int i = RG - 1; i >= 0; --i) {
                const int row = lane * RG + i;
                const double g = dn[i] ? 0.0 : gamma;
                x = static_cast<double>(rr[i]) + g * x;
                l = dn[i] ? 0 : l + 1;
                qs_sts(rt + static_cast<uint32_t>(row) * 4u, static_cast<float>(x));
                qs_sts_s32(lt + static_cast<uint32_t>(row) * 4u, l);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rfull[par]);
        }
    } else if (warp >= kQsConsumers) {
        // ------------------------------------------------------------------ transform warps
        // lane <-> value row of the 128-row chunk (32 rows per warp): 16-byte reads of the swizzled chunk, 4-byte
        // writes along the slab columns (32 consecutive rows: conflict free), and the value estimate for the GAE
        // stage (rl/tvf.py:303-360) as a thread-local dot product over the columns that pass through anyway
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        const int tw = warp - kQsConsumers;
        const int rloc = 32 * tw + lane;
        const int wv = q.want_value;
        int slot = 0;
        uint32_t spar = 0;
        for (int e = 0; e < nenv; ++e) {
            const uint32_t par = static_cast<uint32_t>(e & 1);
            const int64_t env = env0 + e;
            // row 0 is not part of the slab (no target bootstraps from it): straight from global memory, used after
            // the column groups, so its latency hides behind them
            float row0v[NG];
            if (wv && tw == 0) {
#pragma unroll
                for (int jj = 0; jj < NG; ++jj) row0v[jj] = __ldg(p.vs + env * p.vs_es + jj * 32 + lane);
            }
            mbar_wait_sleep(&rfull[par], (static_cast<uint32_t>(e) >> 1) & 1u, kQsPollNs);
            const uint32_t rt = rt_a + par * RT_BYTES;
            float Rr[NCH], dot[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                Rr[c] = qs_lds(rt + static_cast<uint32_t>(c * kQsChunkRows + rloc + 1) * 4u), dot[c] = 0.f;
            for (int g = NG - 1; g >= 0; --g) {
                // the group's columns were last read by the quads of the previous env
                if (e > 0 && q.arrivals[g] > 0) mbar_wait_sleep(&gfree[g], static_cast<uint32_t>(e - 1) & 1u, kQsPollNs);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    mbar_wait(&sfull[slot], spar);
                    const int i = c * kQsChunkRows + rloc;  // slab index: value row r = i + 1
                    const uint32_t src = stage_a + static_cast<uint32_t>(slot) * kQsChunkBytes + static_cast<uint32_t>(rloc) * 128u;
                    const uint32_t dst = wt_a + static_cast<uint32_t>(g * 32 * P + i) * 4u;
                    if (!(q.dbg & 4)) {
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            // four 16-byte reads in flight, then the sixteen column writes
                            float4 v[4];
#pragma unroll
                            for (int cq = 0; cq < 4; ++cq)
                                v[cq] = qs_lds128f(src + (static_cast<uint32_t>((half * 4 + cq) ^ (rloc & 7)) << 4));
#pragma unroll
                            for (int cq = 0; cq < 4; ++cq) {
                                const uint32_t d4 = dst + static_cast<uint32_t>((half * 4 + cq) * 4) * PITCH4;
                                qs_sts(d4, v[cq].x - Rr[c]), qs_sts(d4 + PITCH4, v[cq].y - Rr[c]);
                                qs_sts(d4 + 2 * PITCH4, v[cq].z - Rr[c]), qs_sts(d4 + 3 * PITCH4, v[cq].w - Rr[c]);
                            }
                            if (wv == 2) {
#pragma unroll
                                for (int cq = 0; cq < 4; ++cq) {
                                    const int col = g * 32 + (half * 4 + cq) * 4;
                                    dot[c] = fmaf(v[cq].x, q.vw[col], dot[c]), dot[c] = fmaf(v[cq].y, q.vw[col + 1], dot[c]);
                                    dot[c] = fmaf(v[cq].z, q.vw[col + 2], dot[c]), dot[c] = fmaf(v[cq].w, q.vw[col + 3], dot[c]);
                                }
                            } else if (wv == 1 && half == 1 && g == NG - 1) {
                                dot[c] = v[3].w;
                            }
                            if (c == NCH - 1 && tw == kQsXform - 1 && lane == 31) {  // the final row, for the tail bootstraps
#pragma unroll
                                for (int cq = 0; cq < 4; ++cq) {
                                    const uint32_t f4 = fr_a + static_cast<uint32_t>(g * 32 + (half * 4 + cq) * 4) * 4u;
                                    qs_sts(f4, v[cq].x), qs_sts(f4 + 4, v[cq].y), qs_sts(f4 + 8, v[cq].z), qs_sts(f4 + 12, v[cq].w);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sempty[slot]);
                    if (++slot == q.nstage) slot = 0, spar ^= 1;
                }
                if (lane == 0) {
                    __threadfence_block();
                    atomicAdd(gcount + g, 1u);
                    mbar_arrive(&gfull[g]);
                }
            }
            // the -R column (index V) last: its readers are the lowest horizons, the last quads of an env; column
            // V + 1 stays zero
            if (e > 0 && q.arrivals[kQsSpecial] > 0) mbar_wait_sleep(&gfree[kQsSpecial], static_cast<uint32_t>(e - 1) & 1u, kQsPollNs);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                qs_sts(wt_a + static_cast<uint32_t>(V * P + c * kQsChunkRows + rloc) * 4u, -Rr[c]);
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(gcount + kQsSpecial, 1u);
                mbar_arrive(&gfull[kQsSpecial]);
            }
            if (wv) {
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    q.value_out[static_cast<int64_t>(c * kQsChunkRows + rloc + 1) * q.vo_rs + env] = dot[c];
                if (tw == 0) {
                    float acc = 0.f;
#pragma unroll
                    for (int jj = 0; jj < NG; ++jj)
                        acc = wv == 2 ? fmaf(row0v[jj], q.vw[jj * 32 + lane], acc) : (jj == NG - 1 && lane == 31 ? row0v[jj] : acc);
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    if (lane == 0) q.value_out[env] = acc;
                }
            }
            if (lane == 0) mbar_arrive(&rempty[par]);
        }
    } else {
        // ------------------------------------------------------------------ consumers
        // Units (env, oct, half) are drawn from a shared counter: env order, highest horizons first, 32 units per env.
        // A unit is eight consecutive horizons (two quads, evaluated one after the other) for one half of the env's
        // rows: a thread carries rows half * N / 2 + lane + 32 i, i < RH, and leaves 32 contiguous bytes per row -- a
        // whole sector (16-byte global stores reach L2 as partial-sector writes: 734 instead of 363 us).
        // Phase argument for the parity waits below: units are drawn in order and the transform warps load in order,
        // so a warp that holds a unit of env e + 1 which reads group g can only exist when every unit of env e that
        // reads g has been drawn and either finished or holds another warp; with 16 warps and 32 units per env at
        // least one unit of env e + 1 reading an earlier-loaded group has FINISHED by then (or g is the first group),
        // so gfull[g] has completed phase e and a wait for phase e + 1 cannot alias.  (128 horizons only.)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        constexpr int RH = RG / 2;
        constexpr uint32_t TILE_BYTES = (N / 2) * 32;
        const int total_units = nenv * KQ;
        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(unit_counter, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= total_units) break;
            const int e = u / KQ, j = u - e * KQ;
            const int oct = KQ / 2 - 1 - (j >> 1), half = j & 1;
            const uint32_t par = static_cast<uint32_t>(e & 1);
            mbar_wait(&rfull[par], (static_cast<uint32_t>(e) >> 1) & 1u);
            const uint32_t need = q.need[oct];
#pragma unroll
            for (int g = kQsGroups - 1; g >= 0; --g)
                if ((need >> g) & 1u) {
                    mbar_wait(&gfull[g], par);
                    // the arrival count is the authority (a parity wait can only tell two neighbouring phases apart; the
                    // argument above says that is enough -- this makes a table that breaks it cost a poll, not the data)
                    while (*reinterpret_cast<volatile unsigned int*>(gcount + g) < static_cast<unsigned int>(kQsXform * (e + 1)))
                        __nanosleep(20);
                }
            const uint32_t rt = rt_a + par * RT_BYTES, lt = lim_a + par * RT_BYTES;
            const int t0 = half * (N / 2) + lane;  // the thread's rows: t0 + 32 i
            const uint32_t base_lane = wt_a + static_cast<uint32_t>(t0) * 4u;
            uint32_t key[RH];
            uint32_t endmask = 0;  // bit i: row t0 + 32 i is alive until the rollout ends and its windows may run off
#pragma unroll
            for (int i = 0; i < RH; ++i) {
                const int t = t0 + 32 * i;
                const int lim = qs_lds_s32(lt + static_cast<uint32_t>(t) * 4u);
                key[i] = static_cast<uint32_t>(lim + 1) << kQsOffBits;
                const int m = N - t;
                if (m <= nmax && lim == m) endmask |= 1u << i;
            }
            const bool any_end = __any_sync(0xffffffffu, endmask != 0u);
            float4 prev[RH];  // the lower quad of the oct, kept until the upper one is ready
            uint32_t rx[RH], ry[RH], rz[RH];  // tail records of the horizon about to be evaluated (rows alive to the end)
#pragma unroll
            for (int i = 0; i < RH; ++i) rx[i] = 0u, ry[i] = 0u, rz[i] = 0u;
            if (any_end) {
                const uint4* trow = q.tail4 + static_cast<int64_t>(oct * 8) * q.tail_pitch;
#pragma unroll
                for (int i = 0; i < RH; ++i)
                    if ((endmask >> i) & 1u) qs_ldg_rec(trow + (N - (t0 + 32 * i)), rx[i], ry[i], rz[i]);
            }

#pragma unroll 1
            for (int qq = 0; qq < 2; ++qq) {
                float o[4][RH];
                float rs[4];
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const int k = oct * 8 + qq * 4 + u4;
                    const uint2 hz = qs_lds64(hz_a + static_cast<uint32_t>(k) * 8u);
                    const int maxn = static_cast<int>(hz.x >> 20);
                    float (&acc)[RH] = o[u4];
#pragma unroll
                    for (int i = 0; i < RH; ++i) acc[i] = 0.f;
                    // rows whose n-step windows run off the rollout: one folded bootstrap from the final row (host
                    // weights, one 16-byte record per (horizon, rows left); records exist for rows left < the horizon's
                    // largest n-step).  The records were requested one horizon ago (463.9 -> 447.2 us: a unit of the
                    // upper half no longer waits for L2 in front of every horizon's gathers).
                    if (any_end) {
                        if (maxn > 1) {
#pragma unroll
                            for (int i = 0; i < RH; ++i) {
                                const int m = N - (t0 + 32 * i);
                                if (((endmask >> i) & 1u) && m < maxn) {
                                    const float v0 = qs_lds(fr_a + (rx[i] & 0xffffu) * 4u);
                                    const float v1 = qs_lds(fr_a + (rx[i] >> 16) * 4u);
                                    acc[i] = fmaf(v1, __uint_as_float(rz[i]), v0 * __uint_as_float(ry[i]));
                                }
                            }
                        }
                        if (qq * 4 + u4 < 7) {  // the next horizon's records: their L2 latency hides behind this one's gathers
                            const uint4* trow = q.tail4 + static_cast<int64_t>(k + 1) * q.tail_pitch;
#pragma unroll
                            for (int i = 0; i < RH; ++i)
                                if ((endmask >> i) & 1u) qs_ldg_rec(trow + (N - (t0 + 32 * i)), rx[i], ry[i], rz[i]);
                        }
                    }
                    uint32_t ce = ent_a + (hz.x & 0xfffu) * kQsEntBytes;
                    int left = static_cast<int>((hz.x >> 12) & 0xffu);
                    for (; left >= 2; left -= 2, ce += 2 * kQsEntBytes) qs_eval<2, RH, PITCH4>(ce, base_lane, key, acc);
                    if (left >= 1) qs_eval<1, RH, PITCH4>(ce, base_lane, key, acc);
                    rs[u4] = __uint_as_float(hz.y);
                }
                // out = rscale * R(t) + acc (the entry and tail weights carry the 1 / C of the sample average)
                if (qq == 0) {
#pragma unroll
                    for (int i = 0; i < RH; ++i) {
                        const float R = qs_lds(rt + static_cast<uint32_t>(t0 + 32 * i) * 4u);
                        prev[i].x = fmaf(rs[0], R, o[0][i]), prev[i].y = fmaf(rs[1], R, o[1][i]);
                        prev[i].z = fmaf(rs[2], R, o[2][i]), prev[i].w = fmaf(rs[3], R, o[3][i]);
                    }
                } else {
                    // Both quads are ready: take a free output tile (N / 2 rows x 32 bytes, TMA 32-byte swizzle:
                    // conflict-free 16-byte stores), fill it and hand it to the store warp.  Nothing blocks between
                    // taking the tile and handing it over, and the store warp frees it as soon as the tensor store
                    // has read it, so a handful of tiles serves sixteen warps.
                    int tb = 0;
                    if (!(q.dbg & 1)) {
                        if (lane == 0) {
                            for (;;) {
                                const unsigned int m = *reinterpret_cast<volatile unsigned int*>(tile_free);
                                if (m != 0u) {
                                    tb = __ffs(static_cast<int>(m)) - 1;
                                    if (atomicAnd(tile_free, ~(1u << tb)) & (1u << tb)) break;
                                } else {
                                    __nanosleep(40);
                                }
                            }
                        }
                        tb = __shfl_sync(0xffffffffu, tb, 0);
                    }
                    const uint32_t tile = tile_a + static_cast<uint32_t>(tb) * TILE_BYTES;
                    float smn = INFINITY, smx = -INFINITY;
                    double ssum = 0.0, ssq = 0.0, szero = 0.0;
#pragma unroll
                    for (int i = 0; i < RH; ++i) {
                        const float R = qs_lds(rt + static_cast<uint32_t>(t0 + 32 * i) * 4u);
                        float4 v;
                        v.x = fmaf(rs[0], R, o[0][i]), v.y = fmaf(rs[1], R, o[1][i]);
                        v.z = fmaf(rs[2], R, o[2][i]), v.w = fmaf(rs[3], R, o[3][i]);
                        if (!(q.dbg & 1)) {
                            const uint32_t row = static_cast<uint32_t>(lane + 32 * i);
                            const uint32_t sw = (row >> 2) & 1u;
                            qs_sts128(tile + row * 32u + (sw << 4), prev[i]);
                            qs_sts128(tile + row * 32u + ((sw ^ 1u) << 4), v);
                        }
                        if (q.want_stats && oct == KQ / 2 - 1) {
                            const float x = v.w;
                            ssum += x, ssq += static_cast<double>(x) * x, szero += x == 0.f ? 1.0 : 0.0;
                            smn = fminf(smn, x), smx = fmaxf(smx, x);
                        }
                    }
                    if (!(q.dbg & 1)) {
                        qs_fence_async();  // tile writes -> visible to the async proxy
                        __syncwarp();
                        if (lane == 0) {
                            // job word: valid | tile | oct | half | local env
                            const unsigned int job = 0x80000000u | (static_cast<unsigned int>(tb) << 24) |
                                                     (static_cast<unsigned int>(oct) << 20) |
                                                     (static_cast<unsigned int>(half) << 19) | static_cast<unsigned int>(e);
                            // (every lane's tile writes are fenced above and ordered before this store by the warp barrier)
                            const unsigned int slot = atomicAdd(ring_tail, 1u) % kQsRing;
                            // (a slot is free again long before the ring comes round: at most ntile jobs are open)
                            *reinterpret_cast<volatile unsigned int*>(ring + slot) = job;
                        }
                    }
                    if (q.want_stats && oct == KQ / 2 - 1) {
                        // statistics of the longest horizon: folded per warp into a shared-memory record (two units per
                        // env carry it; registers stay with the gather loop)
                        ssum = warp_sum_f64(ssum), ssq = warp_sum_f64(ssq), szero = warp_sum_f64(szero);
                        smn = warp_min_f32(smn), smx = warp_max_f32(smx);
                        if (lane == 0) {
                            double* rec = reinterpret_cast<double*>(smem + q.off_stat) + warp * 8;
                            rec[0] += 32.0 * RH, rec[1] += ssum, rec[2] += ssq, rec[5] += szero;
                            rec[3] = fmin(rec[3], static_cast<double>(smn)), rec[4] = fmax(rec[4], static_cast<double>(smx));
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < kQsGroups; ++g)
                    if ((need >> g) & 1u) mbar_arrive(&gfree[g]);
                mbar_arrive(&rempty[par]);
            }
        }
        if (q.want_stats) {
            // per-CTA record from the per-warp records
            asm volatile("bar.sync 1, %0;" ::"n"(kQsConsumers * 32) : "memory");
            const double* red = reinterpret_cast<const double*>(smem + q.off_stat);
            const int cw = warp;
            if (cw == 0) {
                double an = 0, as = 0, aq = 0, az = 0, amn = INFINITY, amx = -INFINITY;
                if (lane < kQsConsumers) {
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

static uint32_t qs_align(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

template <int RG>
static int qs_launch_one(const QsParams& q, const CUtensorMap& vmap, const CUtensorMap& omap, int64_t grid,
                         cudaStream_t stream) {
    // The consumers raise their register count to 96 out of the pool the CTA is launched with (768 threads x the
    // kernel's register count): with fewer than 80 registers per thread at launch the request could never be granted
    // and the kernel would hang, so a build that came out smaller declines (PATH_AUTO then takes the transposed slab kernel).
    cudaFuncAttributes attr;
    PPO_CUDA(cudaFuncGetAttributes(&attr, tvf_qslab_kernel<RG>));
    if (attr.numRegs * kQsThreads < kQsConsumers * 32 * 96 + (kQsThreads - kQsConsumers * 32) * 48) {
        set_error("ppo_tvf_targets: quad slab kernel unavailable: built with %d registers per thread, 80 required", attr.numRegs);
        return PPO_ERR_UNSUPPORTED;
    }
    PPO_CUDA(cudaFuncSetAttribute(tvf_qslab_kernel<RG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(q.smem_total)));
    tvf_qslab_kernel<RG><<<static_cast<unsigned>(grid), kQsThreads, q.smem_total, stream>>>(q, vmap, omap);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// stats / partials / ticket: optional fused statistics of the longest horizon (workspace layout of tvf.cu).
int launch_tvf_qslab(const TvfParams& p, double* stats, double* partials, unsigned int* ticket, int64_t max_partials,
                     const float* host_value_weights, float* value_out, int64_t value_row_stride, cudaStream_t stream,
                     bool* stats_fused) {
    *stats_fused = false;
    const int N = p.steps, K = p.n_req, V = p.n_val, nmax = p.tab.nmax, E = p.tab.n_entries;
    auto unsupported = [&](const char* why) {
        set_error("ppo_tvf_targets: quad slab kernel unavailable: %s", why);
        return PPO_ERR_UNSUPPORTED;
    };
    if (N != 128 && N != 256) return unsupported("128 or 256 steps only");
    if (K != kQsK || V != kQsK) return unsupported("128 required horizons on 128 value horizons only");
    if (!p.tab.host_blob || !p.tab.dev_blob) return unsupported("tables carry no host mirror (ppo_tvf_tables_build)");
    if (p.tab.tail_pitch <= nmax) return unsupported("tail records missing (tail_pitch)");
    if (!(p.tab.gamma > 0.0)) return unsupported("tables carry no gamma");
    if ((reinterpret_cast<uintptr_t>(p.vs) & 15) || (p.vs_es * 4) % 16 || (p.vs_rs * 4) % 16 || p.vs_es < V)
        return unsupported("value samples are not 16-byte aligned / strided");
    if (reinterpret_cast<uintptr_t>(p.out) & 31) return unsupported("targets are not 32-byte aligned");
    if (p.envs > INT32_MAX) return unsupported("too many envs");
    if (static_cast<int64_t>(K) * E > kTsMaxEntries) return unsupported("more than 1024 (horizon, n-step) entries");
    if (E > 255 || nmax > 1023) return unsupported("more than 255 n-steps per horizon or an n-step above 1023");
    if (p.tab.tslab_offset_words <= 0) return unsupported("tables carry no launch constants for this kernel");
    const int32_t* hdr = static_cast<const int32_t*>(p.tab.host_blob) + p.tab.tslab_offset_words;
    if (hdr[0] != kTsConstMagic || hdr[1] < 0)
        return unsupported("more than 128 horizons / 1024 entries, or interpolation slots that are not adjacent");
    const int pitch = N | 1;
    if (hdr[2] != pitch || hdr[3] != V) return unsupported("launch constants were built for another shape");
    const int total = hdr[1];

    QsParams q{};
    q.p = p;
    q.gamma = p.tab.gamma;
    q.tail_pitch = p.tab.tail_pitch;
    {
        const char* d = getenv("PPO_TVF_DEBUG");
        q.dbg = d ? atoi(d) : 0;
    }
    const int32_t* dhdr = static_cast<const int32_t*>(p.tab.dev_blob) + p.tab.tslab_offset_words;
    const int64_t tail_words = (4 + 4 * static_cast<int64_t>(K) + 4 * static_cast<int64_t>(E) * K + 3) / 4 * 4;
    q.hz = reinterpret_cast<const TsHorizon*>(dhdr + 4);
    q.ent = reinterpret_cast<const TsEntry*>(dhdr + 4 + 4 * K);
    q.tail4 = reinterpret_cast<const uint4*>(dhdr + tail_words);
    q.n_ent = total;

    // ---- which column groups every oct reads (host mirror of the constants): the gathers of its entries (both
    // slots, whatever their weights -- a zero weight must not meet another env's data) and the final-row slots of its
    // tail records (rows left m = 1 .. largest n-step - 1; the slot falls as m grows, so the two ends bound the range)
    {
        const TsHorizon* hh = reinterpret_cast<const TsHorizon*>(hdr + 4);
        const TsEntry* he = reinterpret_cast<const TsEntry*>(hdr + 4 + 4 * K);
        const int32_t* ht = hdr + tail_words;
        auto group_of = [&](int col) { return col >= V ? kQsSpecial : col / 32; };
        for (int oc = 0; oc < kQsQuads / 2; ++oc) {
            unsigned mask = 0;
            for (int k = 8 * oc; k < 8 * oc + 8; ++k) {
                const TsHorizon& h = hh[k];
                if (h.beg < 0 || h.cnt < 0 || h.beg + h.cnt > total) return unsupported("corrupt launch constants");
                for (int e = h.beg; e < h.beg + h.cnt; ++e) {
                    const int col = static_cast<int>(he[e].off / 4u) / pitch;
                    if (col < 0 || col > V) return unsupported("corrupt launch constants");
                    mask |= 1u << group_of(col);
                    mask |= 1u << group_of(col + 1);
                }
                if (h.maxn > 1) {
                    const int m_lo = 1, m_hi = h.maxn - 1;
                    const uint32_t a = static_cast<uint32_t>(ht[4 * (static_cast<int64_t>(k) * q.tail_pitch + m_lo)]);
                    const uint32_t b = static_cast<uint32_t>(ht[4 * (static_cast<int64_t>(k) * q.tail_pitch + m_hi)]);
                    int c_hi = static_cast<int>(std::max(a >> 16, a & 0xffffu)), c_lo = static_cast<int>(std::min(b >> 16, b & 0xffffu));
                    if (c_hi >= V) c_hi = V - 1;
                    if (c_lo > c_hi) c_lo = c_hi;
                    for (int g = c_lo / 32; g <= c_hi / 32; ++g) mask |= 1u << g;
                }
            }
            q.need[oc] = static_cast<unsigned char>(mask);
            for (int g = 0; g < kQsGroups; ++g) q.arrivals[g] += 2 * ((mask >> g) & 1u);
        }
    }

    int dev = 0, max_smem = 0;
    PPO_CUDA(cudaGetDevice(&dev));
    PPO_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    max_smem -= 1024;  // the kernel's static shared memory
    // slab: V value columns + the -R column + the zero column, and padding for the gathers of masked rows that run
    // past the last column
    const uint32_t wt_bytes = qs_align((static_cast<uint32_t>(V + 2) * pitch + nmax + 64) * 4u, 1024);
    const uint32_t ent_bytes = qs_align(static_cast<uint32_t>(total + 3 * K + 4) * kQsEntBytes, 128);
    const uint32_t hz_bytes = qs_align(static_cast<uint32_t>(K) * 8u, 128);
    const uint32_t rt_bytes = 2u * (static_cast<uint32_t>(N) + 32u) * 4u;  // two env parities
    const uint32_t fr_bytes = qs_align(static_cast<uint32_t>(V) * 4u, 128);
    const uint32_t misc_bytes = 256, stat_bytes = kQsConsumers * 8 * 8;
    const uint32_t bar_bytes = (2 * kQsMaxStages + 2 * kQsGroups + 4) * 8u;
    const uint32_t tile_bytes = static_cast<uint32_t>(N / 2) * 32u;
    q.ntile = kQsMaxTiles;
    if (const char* t = getenv("PPO_QS_TILES")) q.ntile = std::max(1, std::min(kQsMaxTiles, atoi(t)));  // (development)
    const uint32_t rest = q.ntile * tile_bytes + ent_bytes + hz_bytes + 2 * rt_bytes + fr_bytes + misc_bytes + stat_bytes + bar_bytes;
    int nstage = 0;
    while (nstage < kQsMaxStages && wt_bytes + rest + static_cast<uint32_t>(nstage + 1) * kQsChunkBytes <= static_cast<uint32_t>(max_smem))
        ++nstage;
    if (nstage < 2) return unsupported("the env slab does not fit in shared memory");
    q.nstage = nstage;
    q.off_stage = wt_bytes;
    q.off_tile = q.off_stage + static_cast<uint32_t>(nstage) * kQsChunkBytes;
    q.off_ent = q.off_tile + q.ntile * tile_bytes;
    q.off_hz = q.off_ent + ent_bytes;
    q.off_rtab = q.off_hz + hz_bytes;
    q.off_lim = q.off_rtab + rt_bytes;
    q.off_fr = q.off_lim + rt_bytes;
    q.off_misc = q.off_fr + fr_bytes;
    q.off_stat = q.off_misc + misc_bytes;
    q.off_bar = q.off_stat + stat_bytes;
    q.smem_total = q.off_bar + bar_bytes;

    CUtensorMap vmap;  // value samples [N + 1, A, V]: 128 rows x 32 columns per load, 128-byte swizzle
    const uint64_t dims[3] = {static_cast<uint64_t>(V), static_cast<uint64_t>(p.envs), static_cast<uint64_t>(N) + 1};
    const uint64_t strides[2] = {static_cast<uint64_t>(p.vs_es) * 4, static_cast<uint64_t>(p.vs_rs) * 4};
    const uint32_t box[3] = {32u, 1u, static_cast<uint32_t>(kQsChunkRows)};
    int rc = encode_tensor_map(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.vs, dims, strides, box, true);
    if (rc != PPO_OK) return rc;

    CUtensorMap omap;  // targets [N, A, K]: N / 2 rows x 8 horizons per store, 32-byte swizzle
    {
        const uint64_t odims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(p.envs), static_cast<uint64_t>(N)};
        const uint64_t ostrides[2] = {static_cast<uint64_t>(K) * 4, static_cast<uint64_t>(p.envs) * K * 4};
        const uint32_t obox[3] = {8u, 1u, static_cast<uint32_t>(N / 2)};
        rc = encode_tensor_map_swz(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.out, odims, ostrides, obox, 32);
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

    rc = N == 256 ? qs_launch_one<8>(q, vmap, omap, grid, stream) : qs_launch_one<4>(q, vmap, omap, grid, stream);
    if (rc != PPO_OK) return rc;
    *stats_fused = q.want_stats != 0;
    return PPO_OK;
}

}  // namespace ppo
