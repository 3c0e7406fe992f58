// TVF targets, column slab kernel (PPO_TVF_CSLAB): the path for several n-step samples per horizon at 128 / 256 steps.
//
//   out[t,k] = R(t) + 1/C * sum_e cnt_e * gamma^n_e * [n_e <= lim(t)] * lerp_e(W[t+n_e, :])  (+ tail bootstraps),
//   W[r, j] = V[r, j] - R(r),  R = discounted reward-to-go inside the rollout,  lim(t) = steps to the first terminal
// (the return form of tvf_tslab.cu; rl/returns_truncated.py:558-693).  Every target needs two gathered value samples
// per distinct n-step, so the pass is bound by the shared-memory data pipe (128 B per clock per SM), not by HBM.  This
// kernel is built so that the pipe moves nothing but gather payload, at its full width:
//   * Transposed slab W^T[col][row] of ONE env in shared memory, pitch = N.  A consumer thread owns the aligned time
//     rows 4 i .. 4 i + 3 of every 128-row group (i = lane), so a gather is one 16-byte load per thread: 512 contiguous
//     bytes per warp, four conflict-free wavefronts for 4 x 32 rows (a 4-byte gather pays 1.3 clocks per 128 bytes).
//     A gather at n-step n reads the aligned row group "bin" m = (n - 1) / 4, which is shifted against the thread's own
//     rows by s = (n - 1) % 4: the values belong to rows 4 i - s .. 4 i - s + 3, i.e. the first s of them to the
//     previous lane.  They are accumulated in seven registers per row group (three carries + four own rows) and the
//     carries move down one lane once per horizon (three shuffles per 128 rows).  Lane 0 drops its carries (rows
//     before the rollout); the rows lane 31 would receive from a lane 32 lie past the rollout end, where every entry
//     is masked anyway.
//   * n-steps of one horizon that fall into the same bin (n in 4 m - 2 .. 4 m + 1) share their loads: the host sorts
//     the entries by n and flags the first entry of every (column, bin); 1381 instead of 1676 row-group loads per env
//     with the reference's default sample table (eight exponential samples per horizon).
//   * A consumer warp owns four consecutive horizons ("job") for ALL rows of the env, so every entry constant is
//     fetched once per env; the constants live in the kernel parameter (constant) bank and are fetched one entry
//     ahead with LDC -- no shared-memory wavefront at all.
//   * The slab is refilled in column groups of 32: a column group is released when the last job that reads it has
//     finished (per-group arrival counts from the host tables), so the next env's columns stream in (TMA -> staging
//     ring -> transposing writes by two transform warps) behind the jobs that still evaluate the high horizons.
//   * Results leave as one 2 KB tensor store per (job, 128 rows) from a per-warp tile (4-D tensor map {k, env, t / 4,
//     t % 4}: the tile is written with conflict-free 16-byte stores).
// Warp roles: 0-15 consumers (two teams of eight; team jobs handed out by an atomic counter, in env / horizon order),
// 16-19 transform, 20 scan (R and lim tables, float64), 21 TMA producer, 22-23 idle (setmaxnreg works on whole
// warpgroups).  Registers move from the helper warps to the consumers: 80 per thread at launch -> 96 / 48.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "tma.cuh"
#include "tvf_common.cuh"

namespace ppo {

constexpr int kCsConsumers = 16;
constexpr int kCsXform = 4;
constexpr int kCsThreads = (kCsConsumers + kCsXform + 4) * 32;  // two helper warpgroups: transform x 4 | scan, producer, 2 idle
constexpr int kCsChunkRows = 128;                  // value rows per TMA chunk (32 columns wide, 128-byte swizzle)
constexpr uint32_t kCsChunkBytes = kCsChunkRows * 128;
constexpr int kCsMaxStages = 8;
constexpr unsigned kCsPollNs = 200;
constexpr int kCsMaxChunks = 4;                    // N <= 256

struct CsParams {
    TvfParams p;
    double gamma;
    int ng, njobs, n_ent, tail_pitch, want_stats, want_value, nstage;
    int dbg;  // development switches (PPO_TVF_DEBUG): 1 no stores, 2 no TMA loads, 8 no entry loop, 16 no tail term
    int arrivals[kCsGroups];
    unsigned int* sched;  // {next env, finished CTAs}: envs are drawn from a global counter (CTAs differ in speed)
    double* env_stats;    // [envs][8] statistics of the longest horizon per env (want_stats), folded in env order at the end
    uint32_t off_stage, off_tile, off_rtab, off_lim, off_fr, off_ftab, off_nib, off_gpc, off_misc, off_bar, smem_total;
    int xmax, nib_words;      // table horizons 0..xmax of the final value curve; words of the multiplicity table
    const uint32_t* nib_dev;  // multiplicity table (device copy inside the table blob)
    float* value_out;  // [N + 1, A] value estimate sum_j vw[j] * V[r, a, j] (want_value), row stride vo_rs
    int64_t vo_rs;
    unsigned int* ticket;
    double* stats;
    const uint4* tail4;  // [K][tail_pitch] folded tail records {idx lo | hi << 16, cwlo, cwhi, 0}
    float vw[kCsMaxK];
    int vh[kCsMaxK];  // value horizons
    CsHorizon hz[kCsMaxK];
    uint32_t jobmask[kCsMaxK / 4];
    CsEntry ent[kCsMaxEntries + 1];
};

// development: globaltimer stamps of CTA 0 (PPO_TVF_DEBUG bit 256), read by ppo_tvf_cslab_debug_times
__device__ unsigned int g_cs_sched[2 * 64];  // 64 launches in flight at most; the last CTA of a launch zeroes its pair
__device__ unsigned long long g_cs_times[512 + 2 * 160];  // [512 ..): start / end of every CTA
__device__ __forceinline__ void cs_stamp(int dbg, int slot) {
    if ((dbg & 256) && blockIdx.x == 0 && slot < 512) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_cs_times[slot] = t;
    }
}

__device__ __forceinline__ float cs_lds(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 cs_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
template <int IMM>
__device__ __forceinline__ float4 cs_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ void cs_sts(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cs_sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)) : "memory");
}
__device__ __forceinline__ void cs_sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cs_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void cs_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cs_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void cs_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cs_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// acc += [n <= lim] * (lo * wlo + hi * whi), the predicate taken on the packed entry word: w = n << 18 | low bits and
// key = (lim + 1) << 18, so n <= lim  <=>  w < key.
__device__ __forceinline__ void cs_fma_if(float& acc, float lo, float hi, float wlo, float whi, uint32_t w, uint32_t key) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %5, %6;\n\t"
        "@p fma.rn.f32 %0, %1, %3, %0;\n\t"
        "@p fma.rn.f32 %0, %2, %4, %0;\n\t}"
        : "+f"(acc)
        : "f"(lo), "f"(hi), "f"(wlo), "f"(whi), "r"(w), "r"(key));
}

__device__ __forceinline__ uint32_t cs_lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// One entry on the loaded row groups: value q of a group belongs to row 4 i - S + q, accumulator 3 - S + q
// (accumulators 0..2: rows 4 i - 3 .. 4 i - 1 of the previous lane, 3..6: the thread's own rows).
template <int S, int G>
__device__ __forceinline__ void cs_tap(const float4 (&lo)[G], const float4 (&hi)[G], float wlo, float whi, uint32_t w,
                                       float (&acc)[G][7], const uint32_t (&key)[G][7]) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
        cs_fma_if(acc[g][3 - S + 0], lo[g].x, hi[g].x, wlo, whi, w, key[g][3 - S + 0]);
        cs_fma_if(acc[g][3 - S + 1], lo[g].y, hi[g].y, wlo, whi, w, key[g][3 - S + 1]);
        cs_fma_if(acc[g][3 - S + 2], lo[g].z, hi[g].z, wlo, whi, w, key[g][3 - S + 2]);
        cs_fma_if(acc[g][3 - S + 3], lo[g].w, hi[g].w, wlo, whi, w, key[g][3 - S + 3]);
    }
}

// The entries of one horizon with shift S: constants from the parameter bank one entry ahead (`cur` always holds entry i:
// the stream continues across shifts, horizons and quads, so its latency hides behind the gathers), two 16-byte gathers
// per row group, accumulators fixed at compile time.
template <int S, int G>
__device__ __forceinline__ int cs_run(const CsParams& q, int i, int cnt, CsEntry& cur, uint32_t base, float (&acc)[G][7],
                                      const uint32_t (&key)[G][7]) {
    constexpr int COLB = 512 * G;
    const int end = i + cnt;
    for (; i < end; ++i) {
        const CsEntry nxt = q.ent[i + 1];
        const uint32_t a = base + (cur.w & kCsOffMask);
        float4 lo[G], hi[G];
        lo[0] = cs_lds128<0>(a), hi[0] = cs_lds128<COLB>(a);
        if (G > 1) lo[G - 1] = cs_lds128<512>(a), hi[G - 1] = cs_lds128<COLB + 512>(a);
        cs_tap<S, G>(lo, hi, cur.wlo, cur.whi, cur.w, acc, key);
        cur = nxt;
    }
    return i;
}

template <int G>
__global__ void __launch_bounds__(kCsThreads, 1)
tvf_cslab_kernel(const __grid_constant__ CsParams q, const __grid_constant__ CUtensorMap vmap,
                 const __grid_constant__ CUtensorMap omap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int N = 128 * G;
    constexpr int NCH = N / kCsChunkRows;  // TMA chunks per column group
    constexpr int COLB = N * 4;            // bytes between neighbouring slab columns
    const TvfParams& p = q.p;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
    const int K = p.n_req, V = p.n_val, NG = q.ng, NJ = q.njobs;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + q.off_bar);
    uint64_t* sfull = bars;                 // [nstage] TMA chunk landed
    uint64_t* sempty = sfull + kCsMaxStages;   // chunk transposed, slot reusable
    uint64_t* gfull = sempty + kCsMaxStages;   // [5] column group of the current env is in the slab
    uint64_t* gfree = gfull + kCsGroups;    // [5] every job of the env that reads the column group has finished
    uint64_t* rfull = gfree + kCsGroups;    // [2] R / lim tables of env parity ready
    uint64_t* rempty = rfull + 2;           // [2]
    uint64_t* jobready = rempty + 2;        // [team][2] the coordinator has published the team's next job
    uint64_t* tilefull = jobready + 4;      // [team] the eight warps of the team have written their columns of the job tile
    uint64_t* tgrant = tilefull + 2;        // [2] the job tile belongs to job j (parity j & 1): every earlier job has left it

    const uint32_t slab_a = smem_u32(smem);
    const uint32_t stage_a = smem_u32(smem + q.off_stage);
    const uint32_t tile_a = smem_u32(smem + q.off_tile);
    const uint32_t rt_a = smem_u32(smem + q.off_rtab);
    const uint32_t lim_a = smem_u32(smem + q.off_lim);
    const uint32_t fr_a = smem_u32(smem + q.off_fr);
    const uint32_t ftab_a = smem_u32(smem + q.off_ftab);
    const uint32_t nib_a = smem_u32(smem + q.off_nib);
    const uint32_t gpc_a = smem_u32(smem + q.off_gpc);
    const uint32_t vh_a = gpc_a + 512u;  // value horizons (128 words behind the gamma^m / C table)
    constexpr uint32_t RT_BYTES = (N + 16) * 4, LIM_BYTES = (N + 16) * 2;

    // Envs are drawn one at a time from a global counter by the scan warp (SMs differ in speed by 20 % on this pass);
    // local env e of this CTA is global env envq[e & 7], -1 after the last one.
    volatile int* envq_n = reinterpret_cast<volatile int*>(smem + q.off_misc + 92);
    volatile int* envq = reinterpret_cast<volatile int*>(smem + q.off_misc + 96);
    auto env_of = [&](int e) {
        while (*envq_n <= e) __nanosleep(40);
        return envq[e & 7];
    };

    // The reward / terminal columns of the first env are requested before anything else: the first R / lim tables
    // gate the pipeline fill, and these strided loads are pure latency.  Lane l owns rows [l * RPL, (l + 1) * RPL).
    constexpr int RPL = N / 32;
    float rr[RPL];
    uint32_t dw[RPL];
    auto load_columns = [&](int64_t env) {
#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int row = lane * RPL + i;
            rr[i] = __ldg(p.rewards + static_cast<int64_t>(row) * p.r_rs + env);
            const int64_t di = static_cast<int64_t>(row) * p.d_rs + env;
            if (p.done_kind == PPO_DONE_U8) dw[i] = __ldg(static_cast<const unsigned char*>(p.dones) + di);
            else if (p.done_kind == PPO_DONE_F32) dw[i] = __float_as_uint(__ldg(static_cast<const float*>(p.dones) + di));
            else dw[i] = 0u;
        }
    };
    if ((q.dbg & 256) && threadIdx.x == 0 && blockIdx.x < 160) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_cs_times[512 + blockIdx.x] = t;
    }
    int poll_ns = kCsPollNs;
    if (q.dbg >> 16) poll_ns = q.dbg >> 16;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kCsMaxStages; ++i) mbar_init(&sfull[i], 1), mbar_init(&sempty[i], kCsXform);
        for (int g = 0; g < kCsGroups; ++g) {
            mbar_init(&gfull[g], kCsXform);
            mbar_init(&gfree[g], q.arrivals[g] > 0 ? q.arrivals[g] : 1);
        }
        for (int i = 0; i < 2; ++i) mbar_init(&rfull[i], 1), mbar_init(&rempty[i], NJ + kCsXform);
        for (int i = 0; i < 4; ++i) mbar_init(&jobready[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&tilefull[i], 8), mbar_init(&tgrant[i], 1);
        mbar_fence_init();
        tma_prefetch_desc(&vmap);
        tma_prefetch_desc(&omap);
    }
    // Everything starts zero: the zero column behind the -R column, the padding behind the slab, R(N) = 0, lim = 0
    // past the rollout, the statistics and the job counter.
    for (uint32_t i = threadIdx.x * 16u; i < q.off_bar; i += blockDim.x * 16u)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // the multiplicity table of the tail bootstraps and gamma^m / C for the last 128 rows (m = 128 - index)
    for (int i = threadIdx.x; i < q.nib_words; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem + q.off_nib)[i] = __ldg(q.nib_dev + i);
    if (threadIdx.x < 128) {
        const int m = 128 - static_cast<int>(threadIdx.x);
        reinterpret_cast<float*>(smem + q.off_gpc)[threadIdx.x] = m <= p.tab.nmax ? __ldg(p.tab.gpow + m) * p.tab.inv_c : 0.f;
        reinterpret_cast<int*>(smem + q.off_gpc)[128 + threadIdx.x] = q.vh[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        mbar_arrive(&tgrant[0]);  // job 0 owns the tile from the start
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= kCsConsumers) {
        // helper warpgroups: 56 registers for the transform warps, 40 for scan / producer (the scan warp spills a few
        // values to local memory; it runs an env ahead and is latency tolerant)
        if (warp >= kCsConsumers + kCsXform) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");  // second helper warpgroup
        if (warp > kCsConsumers + kCsXform + 1) {
            // ------------------------------------------------------------------ team coordinators (one per team)
            // Hands the team its jobs (32 consecutive horizons of one env; teams alternate within an env and swap
            // roles every env, so both see every horizon group) once their inputs are in shared memory, and sends the
            // team's finished job tile to global memory -- the consumer warps never wait for each other at a barrier.
            const int T = warp - (kCsConsumers + kCsXform + 2);
            const int NH = NJ >> 3;
            auto team_of = [&](int j) { return NH == 1 ? 0 : ((j % NH) + (j / NH)) & 1; };
            auto next_job = [&](int j) {
                while (team_of(j) != T) ++j;
                return j;
            };
            volatile int* slot = reinterpret_cast<volatile int*>(smem + q.off_misc + 72) + T * 2;
            int jpub = next_job(0), jst = jpub;  // next job to publish / to store
            int npub = 0, nst = 0;
            bool done_pub = false;
            while (!done_pub || nst < npub) {
                bool progress = false;
                if (!done_pub && npub - nst < 2 && *envq_n > jpub / NH) {
                    if (envq[(jpub / NH) & 7] < 0) {
                        if (lane == 0) slot[npub & 1] = -1, mbar_arrive(&jobready[T * 2 + (npub & 1)]);
                        done_pub = true, progress = true;
                    } else {
                        const int ee = jpub / NH, hh = jpub - ee * NH;
                        uint32_t um = 0;
#pragma unroll
                        for (int w = 0; w < 8; ++w) um |= q.jobmask[hh * 8 + w];
                        bool ready = mbar_try_wait(&rfull[ee & 1], (static_cast<uint32_t>(ee) >> 1) & 1u);
#pragma unroll
                        for (int g = 0; g < kCsGroups; ++g)
                            if (um >> g & 1u) ready = ready && mbar_try_wait(&gfull[g], static_cast<uint32_t>(ee & 1));
                        if (ready) {
                            if (lane == 0) {
                                if (ee < 4 && hh < 4) cs_stamp(q.dbg, 64 * ee + 12 + 12 * hh + 0);
                                slot[npub & 1] = jpub;
                                mbar_arrive(&jobready[T * 2 + (npub & 1)]);
                            }
                            ++npub, jpub = next_job(jpub + 1), progress = true;
                        }
                    }
                }
                if (nst < npub) {
                    if (mbar_try_wait(&tilefull[T], static_cast<uint32_t>(nst & 1))) {
                        const int ee = jst / NH, hh = jst - ee * NH;
                        if (lane == 0) {
                            if (!(q.dbg & 1)) {
#pragma unroll
                                for (int g = 0; g < G; ++g)
#pragma unroll
                                    for (int c = 0; c < 4; ++c)
                                        cs_store_4d(&omap, tile_a + static_cast<uint32_t>(g * 4 + c) * 4096u, hh * 32,
                                                    envq[ee & 7], c, 32 * g);
                                cs_bulk_commit();
                                cs_bulk_wait_read();
                            }
                            mbar_arrive(&tgrant[(jst + 1) & 1]);  // the tile goes to the next job
                            if (ee < 4 && hh < 4) cs_stamp(q.dbg, 64 * ee + 12 + 12 * hh + 4);
                        }
                        __syncwarp();
                        ++nst, jst = next_job(jst + 1), progress = true;
                    }
                }
                if (!progress) __nanosleep(100);
            }
            if (lane == 0) cs_bulk_wait_all();
            if (lane == 0) cs_stamp(q.dbg, 500 + T);
        } else if (warp == kCsConsumers + kCsXform + 1) {
            // ------------------------------------------------------------------ TMA producer
            if (lane == 0) {
                int slot = 0;
                uint32_t par = 0;
                for (int e = 0;; ++e) {
                    const int env = env_of(e);
                    if (env < 0) break;
                    for (int g = 0; g < NG; ++g)
                        for (int c = 0; c < NCH; ++c) {
                            mbar_wait_sleep(&sempty[slot], par ^ 1, 2 * poll_ns);
                            if (q.dbg & 2) {
                                mbar_arrive(&sfull[slot]);
                            } else {
                                mbar_arrive_expect_tx(&sfull[slot], kCsChunkBytes);
                                tma_load_3d(smem + q.off_stage + static_cast<size_t>(slot) * kCsChunkBytes, &vmap,
                                            &sfull[slot], g * 32, env, c * kCsChunkRows + 1);
                            }
                            if (++slot == q.nstage) slot = 0, par ^= 1;
                        }
                }
            }
        } else if (warp == kCsConsumers + kCsXform) {
            // ------------------------------------------------------------------ scan warp: R(t) and lim(t) per env
            const double gamma = q.gamma;
            for (int e = 0;; ++e) {
                const uint32_t par = static_cast<uint32_t>(e & 1);
                // the tables of env e - 2 must have been consumed before this env is even drawn: drawing late keeps the
                // envs a CTA holds in its pipeline (and with them the imbalance at the end of the launch) small
                mbar_wait_sleep(&rempty[par], ((static_cast<uint32_t>(e) >> 1) & 1u) ^ 1u, 4 * poll_ns);
                int env = -1;
                if (lane == 0) {
                    const unsigned int v = atomicAdd(q.sched, 1u);
                    env = v < static_cast<unsigned int>(p.envs) ? static_cast<int>(v) : -1;
                    envq[e & 7] = env;
                    __threadfence_block();
                    *envq_n = e + 1;
                }
                env = __shfl_sync(0xffffffffu, env, 0);
                if (env < 0) {
                    // no CTA draws after its last draw: the last CTA to get here resets the pair for the next launch
                    if (lane == 0 && atomicAdd(q.sched + 1, 1u) == gridDim.x - 1) q.sched[0] = 0u, q.sched[1] = 0u;
                    break;
                }
                load_columns(env);
                bool dn[RPL];
#pragma unroll
                for (int i = 0; i < RPL; ++i)
                    dn[i] = p.done_kind == PPO_DONE_F32 ? (__uint_as_float(dw[i]) != 0.f) : (dw[i] != 0u);
                // segment maps: R_start = a + bc * R_next;  lim_start = hd ? c : c + lim_next
                double a = 0.0, bc = 1.0;
                int c = 0;
                bool hd = false;
#pragma unroll
                for (int i = RPL - 1; i >= 0; --i) {
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
                const uint32_t rt = rt_a + par * RT_BYTES, lt = lim_a + par * LIM_BYTES;
#pragma unroll
                for (int i = RPL - 1; i >= 0; --i) {
                    const int row = lane * RPL + i;
                    const double g = dn[i] ? 0.0 : gamma;
                    x = static_cast<double>(rr[i]) + g * x;
                    l = dn[i] ? 0 : l + 1;
                    cs_sts(rt + static_cast<uint32_t>(row) * 4u, static_cast<float>(x));
                    cs_sts_u16(lt + static_cast<uint32_t>(row + 4) * 2u, static_cast<uint32_t>(l));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&rfull[par]);
            }
        } else {
            // ------------------------------------------------------------------ transform warps
            asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
            // lane <-> value row of the 128-row chunk (32 rows per warp): 16-byte reads of the swizzled chunk, 4-byte
            // writes along the slab columns (32 consecutive rows: conflict free), and the value estimate for the GAE
            // stage (rl/tvf.py:303-360) as a thread-local dot product over the columns that pass through anyway
            const int tw = warp - kCsConsumers;
            int slot = 0;
            uint32_t spar = 0;
            // row 0 is not part of the slab (no target bootstraps from it): straight from global memory, fetched one env
            // ahead so that its latency is off the critical path
            float row0v[4] = {0.f, 0.f, 0.f, 0.f};
            auto load_row0 = [&](int64_t env) {
                const float* row0 = p.vs + env * p.vs_es;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    if (jj * 32 + lane < V) row0v[jj] = __ldg(row0 + jj * 32 + lane);
            };
            const int wv = q.want_value;
            const int rloc = 32 * tw + lane;
            for (int e = 0;; ++e) {
                const uint32_t par = static_cast<uint32_t>(e & 1);
                const int env = env_of(e);
                if (env < 0) break;
                if (wv && tw == 0) load_row0(env);  // used after the column groups: the latency hides behind them
                if (tw == 0 && lane == 0 && e < 4) cs_stamp(q.dbg, 64 * e + 0);
                mbar_wait(&rfull[par], (static_cast<uint32_t>(e) >> 1) & 1u);
                if (tw == 0 && lane == 0 && e < 4) cs_stamp(q.dbg, 64 * e + 1);
                const uint32_t rt = rt_a + par * RT_BYTES;
                float Rr[NCH], dot[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    Rr[c] = cs_lds(rt + static_cast<uint32_t>(c * kCsChunkRows + rloc + 1) * 4u), dot[c] = 0.f;
                if (e > 0 && q.arrivals[kCsSpecial] > 0) mbar_wait_sleep(&gfree[kCsSpecial], (e - 1) & 1u, poll_ns);
#pragma unroll
                for (int c = 0; c < NCH; ++c)  // the -R column (index V); column V + 1 stays zero
                    cs_sts(slab_a + static_cast<uint32_t>(V * N + c * kCsChunkRows + rloc) * 4u, -Rr[c]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&gfull[kCsSpecial]);
                for (int g = 0; g < NG; ++g) {
                    if (e > 0 && q.arrivals[g] > 0) mbar_wait_sleep(&gfree[g], (e - 1) & 1u, poll_ns);
                    if (tw == 0 && lane == 0 && e < 4) cs_stamp(q.dbg, 64 * e + 2 + 2 * g);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        mbar_wait(&sfull[slot], spar);
                        const int i = c * kCsChunkRows + rloc;  // slab index: value row r = i + 1
                        const uint32_t src = stage_a + static_cast<uint32_t>(slot) * kCsChunkBytes + static_cast<uint32_t>(rloc) * 128u;
                        const uint32_t dst = slab_a + static_cast<uint32_t>(g * 32 * N + i) * 4u;
                        if (!(q.dbg & 4)) {
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                // four 16-byte reads in flight, then the sixteen column writes
                                float4 v[4];
#pragma unroll
                                for (int cq = 0; cq < 4; ++cq)
                                    v[cq] = cs_lds128<0>(src + (static_cast<uint32_t>((half * 4 + cq) ^ (rloc & 7)) << 4));
#pragma unroll
                                for (int cq = 0; cq < 4; ++cq) {
                                    const uint32_t d4 = dst + static_cast<uint32_t>((half * 4 + cq) * 4) * COLB;
                                    cs_sts(d4, v[cq].x - Rr[c]), cs_sts(d4 + COLB, v[cq].y - Rr[c]);
                                    cs_sts(d4 + 2 * COLB, v[cq].z - Rr[c]), cs_sts(d4 + 3 * COLB, v[cq].w - Rr[c]);
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
                                if (c == NCH - 1 && tw == kCsXform - 1 && lane == 31) {  // the final row, for the tail bootstraps
#pragma unroll
                                    for (int cq = 0; cq < 4; ++cq) {
                                        const uint32_t f4 = fr_a + static_cast<uint32_t>(g * 32 + (half * 4 + cq) * 4) * 4u;
                                        cs_sts(f4, v[cq].x), cs_sts(f4 + 4, v[cq].y), cs_sts(f4 + 8, v[cq].z), cs_sts(f4 + 12, v[cq].w);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sempty[slot]);
                        if (++slot == q.nstage) slot = 0, spar ^= 1;
                    }
                    if (q.xmax > 0 && q.vh[g * 32 > 0 ? g * 32 - 1 : 0] < q.xmax) {
                        // the final value curve on the integer horizons of this column group's slots: x in
                        // (vh[s-1], vh[s]] is a line between the final-row values s - 1 and s (rl/returns_truncated.py:
                        // 142-174); the consumers' tail bootstraps read it instead of gathering per target.
                        // lane <-> slot, the four warps interleave the horizons of a slot
                        asm volatile("bar.sync 4, %0;" ::"n"(kCsXform * 32) : "memory");  // the final row of the group is in `fr`
                        const int sl = g * 32 + lane;
                        int xl = 0, dx = 0;
                        float fa = 0.f, slope = 0.f;
                        if (sl < V) {
                            const int xr = static_cast<int>(cs_lds_u32(vh_a + static_cast<uint32_t>(sl) * 4u));
                            xl = sl > 0 ? static_cast<int>(cs_lds_u32(vh_a + static_cast<uint32_t>(sl - 1) * 4u)) : 0;
                            const float fb = cs_lds(fr_a + static_cast<uint32_t>(sl) * 4u);
                            fa = sl > 0 ? cs_lds(fr_a + static_cast<uint32_t>(sl - 1) * 4u) : fb;
                            slope = __fdividef(fb - fa, static_cast<float>(max(xr - xl, 1)));
                            dx = max(min(xr, q.xmax) - xl, 0);
                        }
                        for (int jx = tw; __any_sync(0xffffffffu, jx < dx); jx += kCsXform)
                            if (jx < dx) cs_sts(ftab_a + static_cast<uint32_t>(xl + 1 + jx) * 4u, fmaf(static_cast<float>(jx + 1), slope, fa));
                        __syncwarp();
                    }
                    if (lane == 0) mbar_arrive(&gfull[g]);
                    if (tw == 0 && lane == 0 && e < 4) cs_stamp(q.dbg, 64 * e + 3 + 2 * g);
                }
                if (wv) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        q.value_out[static_cast<int64_t>(c * kCsChunkRows + rloc + 1) * q.vo_rs + env] = dot[c];
                    if (tw == 0) {
                        float acc = 0.f;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = jj * 32 + lane;
                            if (j < V) acc = wv == 2 ? fmaf(row0v[jj], q.vw[j], acc) : (j == V - 1 ? row0v[jj] : acc);
                        }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                        if (lane == 0) q.value_out[env] = acc;
                    }
                }
                if (lane == 0) mbar_arrive(&rempty[par]);
            }
        }
    } else {
        // ---------------------------------------------------------------------- consumers
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        const uint32_t base = slab_a + static_cast<uint32_t>(lane) * 16u;
        // Two teams of eight warps: a team takes 32 consecutive horizons of one env (a "team job", published by the
        // team's coordinator warp), each warp one quad of them for all rows (the quads rotate over the warps from job
        // to job: their costs differ); the team's results leave as 128-byte rows (a 16-byte row costs the TMA unit
        // as much as a 128-byte one) from a job tile that the teams use in job order.
        const int team = warp >> 3, w8 = warp & 7;
        volatile int* slot = reinterpret_cast<volatile int*>(smem + q.off_misc + 72) + team * 2;
        const int NH = NJ >> 3;  // team jobs per env
        for (int it = 0;; ++it) {
            mbar_wait(&jobready[team * 2 + (it & 1)], (static_cast<uint32_t>(it) >> 1) & 1u);
            const int j = slot[it & 1];
            if (j < 0) break;
            const int e = j / NH, hg = j - e * NH, qpos = (w8 + it) & 7, qd = hg * 8 + qpos;
            const bool stamp = w8 == 0 && lane == 0 && e < 4 && hg < 4;
            const uint32_t par = static_cast<uint32_t>(e & 1);
            const uint32_t mask = q.jobmask[qd];
            if (stamp) cs_stamp(q.dbg, 64 * e + 12 + 12 * hg + 1);
            const uint32_t rt = rt_a + par * RT_BYTES, lt = lim_a + par * LIM_BYTES;
            // row limits of rows 4 i - 3 .. 4 i + 3 per group (the first three are the previous lane's last rows; the
            // table starts four entries before row 0, which read as 0: no entry passes)
            uint32_t key[G][7];
            uint32_t endm = 0;  // bit 4 g + j: row 128 g + 4 i + j is alive until the rollout ends
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int r0 = 128 * g + 4 * lane;
                const uint2 a = cs_lds64(lt + static_cast<uint32_t>(r0) * 2u), b = cs_lds64(lt + static_cast<uint32_t>(r0 + 4) * 2u);
                const uint32_t lim[7] = {a.x >> 16, a.y & 0xffffu, a.y >> 16, b.x & 0xffffu, b.x >> 16, b.y & 0xffffu, b.y >> 16};
#pragma unroll
                for (int i = 0; i < 7; ++i) key[g][i] = (lim[i] + 1u) << 18;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (lim[3 + i] == static_cast<uint32_t>(N - r0 - i)) endm |= 1u << (4 * g + i);
            }
            float o[G][4][4];
            uint32_t live = 0;
            CsEntry cur = q.ent[q.hz[qd * 4].w0 & 0x7ffu];  // first entry of the quad (fetched before the row limits)
            if (q.dbg & 32) {
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[g][c][0] = o[g][c][1] = o[g][c][2] = o[g][c][3] = 0.f;
            } else
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = qd * 4 + u;
                const CsHorizon hz = q.hz[k];
                live |= ((hz.w0 >> 23) & 1u) << u;
                float acc[G][7];
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int i = 0; i < 7; ++i) acc[g][i] = 0.f;
                if (!(q.dbg & 8)) {
                    int i = static_cast<int>(hz.w0 & 0x7ffu);  // == ei: the lists of consecutive horizons are contiguous
                    i = cs_run<0, G>(q, i, static_cast<int>(hz.w1 & 0xffu), cur, base, acc, key);
                    i = cs_run<1, G>(q, i, static_cast<int>((hz.w1 >> 8) & 0xffu), cur, base, acc, key);
                    i = cs_run<2, G>(q, i, static_cast<int>((hz.w1 >> 16) & 0xffu), cur, base, acc, key);
                    i = cs_run<3, G>(q, i, static_cast<int>(hz.w1 >> 24), cur, base, acc, key);
                }
                // carries: accumulators 0 .. 2 belong to the last three rows of the previous lane (lane 0 of a group
                // hands over to lane 31 of the group before; lane 0 of the first group drops them)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float r[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) r[g] = __shfl_sync(0xffffffffu, acc[g][c], (lane + 1) & 31);
#pragma unroll
                    for (int g = 0; g < G; ++g) acc[g][4 + c] += lane == 31 ? (g + 1 < G ? r[g + 1 < G ? g + 1 : 0] : 0.f) : r[g];
                }
                // rows whose n-step windows run off the rollout bootstrap from the final row:
                //   tail[t, k] = cnt_k(m) / C * gamma^m * F(h_k - m),  m = N - t,  for rows alive until the rollout ends
                const int maxn = static_cast<int>((hz.w0 >> 11) & 0x3ffu);
                const int mode = static_cast<int>((hz.w0 >> 21) & 3u);
                if (!(q.dbg & 16) && mode == kCsTailGeneric) {
                    // folded records from the horizon-major tables in global memory (log interpolation, long windows)
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        if (N - (128 * g + 127) <= maxn && __any_sync(0xffffffffu, (endm >> (4 * g)) & 0xfu)) {
                            const uint4* trow = q.tail4 + static_cast<int64_t>(k) * q.tail_pitch;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int m = N - (128 * g + 4 * lane + c);
                                if (((endm >> (4 * g + c)) & 1u) && m <= maxn) {
                                    const uint4 rec = __ldg(trow + m);
                                    const float v0 = cs_lds(fr_a + (rec.x & 0xffffu) * 4u);
                                    const float v1 = cs_lds(fr_a + (rec.x >> 16) * 4u);
                                    acc[g][3 + c] += fmaf(v1, __uint_as_float(rec.z), v0 * __uint_as_float(rec.y));
                                }
                            }
                        }
                    }
                } else if (!(q.dbg & 16) && mode != kCsTailNone) {
                    // shared-memory form (n-steps up to 128: only the last row group has such rows): multiplicities from
                    // the 4-bit table, gamma^m / C per row, F from the per-env table of the final value curve (mode 1)
                    // or as a line in m when every h_k - m lies in one interpolation slot (mode 2)
                    const uint32_t em = (endm >> (4 * (G - 1))) & 0xfu;
                    const int m0 = 128 - 4 * lane;  // row c of the group has m = m0 - c
                    if (__any_sync(0xffffffffu, em != 0u && m0 - 3 <= maxn)) {
                        const uint32_t il = hz.w2 + static_cast<uint32_t>(m0 - 4);  // nibble of m = m0 - 3 (row c = 3)
                        const uint32_t wa = cs_lds_u32(nib_a + (il >> 3) * 4u), wb = cs_lds_u32(nib_a + (il >> 3) * 4u + 4u);
                        const uint32_t bits = __funnelshift_r(wa, wb, (il & 7u) * 4u);
                        const float4 gp = cs_lds128<0>(gpc_a + static_cast<uint32_t>(lane) * 16u);
                        float F[4];
                        if (mode == kCsTailLinear) {
                            const uint32_t sc = static_cast<uint32_t>(hz.w3) & 0xffu;
                            const float c0 = static_cast<float>(hz.w3 >> 8);
                            const float fa = cs_lds(fr_a + (sc - 1u) * 4u), fb = cs_lds(fr_a + sc * 4u);
                            const float slope = (fb - fa) * hz.w4;
#pragma unroll
                            for (int c = 0; c < 4; ++c) F[c] = fmaf(c0 - static_cast<float>(m0 - c), slope, fa);
                        } else {
                            const int x0 = hz.w3 + (N - 128) + 4 * lane;  // table horizon h_k - m of row c = 0
#pragma unroll
                            for (int c = 0; c < 4; ++c) F[c] = cs_lds(ftab_a + static_cast<uint32_t>(max(x0 + c, 0)) * 4u);
                        }
                        const float gpv[4] = {gp.x, gp.y, gp.z, gp.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float cntf = static_cast<float>((bits >> (4 * (3 - c))) & 15u);
                            if ((em >> c & 1u) && m0 - c <= maxn) acc[G - 1][3 + c] = fmaf(cntf * gpv[c], F[c], acc[G - 1][3 + c]);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[g][c][u] = acc[g][3 + c];
            }
            // + R(t) (the entry and tail weights carry the 1 / C of the average; h = 0 horizons stay zero)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 R = cs_lds128<0>(rt + static_cast<uint32_t>(128 * g + 4 * lane) * 4u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (live >> u & 1u) o[g][0][u] += R.x, o[g][1][u] += R.y, o[g][2][u] += R.z, o[g][3][u] += R.w;
                }
            }
            // the slab columns and the R / lim tables of this job are no longer needed
            if (stamp) cs_stamp(q.dbg, 64 * e + 12 + 12 * hg + 2);
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < kCsGroups; ++g)
                    if (mask >> g & 1u) mbar_arrive(&gfree[g]);
                mbar_arrive(&rempty[par]);
            }
            if (q.want_stats && qd == NJ - 1) {
                // statistics of the longest horizon: one record per env (folded in env order at the end of the launch,
                // so the sums do not depend on which CTA took which env)
                double sm = 0.0, sq = 0.0;
                float mn = INFINITY, mx = -INFINITY;
                int z = 0;
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float x = o[g][c][3];
                        sm += x, sq += static_cast<double>(x) * x, z += x == 0.f ? 1 : 0;
                        mn = fminf(mn, x), mx = fmaxf(mx, x);
                    }
                sm = warp_sum_f64(sm), sq = warp_sum_f64(sq);
                mn = warp_min_f32(mn), mx = warp_max_f32(mx);
                z = __reduce_add_sync(0xffffffffu, z);
                if (lane == 0) {
                    double* rec = q.env_stats + static_cast<int64_t>(envq[e & 7]) * 8;
                    rec[0] = static_cast<double>(N), rec[1] = sm, rec[2] = sq, rec[3] = static_cast<double>(mn);
                    rec[4] = static_cast<double>(mx), rec[5] = static_cast<double>(z);
                }
            }
            // results: job tile [128-row group][t % 4][t / 4][32 horizons] (128-byte rows in TMA's 128-byte swizzle: the
            // eight lanes of a quarter-warp store to eight different bank groups); the coordinator sends it off once
            // all eight warps have written their quads
            mbar_wait(&tgrant[j & 1], (static_cast<uint32_t>(j) >> 1) & 1u);
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    cs_sts128(tile_a + static_cast<uint32_t>(g * 4 + c) * 4096u + static_cast<uint32_t>(lane) * 128u +
                                  (static_cast<uint32_t>(qpos ^ (lane & 7)) << 4),
                              o[g][c][0], o[g][c][1], o[g][c][2], o[g][c][3]);
            cs_fence_async();  // tile writes -> visible to the bulk-copy (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&tilefull[team]);
            if (stamp) cs_stamp(q.dbg, 64 * e + 12 + 12 * hg + 3);
        }
        if (lane == 0 && (warp & 7) == 0) cs_stamp(q.dbg, 502 + (warp >> 3));
        if ((q.dbg & 256) && threadIdx.x == 0 && blockIdx.x < 160) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            g_cs_times[512 + 160 + blockIdx.x] = t;
        }
        if (q.want_stats) {
            // the last CTA folds the env records: thread t takes envs t, t + 512, ... in order, then a fixed tree
            __threadfence();
            asm volatile("bar.sync 3, %0;" ::"n"(kCsConsumers * 32) : "memory");
            int* flag = reinterpret_cast<int*>(smem + q.off_misc + 64);
            if (threadIdx.x == 0) *flag = atomicAdd(q.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
            asm volatile("bar.sync 3, %0;" ::"n"(kCsConsumers * 32) : "memory");
            if (*flag) {
                __threadfence();
                double an = 0, as = 0, aq = 0, az = 0, amn = INFINITY, amx = -INFINITY;
                for (int64_t g = threadIdx.x; g < p.envs; g += kCsConsumers * 32) {
                    const double* rec = q.env_stats + g * 8;
                    an += __ldcg(rec), as += __ldcg(rec + 1), aq += __ldcg(rec + 2), az += __ldcg(rec + 5);
                    amn = fmin(amn, __ldcg(rec + 3)), amx = fmax(amx, __ldcg(rec + 4));
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    an += shfl_down_f64(an, d), as += shfl_down_f64(as, d), aq += shfl_down_f64(aq, d);
                    az += shfl_down_f64(az, d);
                    amn = fmin(amn, shfl_down_f64(amn, d)), amx = fmax(amx, shfl_down_f64(amx, d));
                }
                double* red = reinterpret_cast<double*>(smem + q.off_tile);  // the job tile is idle by now
                if (lane == 0) {
                    red[warp * 8 + 0] = an, red[warp * 8 + 1] = as, red[warp * 8 + 2] = aq, red[warp * 8 + 3] = amn;
                    red[warp * 8 + 4] = amx, red[warp * 8 + 5] = az;
                }
                asm volatile("bar.sync 3, %0;" ::"n"(kCsConsumers * 32) : "memory");
                if (threadIdx.x == 0) {
                    an = 0, as = 0, aq = 0, az = 0, amn = INFINITY, amx = -INFINITY;
                    for (int w = 0; w < kCsConsumers; ++w) {
                        an += red[w * 8 + 0], as += red[w * 8 + 1], aq += red[w * 8 + 2], az += red[w * 8 + 5];
                        amn = fmin(amn, red[w * 8 + 3]), amx = fmax(amx, red[w * 8 + 4]);
                    }
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

static uint32_t cs_align(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

template <int G>
static int cs_launch_variant(const CsParams& q, const CUtensorMap& vmap, const CUtensorMap& omap, int64_t grid,
                             cudaStream_t stream) {
    static std::atomic<uint32_t> configured{0};  // largest dynamic shared-memory size set so far (per instantiation)
    if (configured.load() < q.smem_total) {
        PPO_CUDA(cudaFuncSetAttribute(tvf_cslab_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(q.smem_total)));
        configured.store(q.smem_total);
    }
    tvf_cslab_kernel<G><<<static_cast<unsigned>(grid), kCsThreads, q.smem_total, stream>>>(q, vmap, omap);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// stats / partials / ticket: optional fused statistics of the longest horizon (workspace layout of tvf.cu).
// partials: workspace for the per-env records of the fused statistics, `partial_doubles` doubles (>= 8 * envs)
int launch_tvf_cslab(const TvfParams& p, double* stats, double* partials, unsigned int* ticket, int64_t partial_doubles,
                     const float* host_value_weights, float* value_out, int64_t value_row_stride, cudaStream_t stream,
                     bool* stats_fused) {
    *stats_fused = false;
    const int N = p.steps, K = p.n_req, V = p.n_val, nmax = p.tab.nmax;
    auto unsupported = [&](const char* why) {
        set_error("ppo_tvf_targets: column slab kernel unavailable: %s", why);
        return PPO_ERR_UNSUPPORTED;
    };
    if (N != 128 && N != 256) return unsupported("steps must be 128 or 256");
    if (V % 32 != 0 || V > 128) return unsupported("value horizons must be 32, 64, 96 or 128");
    // (32 horizons would be one team job per env: the two-team job rotation needs at least two)
    if (K != 64 && K != 96 && K != 128) return unsupported("required horizons must be 64, 96 or 128");
    if (!p.tab.host_blob || !p.tab.dev_blob || !p.tab.gpow) return unsupported("tables carry no host mirror (ppo_tvf_tables_build)");
    if (p.tab.cslab_offset_words <= 0 || p.tab.tslab_offset_words <= 0)
        return unsupported("tables carry no launch constants for this kernel");
    if (!(p.tab.gamma > 0.0)) return unsupported("tables carry no gamma");
    if (p.tab.tail_pitch <= nmax) return unsupported("horizon-major tail tables missing");
    if ((reinterpret_cast<uintptr_t>(p.vs) & 15) || (p.vs_es * 4) % 16 || (p.vs_rs * 4) % 16 || p.vs_es < V)
        return unsupported("value samples are not 16-byte aligned / strided");
    if (reinterpret_cast<uintptr_t>(p.out) & 15) return unsupported("targets are not 16-byte aligned");
    if (p.envs > INT32_MAX) return unsupported("too many envs");

    const int32_t* hdr = static_cast<const int32_t*>(p.tab.host_blob) + p.tab.cslab_offset_words;
    if (hdr[0] != kCsConstMagic || hdr[1] < 0)
        return unsupported("more than 128 horizons / 1024 entries, or interpolation slots that are not adjacent");
    if (hdr[2] != N || hdr[3] != V || hdr[4] != K / 4) return unsupported("launch constants were built for another shape");

    static thread_local CsParams q;  // 18 KB: not on the stack
    q.p = p;
    q.gamma = p.tab.gamma;
    q.ng = V / 32, q.njobs = K / 4, q.n_ent = hdr[1], q.tail_pitch = p.tab.tail_pitch;
    {
        const char* d = getenv("PPO_TVF_DEBUG");
        q.dbg = d ? atoi(d) : 0;
    }
    auto pad4 = [](int64_t w) { return (w + 3) / 4 * 4; };
    const CsHorizon* hzw = reinterpret_cast<const CsHorizon*>(hdr + kCsHeaderWords);
    const uint32_t* jobw = reinterpret_cast<const uint32_t*>(hzw + K);
    const int32_t* vhw = reinterpret_cast<const int32_t*>(jobw + pad4(K / 4));
    const CsEntry* ce = reinterpret_cast<const CsEntry*>(vhw + 128);
    const int64_t ek = static_cast<int64_t>(p.tab.n_entries) * K;
    for (int g = 0; g < kCsGroups; ++g) q.arrivals[g] = hdr[5 + g];
    memcpy(q.hz, hzw, sizeof(CsHorizon) * K);
    memcpy(q.jobmask, jobw, sizeof(uint32_t) * (K / 4));
    memcpy(q.vh, vhw, sizeof(int32_t) * 128);
    memcpy(q.ent, ce, sizeof(CsEntry) * (q.n_ent + 1));
    q.xmax = hdr[10], q.nib_words = hdr[11];
    q.nib_dev = reinterpret_cast<const uint32_t*>(static_cast<const int32_t*>(p.tab.dev_blob) + p.tab.cslab_offset_words +
                                                 (reinterpret_cast<const int32_t*>(ce + ek + 1) - hdr));
    // the interleaved tail records of the transposed slab section (device copy)
    {
        const int64_t E = p.tab.n_entries;
        const int32_t* dts = static_cast<const int32_t*>(p.tab.dev_blob) + p.tab.tslab_offset_words;
        q.tail4 = reinterpret_cast<const uint4*>(dts + (4 + 4 * K + 4 * E * K + 3) / 4 * 4);
    }

    int dev = 0, max_smem = 0;
    PPO_CUDA(cudaGetDevice(&dev));
    PPO_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // slab: V value columns, the -R column, a zero column; gathers may run past the last column by the largest bin
    const uint32_t slab_bytes = cs_align((static_cast<uint32_t>(V + 2) * N + nmax + 16) * 4u, 1024);
    const uint32_t ftab_bytes = cs_align(static_cast<uint32_t>(q.xmax + 8) * 4u, 128);
    const uint32_t nib_bytes = cs_align(static_cast<uint32_t>(q.nib_words + 24) * 4u, 128);
    const uint32_t rest = 2u * 16384u + cs_align(2u * (N + 16) * 4u, 128) + cs_align(2u * (N + 16) * 2u, 128) +
                          cs_align(static_cast<uint32_t>(V) * 4u, 128) + ftab_bytes + nib_bytes + 1024u + 128u +
                          (2 * kCsMaxStages + 2 * kCsGroups + 4 + 8) * 8u;
    int nstage = 0;  // staging ring: as deep as shared memory allows (the TMA loads are latency bound)
    while (nstage < kCsMaxStages && slab_bytes + rest + static_cast<uint32_t>(nstage + 1) * kCsChunkBytes <= static_cast<uint32_t>(max_smem))
        ++nstage;
    if (getenv("PPO_CSLAB_STAGES")) nstage = atoi(getenv("PPO_CSLAB_STAGES")) < nstage ? atoi(getenv("PPO_CSLAB_STAGES")) : nstage;
    if (nstage < 2) return unsupported("the env slab does not fit in shared memory");
    q.nstage = nstage;
    q.off_stage = slab_bytes;
    q.off_tile = q.off_stage + static_cast<uint32_t>(nstage) * kCsChunkBytes;
    q.off_rtab = q.off_tile + 2u * 16384u;
    q.off_lim = q.off_rtab + cs_align(2u * (N + 16) * 4u, 128);
    q.off_fr = q.off_lim + cs_align(2u * (N + 16) * 2u, 128);
    q.off_ftab = q.off_fr + cs_align(static_cast<uint32_t>(V) * 4u, 128);
    q.off_nib = q.off_ftab + ftab_bytes;
    q.off_gpc = q.off_nib + nib_bytes;
    q.off_misc = q.off_gpc + 1024u;
    q.off_bar = q.off_misc + 128;
    q.smem_total = q.off_bar + (2 * kCsMaxStages + 2 * kCsGroups + 4 + 8) * 8u;
    if (q.smem_total > static_cast<uint32_t>(max_smem)) return unsupported("the env slab does not fit in shared memory");

    CUtensorMap vmap;  // value samples [N + 1, A, V]: 64 rows x 32 columns per load, 128-byte swizzle
    {
        const uint64_t dims[3] = {static_cast<uint64_t>(V), static_cast<uint64_t>(p.envs), static_cast<uint64_t>(N) + 1};
        const uint64_t strides[2] = {static_cast<uint64_t>(p.vs_es) * 4, static_cast<uint64_t>(p.vs_rs) * 4};
        const uint32_t box[3] = {32u, 1u, static_cast<uint32_t>(kCsChunkRows)};
        const int rc = encode_tensor_map(&vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.vs, dims, strides, box, true);
        if (rc != PPO_OK) return rc;
    }
    CUtensorMap omap;  // targets [N, A, K] as {k, env, t % 4, t / 4}: 32 horizons x 32 rows of one t % 4 per store
    {
        const uint64_t row_bytes = static_cast<uint64_t>(p.envs) * K * 4;
        const uint64_t odims[4] = {static_cast<uint64_t>(K), static_cast<uint64_t>(p.envs), 4u, static_cast<uint64_t>(N / 4)};
        const uint64_t ostrides[3] = {static_cast<uint64_t>(K) * 4, row_bytes, 4 * row_bytes};
        const uint32_t obox[4] = {32u, 1u, 1u, 32u};
        const int rc = encode_tensor_map(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, odims, ostrides, obox, true);
        if (rc != PPO_OK) return rc;
    }

    int64_t grid = sm_count();
    if (grid > p.envs) grid = p.envs;
    {
        // scheduler pair of this launch (zero before the first launch, zeroed again by the last CTA of every launch)
        static std::atomic<unsigned int> launches{0};
        unsigned int* base = nullptr;
        PPO_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_cs_sched));
        q.sched = base + 2 * (launches.fetch_add(1) % 64);
    }
    q.want_value = value_out ? (host_value_weights ? 2 : 1) : 0;
    q.value_out = value_out, q.vo_rs = value_row_stride;
    memset(q.vw, 0, sizeof(q.vw));
    if (host_value_weights)
        for (int j = 0; j < V; ++j) q.vw[j] = host_value_weights[j];
    q.want_stats = (stats && partials && ticket && partial_doubles >= 8 * p.envs) ? 1 : 0;
    q.stats = stats, q.env_stats = partials, q.ticket = ticket;

    const int rc = N == 256 ? cs_launch_variant<2>(q, vmap, omap, grid, stream) : cs_launch_variant<1>(q, vmap, omap, grid, stream);
    if (rc != PPO_OK) return rc;
    *stats_fused = q.want_stats != 0;
    return PPO_OK;
}

}  // namespace ppo

// development aid: the stamps of the last launch with PPO_TVF_DEBUG bit 256 (nanoseconds, CTA 0)
extern "C" int ppo_tvf_cslab_debug_times(unsigned long long* out832) {
    PPO_CUDA(cudaDeviceSynchronize());
    PPO_CUDA(cudaMemcpyFromSymbol(out832, ppo::g_cs_times, sizeof(unsigned long long) * 832));
    return PPO_OK;
}
