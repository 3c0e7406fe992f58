// Reverse-time scans over [N, A] rollout buffers: GAE + TD(lambda), bootstrapped discounted
// returns, advantage moments and normalisation.
//
// Replaces rl/returns.py:7-67 and the inline normalisation at rl/rollout.py:1887-1897 of the
// reference (see include/ppo_b200.h for the per-entry-point citations).
//
// Parallel structure.  The recurrences are affine in the carried value,
//     x_t = b_t + c_t * x_{t+1},
// and affine maps compose associatively, so the time axis is split as well as the env axis:
// a CTA owns 32 adjacent envs (one per lane, so every global access is a 128-byte coalesced
// row segment) and 16 time chunks (one per warp).  Each thread (1) loads its chunk once into
// registers and composes it into one map (B, C), (2) exchanges maps through shared memory and
// folds the later chunks into its carry-in, (3) replays its chunk from registers and streams the
// outputs.  HBM traffic is exactly the algorithmic bytes: every input is read once, every
// output written once.
//
// Numerics.  With boolean terminals the reference carries float64 and rounds only the stored
// row (rl/returns.py:26-28); the scan below carries float64 as well and performs the replay with
// the reference's operation order and explicit (non-fused) roundings, so outputs differ from the
// reference only through the re-associated carry-in of a chunk (a few float64 ulps before the
// float32 rounding).
#include <stdlib.h>

#include "common.cuh"

namespace ppo {

constexpr int kScanWarps = 16;  // time chunks per CTA, one warp each
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kCachedLen = 16;  // longest chunk kept in registers

struct ScanStatsIo {
    double* stats;         // final record (PPO_STATS_LEN) or nullptr
    double* partials;      // [grid, PPO_STATS_LEN]
    unsigned int* ticket;  // zero on entry, zero again on exit
};

struct GaeParams {
    const float* rewards;
    int64_t r_rs;
    const float* values;
    int64_t v_rs, v_es;
    const float* final_value;
    int64_t f_es;
    const void* dones;
    int64_t d_rs;
    int64_t steps, envs;
    double gl_adv, gl_ret;  // gamma * lamb (float64 product, rl/returns.py:27)
    float gamma32;          // gamma as the float32 scalar of `gamma * value_next_t` (rl/returns.py:26)
    float* adv;
    float* ret;
    int chunk_len;
    ScanStatsIo io;
};

struct ThreadStats {
    double n = 0, s_adv = 0, q_adv = 0, s_ret = 0, q_ret = 0;
    double s_val = 0, q_val = 0, n_val = 0;  // value estimates read by the scan (rl/rollout.py:1205)
    double z_adv = 0, z_ret = 0;             // exact zeros (rl/logger.py:201-229 `sparsity`)
    __device__ __forceinline__ void add_value(float v) {
        s_val += v;
        q_val += static_cast<double>(v) * v;
        n_val += 1.0;
    }
    float mn_adv = INFINITY, mx_adv = -INFINITY, mn_ret = INFINITY, mx_ret = -INFINITY;
    // sums, squares and extrema only: the caller sets `n` from the element count it owns
    __device__ __forceinline__ void add_no_count(float a, float r) {
        z_adv += a == 0.f ? 1.0 : 0.0;
        z_ret += r == 0.f ? 1.0 : 0.0;
        s_adv += a;
        q_adv += static_cast<double>(a) * a;
        s_ret += r;
        q_ret += static_cast<double>(r) * r;
        mn_adv = fminf(mn_adv, a);
        mx_adv = fmaxf(mx_adv, a);
        mn_ret = fminf(mn_ret, r);
        mx_ret = fmaxf(mx_ret, r);
    }
    __device__ __forceinline__ void add(float a, float r) {
        n += 1.0;
        z_adv += a == 0.f ? 1.0 : 0.0;
        z_ret += r == 0.f ? 1.0 : 0.0;
        s_adv += a;
        q_adv += static_cast<double>(a) * a;
        s_ret += r;
        q_ret += static_cast<double>(r) * r;
        mn_adv = fminf(mn_adv, a);
        mx_adv = fmaxf(mx_adv, a);
        mn_ret = fminf(mn_ret, r);
        mx_ret = fmaxf(mx_ret, r);
    }
};

// Exchange of the per-GPU statistics records over NVLink peer memory (env-sharded runs): every rank owns a mailbox
// [2][world][16] doubles and a flag word per (parity, sender), mapped into its peers; the CTA that folds the local record
// stores it into slot `rank` of every mailbox (its own included) and then raises flag `rank` there to the launch's
// epoch.  Records and flags are double-buffered by the parity of the epoch: a rank that is one launch ahead writes the
// other half, so it cannot overwrite a record a slower rank's CTAs are still reading (it cannot be two launches ahead:
// its launch k + 1 completes only after every rank has raised flag k + 1, i.e. has finished launch k).
struct PeerExchange {
    int world, rank;               // world <= kMaxPeers; world == 0: no exchange
    unsigned long long epoch;      // strictly increasing per launch: flags never need a reset
    double* mail[PPO_MAX_PEERS];              // mail[p]: rank p's mailbox, [2][world][PPO_STATS_LEN]
    unsigned long long* flags[PPO_MAX_PEERS];  // flags[p]: rank p's flag words, [2][world]
    double* global_stats;          // merged record of all ranks (optional), written by CTA 0
    unsigned long long* status;    // optional error word (device or mapped host memory): epoch of a timed-out wait
    unsigned long long timeout_ns;
    __device__ __forceinline__ int half() const { return static_cast<int>(epoch & 1ull) * world; }
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Combine per-thread statistics into one record per CTA, then let the last CTA to finish fold all
// CTA records in index order (deterministic, no float atomics) into io.stats.
__device__ void publish_stats(const ThreadStats& ts, const ScanStatsIo& io, double (*red)[PPO_STATS_LEN],
                              unsigned int* ready = nullptr, const PeerExchange* px = nullptr) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    double n = warp_sum_f64(ts.n);
    double s_adv = warp_sum_f64(ts.s_adv), q_adv = warp_sum_f64(ts.q_adv);
    double s_ret = warp_sum_f64(ts.s_ret), q_ret = warp_sum_f64(ts.q_ret);
    float mn_adv = warp_min_f32(ts.mn_adv), mx_adv = warp_max_f32(ts.mx_adv);
    float mn_ret = warp_min_f32(ts.mn_ret), mx_ret = warp_max_f32(ts.mx_ret);
    double s_val = warp_sum_f64(ts.s_val), q_val = warp_sum_f64(ts.q_val), n_val = warp_sum_f64(ts.n_val);
    double z_adv = warp_sum_f64(ts.z_adv), z_ret = warp_sum_f64(ts.z_ret);
    if (lane == 0) {
        red[w][PPO_STAT_ADV_ZEROS] = z_adv;
        red[w][PPO_STAT_RET_ZEROS] = z_ret;
        red[w][PPO_STAT_VAL_SUM] = s_val;
        red[w][PPO_STAT_VAL_SUMSQ] = q_val;
        red[w][PPO_STAT_VAL_COUNT] = n_val;
        red[w][PPO_STAT_COUNT] = n;
        red[w][PPO_STAT_ADV_SUM] = s_adv;
        red[w][PPO_STAT_ADV_SUMSQ] = q_adv;
        red[w][PPO_STAT_ADV_MIN] = mn_adv;
        red[w][PPO_STAT_ADV_MAX] = mx_adv;
        red[w][PPO_STAT_RET_SUM] = s_ret;
        red[w][PPO_STAT_RET_SUMSQ] = q_ret;
        red[w][PPO_STAT_RET_MIN] = mn_ret;
        red[w][PPO_STAT_RET_MAX] = mx_ret;
    }
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x < PPO_STATS_LEN) {
        const int j = threadIdx.x;
        const bool is_min = (j == PPO_STAT_ADV_MIN || j == PPO_STAT_RET_MIN);
        const bool is_max = (j == PPO_STAT_ADV_MAX || j == PPO_STAT_RET_MAX);
        double acc = 0;
        if (j <= PPO_STAT_RET_ZEROS) {
            acc = red[0][j];
            for (int i = 1; i < nwarps; ++i) {
                const double x = red[i][j];
                acc = is_min ? fmin(acc, x) : is_max ? fmax(acc, x) : acc + x;
            }
        }
        io.partials[static_cast<int64_t>(blockIdx.x) * PPO_STATS_LEN + j] = acc;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(io.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // Last CTA: fold all CTA records.  The records are [grid][16] doubles, so consecutive threads read consecutive
    // words: thread t owns slot t % 16 of records t / 16, t / 16 + groups, ... (loads issued 32 at a time), and
    // the per-group partials are combined in group order.  Fixed order everywhere: the result does not depend on
    // scheduling.
    {
        const int j = threadIdx.x & (PPO_STATS_LEN - 1);
        const int g = threadIdx.x / PPO_STATS_LEN, groups = blockDim.x / PPO_STATS_LEN;
        const bool is_min = (j == PPO_STAT_ADV_MIN || j == PPO_STAT_RET_MIN);
        const bool is_max = (j == PPO_STAT_ADV_MAX || j == PPO_STAT_RET_MAX);
        double acc = is_min ? INFINITY : is_max ? -INFINITY : 0.0;
        const unsigned int G = gridDim.x;
        constexpr int kFoldBatch = 32;  // loads in flight per thread: the fold is a latency chain at the very end
        for (unsigned int b0 = g; b0 < G; b0 += kFoldBatch * groups) {
            double x[kFoldBatch];
#pragma unroll
            for (int u = 0; u < kFoldBatch; ++u) {
                const unsigned int b = b0 + u * groups;
                x[u] = b < G ? __ldcg(io.partials + static_cast<int64_t>(b) * PPO_STATS_LEN + j)
                             : (is_min ? INFINITY : is_max ? -INFINITY : 0.0);
            }
#pragma unroll
            for (int u = 0; u < kFoldBatch; ++u) acc = is_min ? fmin(acc, x[u]) : is_max ? fmax(acc, x[u]) : acc + x[u];
        }
        __syncthreads();  // every warp has passed its use of `red` above
        double* fold = &red[0][0];  // [groups][16], groups <= nwarps * 2
        fold[g * PPO_STATS_LEN + j] = acc;
        __syncthreads();
        if (threadIdx.x < PPO_STATS_LEN) {
            double tot = fold[j];
            for (int i = 1; i < groups; ++i) {
                const double x = fold[i * PPO_STATS_LEN + j];
                tot = is_min ? fmin(tot, x) : is_max ? fmax(tot, x) : tot + x;
            }
            io.stats[j] = j <= PPO_STAT_RET_ZEROS ? tot : 0.0;
            __threadfence();
            if (px) {  // the local record goes to slot `rank` of every mailbox (remote stores over NVLink)
                for (int r = 0; r < px->world; ++r)
                    px->mail[r][(px->half() + px->rank) * PPO_STATS_LEN + j] = j <= PPO_STAT_RET_ZEROS ? tot : 0.0;
                __threadfence_system();
            }
        }
        if (px) {
            __syncthreads();  // all sixteen slots are on their way; the release stores order them before the flags
            if (threadIdx.x < px->world) st_release_sys(px->flags[threadIdx.x] + px->half() + px->rank, px->epoch);
        }
        if (ready) {  // the folded record is visible: release the CTAs that wait for it
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(ready, 1u);
        }
    }
    if (threadIdx.x == 0) *io.ticket = 0u;
}

template <int DONE, bool CACHED, bool STATS>
__global__ void __launch_bounds__(kScanThreads) gae_kernel(const GaeParams p) {
    __shared__ double sB[2][kScanWarps][32];
    __shared__ double sC[2][kScanWarps][32];
    __shared__ double red[2 * kScanWarps][PPO_STATS_LEN];  // per-warp records, re-used as [threads / 16][16] by the final fold

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * 32 + lane;
    const bool live = env < p.envs;
    const int64_t lo = static_cast<int64_t>(w) * p.chunk_len;
    const int64_t hi = min(p.steps, lo + p.chunk_len);
    const int len = hi > lo ? static_cast<int>(hi - lo) : 0;  // warp-uniform

    const float* rp = p.rewards + lo * p.r_rs + env;
    const float* vp = p.values + lo * p.v_rs + env * p.v_es;
    const int64_t dbase = lo * p.d_rs + env;

    float r[kCachedLen], v[kCachedLen], al[kCachedLen];
    float vtop = 0.f;  // value following the last step of this chunk
    if (live && len > 0) {
        vtop = (hi == p.steps) ? p.final_value[env * p.f_es] : p.values[hi * p.v_rs + env * p.v_es];
    }
    if (CACHED) {
#pragma unroll
        for (int i = 0; i < kCachedLen; ++i) {
            r[i] = 0.f, v[i] = 0.f, al[i] = 1.f;
            if (live && i < len) {
                r[i] = __ldg(rp + i * p.r_rs);
                v[i] = __ldg(vp + i * p.v_rs);
                al[i] = static_cast<float>(alive_of<DONE>(p.dones, dbase + i * p.d_rs));
            }
        }
    }

    // one step of both recurrences, reference operation order (rl/returns.py:26-28)
    auto step = [&](float rr, float vv, float vn, double nt, double& x_adv, double& x_ret) {
        const double boot = static_cast<double>(__fmul_rn(p.gamma32, vn)) * nt;
        const double delta = __dsub_rn(__dadd_rn(static_cast<double>(rr), boot), static_cast<double>(vv));
        x_adv = __dadd_rn(delta, __dmul_rn(__dmul_rn(p.gl_adv, nt), x_adv));
        x_ret = __dadd_rn(delta, __dmul_rn(__dmul_rn(p.gl_ret, nt), x_ret));
    };

    // ---- phase 1: compose the chunk into x_lo = B + C * x_in
    double b_adv = 0, c_adv = 1, b_ret = 0, c_ret = 1;
    if (live) {
        if (CACHED) {
#pragma unroll
            for (int i = kCachedLen - 1; i >= 0; --i) {
                if (i < len) {
                    const float vn = (i == len - 1) ? vtop : v[(i + 1) % kCachedLen];
                    const double nt = al[i];
                    step(r[i], v[i], vn, nt, b_adv, b_ret);
                    c_adv *= p.gl_adv * nt;
                    c_ret *= p.gl_ret * nt;
                }
            }
        } else {
            float vn = vtop;
            for (int i = len - 1; i >= 0; --i) {
                const float rr = __ldg(rp + i * p.r_rs), vv = __ldg(vp + i * p.v_rs);
                const double nt = alive_of<DONE>(p.dones, dbase + i * p.d_rs);
                step(rr, vv, vn, nt, b_adv, b_ret);
                c_adv *= p.gl_adv * nt;
                c_ret *= p.gl_ret * nt;
                vn = vv;
            }
        }
    }
    sB[0][w][lane] = b_adv, sC[0][w][lane] = c_adv;
    sB[1][w][lane] = b_ret, sC[1][w][lane] = c_ret;
    __syncthreads();

    // ---- phase 2: carry-in = later chunks applied to A_N = 0
    double x_adv = 0, x_ret = 0;
    for (int j = kScanWarps - 1; j > w; --j) {
        x_adv = sB[0][j][lane] + sC[0][j][lane] * x_adv;
        x_ret = sB[1][j][lane] + sC[1][j][lane] * x_ret;
    }

    // ---- phase 3: replay and stream the outputs
    ThreadStats ts;
    if (live) {
        float* ap = p.adv ? p.adv + lo * p.envs + env : nullptr;
        float* tp = p.ret ? p.ret + lo * p.envs + env : nullptr;
        if (CACHED) {
#pragma unroll
            for (int i = kCachedLen - 1; i >= 0; --i) {
                if (i < len) {
                    const float vn = (i == len - 1) ? vtop : v[(i + 1) % kCachedLen];
                    step(r[i], v[i], vn, al[i], x_adv, x_ret);
                    const float a = static_cast<float>(x_adv);
                    const float t = __fadd_rn(static_cast<float>(x_ret), v[i]);  // rl/returns.py:66-67
                    if (ap) store_stream(ap + i * p.envs, a);
                    if (tp) store_stream(tp + i * p.envs, t);
                    if (STATS) ts.add(a, t), ts.add_value(v[i]);
                }
            }
        } else {
            float vn = vtop;
            for (int i = len - 1; i >= 0; --i) {
                const float rr = __ldg(rp + i * p.r_rs), vv = __ldg(vp + i * p.v_rs);
                const double nt = alive_of<DONE>(p.dones, dbase + i * p.d_rs);
                step(rr, vv, vn, nt, x_adv, x_ret);
                const float a = static_cast<float>(x_adv);
                const float t = __fadd_rn(static_cast<float>(x_ret), vv);
                if (ap) store_stream(ap + i * p.envs, a);
                if (tp) store_stream(tp + i * p.envs, t);
                if (STATS) ts.add(a, t), ts.add_value(vv);
                vn = vv;
            }
        }
        if (STATS && len > 0 && hi == p.steps) ts.add_value(vtop);  // the final estimate, row N
    }
    if (STATS) publish_stats(ts, p.io, red);
}

// ------------------------------------------------------------------------------------------
// Tile form of the same scan for bandwidth: a CTA stages its whole [N, EW] tile of rewards, values
// and terminals in shared memory with asynchronous copies (cp.async: no registers, nothing waits
// until the tile is complete), so several CTAs per SM keep loads in flight while others compute,
// and the per-thread state is a handful of registers instead of the cached chunk.  EW envs (16 or
// 32) x kTileChunks time chunks per CTA; the arithmetic is the one of gae_kernel (float64 carries,
// reference operation order), with the 0/1 terminal mask applied as a select instead of a multiply
// and the two recurrences merged when lambda_policy == lambda_value.
constexpr int kTileChunks = 8;

__device__ __forceinline__ float norm_one(float x, float mean, float denom, float clip) {
    float y = __fdiv_rn(__fsub_rn(x, mean), denom);  // rl/rollout.py:1891
    if (clip > 0.f) y = fminf(fmaxf(y, -clip), clip);  // rl/rollout.py:1895-1896
    return y;
}


struct GaeTileParams {
    GaeParams g;
    int chunk_len;      // rows per time chunk (odd when EW = 16: the two chunks of a warp then hit different banks)
    int vec_r, vec_v, vec_d;  // bytes per asynchronous copy: 16 / 4, and 1 (synchronous byte loads) for the terminals
    int same_lambda;
    // fused advantage normalisation (cooperative launch, every CTA resident): the CTAs wait for the folded statistics
    // and normalise their own tile of advantages out of shared memory
    float* norm_out;           // [N * A] or nullptr
    float norm_eps, norm_clip;
    unsigned int* ready;       // two words: statistics-ready flag, exit counter (both zero between launches)
    PeerExchange px;           // world > 1: the moments come from every rank's record (ppo_gae_normalized_peers)
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4b(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <int DONE, int EW, bool STATS, bool NORM = false>
__global__ void __launch_bounds__(EW * kTileChunks) gae_tile_kernel(const GaeTileParams q) {
    extern __shared__ __align__(16) unsigned char tile[];
    const GaeParams& p = q.g;
    constexpr int T = EW * kTileChunks;
    constexpr int DB = DONE == PPO_DONE_U8 ? 1 : DONE == PPO_DONE_F32 ? 4 : 0;  // bytes per terminal
    const int N = static_cast<int>(p.steps);
    float* sr = reinterpret_cast<float*>(tile);                  // [N][EW]
    float* sv = sr + static_cast<size_t>(N) * EW;                // [N + 1][EW], row N = final estimate
    unsigned char* sd = reinterpret_cast<unsigned char*>(sv + static_cast<size_t>(N + 1) * EW);  // [N][EW] x DB
    double* xB = reinterpret_cast<double*>(sd + ((static_cast<size_t>(N) * EW * DB + 15) / 16) * 16);  // [2][chunks][EW]
    double* xC = xB + 2 * kTileChunks * EW;
    double (*red)[PPO_STATS_LEN] = reinterpret_cast<double (*)[PPO_STATS_LEN]>(xC + 2 * kTileChunks * EW);

    const int tid = threadIdx.x;
    const int64_t env0 = static_cast<int64_t>(blockIdx.x) * EW;
    const uint32_t sr_a = static_cast<uint32_t>(__cvta_generic_to_shared(sr));
    const uint32_t sv_a = static_cast<uint32_t>(__cvta_generic_to_shared(sv));
    const uint32_t sd_a = static_cast<uint32_t>(__cvta_generic_to_shared(sd));

    // ---- stage the tile
    if (q.vec_r == 16) {
        for (int i = tid; i < N * (EW / 4); i += T) {
            const int row = i / (EW / 4), seg = i % (EW / 4);
            if (env0 + seg * 4 < p.envs) cp_async_16(sr_a + (row * EW + seg * 4) * 4, p.rewards + row * p.r_rs + env0 + seg * 4);
        }
    } else {
        for (int i = tid; i < N * EW; i += T) {
            const int row = i / EW, e = i % EW;
            if (env0 + e < p.envs) cp_async_4b(sr_a + i * 4, p.rewards + row * p.r_rs + env0 + e);
        }
    }
    if (q.vec_v == 16) {
        for (int i = tid; i < N * (EW / 4); i += T) {
            const int row = i / (EW / 4), seg = i % (EW / 4);
            if (env0 + seg * 4 < p.envs) cp_async_16(sv_a + (row * EW + seg * 4) * 4, p.values + row * p.v_rs + env0 + seg * 4);
        }
    } else if (p.v_es == 1) {
        for (int i = tid; i < N * EW; i += T) {
            const int row = i / EW, e = i % EW;
            if (env0 + e < p.envs) cp_async_4b(sv_a + i * 4, p.values + row * p.v_rs + env0 + e);
        }
    } else {
        // strided value column (e.g. the last horizon of a [N+1, A, K] value buffer, rl/tvf.py:325-340): one sector
        // per element, a latency-bound gather -> plain loads, sixteen in flight per thread
        constexpr int U = 16;
        for (int i0 = tid; i0 < N * EW; i0 += T * U) {
            float x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * T, row = i / EW, e = i % EW;
                x[u] = (i < N * EW && env0 + e < p.envs) ? __ldg(p.values + row * p.v_rs + (env0 + e) * p.v_es) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * T;
                if (i < N * EW) sv[i] = x[u];
            }
        }
    }
    if (tid < EW && env0 + tid < p.envs) cp_async_4b(sv_a + (N * EW + tid) * 4, p.final_value + (env0 + tid) * p.f_es);
    if (DONE != PPO_DONE_NONE) {
        const unsigned char* dn = static_cast<const unsigned char*>(p.dones);
        constexpr int RB = EW * DB;  // bytes per tile row
        if (q.vec_d == 16) {
            for (int i = tid; i < N * (RB / 16); i += T) {
                const int row = i / (RB / 16), seg = i % (RB / 16);
                if (env0 * DB + seg * 16 < p.envs * DB)
                    cp_async_16(sd_a + row * RB + seg * 16, dn + (row * p.d_rs + env0) * DB + seg * 16);
            }
        } else if (q.vec_d == 4) {
            for (int i = tid; i < N * (RB / 4); i += T) {
                const int row = i / (RB / 4), seg = i % (RB / 4);
                if (env0 * DB + seg * 4 < p.envs * DB)
                    cp_async_4b(sd_a + row * RB + seg * 4, dn + (row * p.d_rs + env0) * DB + seg * 4);
            }
        } else {
            for (int i = tid; i < N * EW; i += T) {
                const int row = i / EW, e = i % EW;
                if (env0 + e < p.envs) sd[i] = dn[row * p.d_rs + env0 + e];
            }
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int e = tid % EW, w = tid / EW;
    const int64_t env = env0 + e;
    const bool live = env < p.envs;
    const int lo = w * q.chunk_len;
    const int hi = min(N, lo + q.chunk_len);
    const bool same = q.same_lambda != 0;

    // terminal -> (boot, k_adv, k_ret): 0/1 masks are selects (bit-identical to the reference's multiply by 0.0 / 1.0)
    auto coeffs = [&](int i, float vn, double& boot, double& ka, double& kr) {
        const double gv = static_cast<double>(__fmul_rn(p.gamma32, vn));
        if (DONE == PPO_DONE_U8) {
            const bool d = sd[i * EW + e] != 0;
            boot = d ? 0.0 : gv, ka = d ? 0.0 : p.gl_adv, kr = d ? 0.0 : p.gl_ret;
        } else if (DONE == PPO_DONE_F32) {
            const double nt = static_cast<double>(1.0f - reinterpret_cast<const float*>(sd)[i * EW + e]);
            boot = gv * nt, ka = __dmul_rn(p.gl_adv, nt), kr = __dmul_rn(p.gl_ret, nt);
        } else {
            boot = gv, ka = p.gl_adv, kr = p.gl_ret;
        }
    };

    // ---- phase 1: compose the chunk into x_lo = B + C * x_in
    double b_adv = 0, c_adv = 1, b_ret = 0, c_ret = 1;
    if (live) {
        for (int i = hi - 1; i >= lo; --i) {
            double boot, ka, kr;
            coeffs(i, sv[(i + 1) * EW + e], boot, ka, kr);
            const double delta = __dsub_rn(__dadd_rn(static_cast<double>(sr[i * EW + e]), boot),
                                           static_cast<double>(sv[i * EW + e]));
            b_adv = __dadd_rn(delta, __dmul_rn(ka, b_adv));
            c_adv *= ka;
            if (!same) {
                b_ret = __dadd_rn(delta, __dmul_rn(kr, b_ret));
                c_ret *= kr;
            }
        }
    }
    xB[w * EW + e] = b_adv, xC[w * EW + e] = c_adv;
    if (!same) xB[(kTileChunks + w) * EW + e] = b_ret, xC[(kTileChunks + w) * EW + e] = c_ret;
    __syncthreads();

    // ---- phase 2: carry-in = later chunks applied to A_N = 0
    double x_adv = 0, x_ret = 0;
    for (int j = kTileChunks - 1; j > w; --j) {
        x_adv = xB[j * EW + e] + xC[j * EW + e] * x_adv;
        if (!same) x_ret = xB[(kTileChunks + j) * EW + e] + xC[(kTileChunks + j) * EW + e] * x_ret;
    }

    // ---- phase 3: replay and stream the outputs
    ThreadStats ts;
    if (live) {
        for (int i = hi - 1; i >= lo; --i) {
            double boot, ka, kr;
            const float vv = sv[i * EW + e];
            coeffs(i, sv[(i + 1) * EW + e], boot, ka, kr);
            const double delta = __dsub_rn(__dadd_rn(static_cast<double>(sr[i * EW + e]), boot), static_cast<double>(vv));
            x_adv = __dadd_rn(delta, __dmul_rn(ka, x_adv));
            if (!same) x_ret = __dadd_rn(delta, __dmul_rn(kr, x_ret));
            const float a = static_cast<float>(x_adv);
            const float t = __fadd_rn(same ? a : static_cast<float>(x_ret), vv);  // rl/returns.py:66-67
            if (p.adv) store_stream(p.adv + static_cast<int64_t>(i) * p.envs + env, a);
            if (p.ret) store_stream(p.ret + static_cast<int64_t>(i) * p.envs + env, t);
            if (STATS) ts.add_no_count(a, t), ts.add_value(vv);
            if (NORM) sr[i * EW + e] = a;  // the reward slot is dead: the raw advantage waits here for the statistics
        }
        ts.n = hi > lo ? static_cast<double>(hi - lo) : 0.0;
        if (STATS && hi > lo && hi == N) ts.add_value(sv[N * EW + e]);  // the final estimate, row N
    }
    const bool peers = NORM && q.px.world > 1;
    if (STATS) publish_stats(ts, p.io, red, NORM && !peers ? q.ready : nullptr, peers ? &q.px : nullptr);
    if (NORM) {
        // Grid-wide hand-off: the CTA that folds the records raises the flag (publish_stats), everyone else spins on
        // it -- legal because the launch is cooperative (all CTAs resident).  Then (x - mean) / (std + eps) exactly as
        // normalize_kernel computes it (rl/rollout.py:1887-1897), from the tile in shared memory.
        // Env-sharded (peers): the flags are the ones the folding CTAs of ALL ranks raise in this rank's mailbox; the
        // moments are summed in rank order, so every rank normalises with the same bits.
        __shared__ double s_mom[3];
        if (threadIdx.x == 0) {
            if (!peers) {
                while (*reinterpret_cast<volatile unsigned int*>(q.ready) == 0u) __nanosleep(32);
                __threadfence();
                s_mom[0] = __ldcg(p.io.stats + PPO_STAT_COUNT);
                s_mom[1] = __ldcg(p.io.stats + PPO_STAT_ADV_SUM);
                s_mom[2] = __ldcg(p.io.stats + PPO_STAT_ADV_SUMSQ);
            } else {
                const unsigned long long* mine = q.px.flags[q.px.rank] + q.px.half();
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
                bool ok = true;
                for (int r = 0; r < q.px.world && ok; ++r) {
                    while (ld_acquire_sys(mine + r) < q.px.epoch) {
                        __nanosleep(64);
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                        if (t1 - t0 > q.px.timeout_ns) {  // a peer never arrived: fail loudly instead of hanging
                            ok = false;
                            break;
                        }
                    }
                }
                if (!ok && q.px.status) {  // the host checks this word (ReturnPass.check_peers); NaN advantages besides
                    *reinterpret_cast<volatile unsigned long long*>(q.px.status) = q.px.epoch;
                    __threadfence_system();
                }
                const volatile double* mail = q.px.mail[q.px.rank] + q.px.half() * PPO_STATS_LEN;
                double n = 0.0, sm = 0.0, sq = 0.0;
                for (int r = 0; r < q.px.world; ++r) {
                    n += mail[r * PPO_STATS_LEN + PPO_STAT_COUNT];
                    sm += mail[r * PPO_STATS_LEN + PPO_STAT_ADV_SUM];
                    sq += mail[r * PPO_STATS_LEN + PPO_STAT_ADV_SUMSQ];
                }
                s_mom[0] = ok ? n : NAN, s_mom[1] = sm, s_mom[2] = sq;
            }
        }
        __syncthreads();
        if (peers && blockIdx.x == 0 && threadIdx.x < PPO_STATS_LEN && q.px.global_stats) {
            // merged record of all ranks, in rank order (same reduction as ppo_merge_stats)
            const int j = threadIdx.x;
            const bool is_min = (j == PPO_STAT_ADV_MIN || j == PPO_STAT_RET_MIN);
            const bool is_max = (j == PPO_STAT_ADV_MAX || j == PPO_STAT_RET_MAX);
            const volatile double* mail = q.px.mail[q.px.rank] + q.px.half() * PPO_STATS_LEN;
            double acc = is_min ? INFINITY : is_max ? -INFINITY : 0.0;
            for (int r = 0; r < q.px.world; ++r) {
                const double x = mail[r * PPO_STATS_LEN + j];
                acc = is_min ? fmin(acc, x) : is_max ? fmax(acc, x) : acc + x;
            }
            q.px.global_stats[j] = j <= PPO_STAT_RET_ZEROS ? acc : 0.0;
        }
        const double n = s_mom[0];
        const double mean64 = s_mom[1] / n;
        const double var64 = fmax(s_mom[2] / n - mean64 * mean64, 0.0);
        const float mean = static_cast<float>(mean64);
        const float denom = __fadd_rn(static_cast<float>(sqrt(var64)), q.norm_eps);
        if (live) {
            for (int i = hi - 1; i >= lo; --i)
                store_stream(q.norm_out + static_cast<int64_t>(i) * p.envs + env,
                             norm_one(sr[i * EW + e], mean, denom, q.norm_clip));
        }
        // the last CTA to leave clears the flag and the counter for the next launch
        __syncthreads();
        if (threadIdx.x == 0 && !peers) {
            if (atomicAdd(q.ready + 1, 1u) == gridDim.x - 1) {
                q.ready[0] = 0u;
                q.ready[1] = 0u;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Bootstrapped discounted returns (rl/returns.py:32-55): R_t = r_t + (R_{t+1} * g_t) * (1 - done_t).
struct ReturnsParams {
    const float* rewards;
    int64_t r_rs;
    const void* dones;
    int64_t d_rs;
    const float* final_value;
    int64_t f_es;
    const float* gamma_array;  // nullptr: scalar
    int64_t g_rs;
    float gamma32;
    int64_t steps, envs;
    float* out;
    int chunk_len;
};

template <int DONE>
__global__ void __launch_bounds__(kScanThreads) returns_kernel(const ReturnsParams p) {
    __shared__ double sB[kScanWarps][32];
    __shared__ double sC[kScanWarps][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * 32 + lane;
    const bool live = env < p.envs;
    const int64_t lo = static_cast<int64_t>(w) * p.chunk_len;
    const int64_t hi = min(p.steps, lo + p.chunk_len);
    const int len = hi > lo ? static_cast<int>(hi - lo) : 0;
    const bool top = (hi == p.steps) && len > 0;  // the chunk that starts from the final value

    const float* rp = p.rewards + lo * p.r_rs + env;
    const float* gp = p.gamma_array ? p.gamma_array + lo * p.g_rs + env : nullptr;
    const int64_t dbase = lo * p.d_rs + env;
    const float fin = (live && top) ? p.final_value[env * p.f_es] : 0.f;

    // The reference's first product is float32 * float32 (final estimate times a float32 gamma
    // row, rl/returns.py:53); every later one is float64 * float32.
    auto run = [&](double x, bool from_final, bool store) {
        for (int i = len - 1; i >= 0; --i) {
            const float g = gp ? __ldg(gp + i * p.g_rs) : p.gamma32;
            const double nt = alive_of<DONE>(p.dones, dbase + i * p.d_rs);
            const double prod = (from_final && i == len - 1)
                                    ? static_cast<double>(__fmul_rn(static_cast<float>(x), g))
                                    : __dmul_rn(x, static_cast<double>(g));
            x = __dadd_rn(static_cast<double>(__ldg(rp + i * p.r_rs)), __dmul_rn(prod, nt));
            if (store) store_stream(p.out + (lo + i) * p.envs + env, static_cast<float>(x));
        }
        return x;
    };

    double b = 0, c = 1;
    if (live && len > 0) {
        if (top) {
            b = run(static_cast<double>(fin), true, false);
            c = 0;
        } else {
            for (int i = len - 1; i >= 0; --i) {
                const float g = gp ? __ldg(gp + i * p.g_rs) : p.gamma32;
                const double k = static_cast<double>(g) * alive_of<DONE>(p.dones, dbase + i * p.d_rs);
                b = static_cast<double>(__ldg(rp + i * p.r_rs)) + k * b;
                c *= k;
            }
        }
    }
    sB[w][lane] = b, sC[w][lane] = c;
    __syncthreads();
    if (!live || len == 0) return;
    if (top) {
        run(static_cast<double>(fin), true, true);
    } else {
        double x = 0;
        for (int j = kScanWarps - 1; j > w; --j) x = sB[j][lane] + sC[j][lane] * x;
        run(x, false, true);
    }
}

// ------------------------------------------------------------------------------------------
// Moments of a float32 vector (same record layout, ADV_* slots) and the normalisation itself.
struct MomentsParams {
    const float* x;
    int64_t count;
    ScanStatsIo io;
};

__global__ void __launch_bounds__(512) moments_kernel(const MomentsParams p) {
    __shared__ double red[32][PPO_STATS_LEN];
    ThreadStats ts;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.count; i += stride) {
        ts.add(__ldg(p.x + i), 0.f);
    }
    publish_stats(ts, p.io, red);
}

// Extrinsic + intrinsic advantage (rl/rollout.py:1225-1231): out = ext + float32(scale) * int, with the statistics of
// the sum in the ADV_* slots (what the normalisation needs) and of the scaled intrinsic part in the RET_* slots
// (`adv_int`, `adv_ratio` of the reference's log).
struct CombineParams {
    const float* ext;
    const float* intr;
    float scale;
    float* out;
    int64_t count;
    ScanStatsIo io;
};

__global__ void __launch_bounds__(512) adv_combine_kernel(const CombineParams p) {
    __shared__ double red[32][PPO_STATS_LEN];
    ThreadStats ts;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.count; i += stride) {
        const float b = __fmul_rn(p.scale, __ldg(p.intr + i));
        const float x = __fadd_rn(__ldg(p.ext + i), b);
        p.out[i] = x;
        ts.add(x, b);
    }
    publish_stats(ts, p.io, red);
}

struct NormParams {
    const float* x;
    float* out;
    int64_t count;
    const double* stats;
    float eps;
    float clip;  // <= 0: no clipping
};

__global__ void __launch_bounds__(256) normalize_kernel(const NormParams p) {
    const double n = p.stats[PPO_STAT_COUNT];
    const double mean64 = p.stats[PPO_STAT_ADV_SUM] / n;
    const double var64 = fmax(p.stats[PPO_STAT_ADV_SUMSQ] / n - mean64 * mean64, 0.0);
    const float mean = static_cast<float>(mean64);
    const float denom = __fadd_rn(static_cast<float>(sqrt(var64)), p.eps);
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0;
    const int64_t n4 = vec ? p.count / 4 : 0;
    for (int64_t i = tid; i < n4; i += stride) {
        float4 q = __ldg(reinterpret_cast<const float4*>(p.x) + i);
        q.x = norm_one(q.x, mean, denom, p.clip);
        q.y = norm_one(q.y, mean, denom, p.clip);
        q.z = norm_one(q.z, mean, denom, p.clip);
        q.w = norm_one(q.w, mean, denom, p.clip);
        __stcs(reinterpret_cast<float4*>(p.out) + i, q);
    }
    for (int64_t i = n4 * 4 + tid; i < p.count; i += stride) {
        p.out[i] = norm_one(__ldg(p.x + i), mean, denom, p.clip);
    }
}

// ------------------------------------------------------------------------------------------
// host side

static int64_t scan_grid(int64_t envs) { return (envs + 31) / 32; }

static int split_workspace(void* workspace, int64_t bytes, int64_t blocks, ScanStatsIo* io) {
    const int64_t need = 256 + blocks * PPO_STATS_LEN * static_cast<int64_t>(sizeof(double));
    if (workspace == nullptr || bytes < need) {
        set_error("workspace of %lld bytes required, got %lld", (long long)need, (long long)bytes);
        return PPO_ERR_WORKSPACE;
    }
    io->ticket = static_cast<unsigned int*>(workspace);
    io->partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    return PPO_OK;
}

constexpr int kMomentsMaxBlocks = 1184;  // 8 CTAs of 512 threads per SM would not fit; 148 * 8 is the cap we size for

static thread_local bool t_tile_dry_run = false;  // ppo_gae_normalized_fits: decide, do not launch

template <int DONE, int EW>
static int launch_gae_tile_variant(const GaeTileParams& q, bool stats, int64_t grid, size_t smem, cudaStream_t s) {
    if (q.norm_out) {
        // fused normalisation: every CTA must be resident (they wait for the folded statistics), so the launch is
        // cooperative and only made when the whole grid fits
        auto kernel = gae_tile_kernel<DONE, EW, true, true>;
        PPO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int per_sm = 0;
        PPO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, EW * kTileChunks, smem));
        if (static_cast<int64_t>(per_sm) * sm_count() < grid) return PPO_ERR_UNSUPPORTED;
        if (t_tile_dry_run) return PPO_OK;
        GaeTileParams copy = q;
        void* args[] = {&copy};
        PPO_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(static_cast<unsigned>(grid)),
                                             dim3(EW * kTileChunks), args, smem, s));
        return PPO_OK;
    }
    if (stats) {
        PPO_CUDA(cudaFuncSetAttribute(gae_tile_kernel<DONE, EW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        gae_tile_kernel<DONE, EW, true><<<static_cast<unsigned>(grid), EW * kTileChunks, smem, s>>>(q);
    } else {
        PPO_CUDA(cudaFuncSetAttribute(gae_tile_kernel<DONE, EW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        gae_tile_kernel<DONE, EW, false><<<static_cast<unsigned>(grid), EW * kTileChunks, smem, s>>>(q);
    }
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// The tile kernel when a whole [N, EW] tile fits in shared memory; PPO_ERR_UNSUPPORTED otherwise (the caller then
// launches gae_kernel).  `p.io.stats` must already be set when statistics are wanted.
static int launch_gae_tile(GaeParams p, int done_kind, bool want_stats, void* workspace, int64_t workspace_bytes,
                           cudaStream_t s, float* norm_out = nullptr, float norm_eps = 0.f, float norm_clip = 0.f,
                           const PeerExchange* px = nullptr) {
    const int64_t N = p.steps, A = p.envs;
    if (N > 4096) return PPO_ERR_UNSUPPORTED;
    // A strided value column (one DRAM sector per element) is a latency-bound gather; the register-cached kernel
    // keeps more of those loads in flight and measures faster (30 vs 37 us at 256 x 4096).
    if (p.v_es != 1) return PPO_ERR_UNSUPPORTED;
    {
        const char* sw = getenv("PPO_GAE_TILE");  // development switch: 0 = always the register-cached kernel
        if (sw && atoi(sw) == 0) return PPO_ERR_UNSUPPORTED;
    }
    // few envs: 16 per CTA so that the grid still covers the SMs
    int ew = (A + 31) / 32 < 2 * static_cast<int64_t>(sm_count()) ? 16 : 32;
    {
        const char* sw = getenv("PPO_GAE_EW");  // development switch
        if (sw && (atoi(sw) == 16 || atoi(sw) == 32)) ew = atoi(sw);
    }
    const int db = done_kind == PPO_DONE_U8 ? 1 : done_kind == PPO_DONE_F32 ? 4 : 0;
    const size_t smem = static_cast<size_t>(N) * ew * 4 + static_cast<size_t>(N + 1) * ew * 4 +
                        (static_cast<size_t>(N) * ew * db + 15) / 16 * 16 +
                        static_cast<size_t>(4) * kTileChunks * ew * sizeof(double) +
                        static_cast<size_t>(ew * kTileChunks / 16) * PPO_STATS_LEN * sizeof(double);
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return PPO_ERR_UNSUPPORTED;
    if (smem > static_cast<size_t>(max_smem) / 2) return PPO_ERR_UNSUPPORTED;  // keep at least two CTAs per SM
    const int64_t grid = (A + ew - 1) / ew;
    if (want_stats) {
        if (split_workspace(workspace, workspace_bytes, grid, &p.io) != PPO_OK) return PPO_ERR_UNSUPPORTED;
    }
    GaeTileParams q{};
    q.g = p;
    q.chunk_len = static_cast<int>((N + kTileChunks - 1) / kTileChunks);
    if (ew == 16) q.chunk_len |= 1;
    auto aligned16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    q.vec_r = (aligned16(p.rewards) && p.r_rs % 4 == 0 && A % 4 == 0) ? 16 : 4;
    q.vec_v = (p.v_es == 1 && aligned16(p.values) && p.v_rs % 4 == 0 && A % 4 == 0) ? 16 : 4;
    q.vec_d = 1;
    if (db) {
        const int64_t row_bytes = p.d_rs * db, a_bytes = A * db;
        if (aligned16(p.dones) && row_bytes % 16 == 0 && a_bytes % 16 == 0) q.vec_d = 16;
        else if ((reinterpret_cast<uintptr_t>(p.dones) & 3) == 0 && row_bytes % 4 == 0 && a_bytes % 4 == 0) q.vec_d = 4;
        else if (db == 4) q.vec_d = 4;
    }
    q.same_lambda = p.gl_adv == p.gl_ret ? 1 : 0;
    if (norm_out) {
        if (!want_stats) return PPO_ERR_UNSUPPORTED;
        q.norm_out = norm_out, q.norm_eps = norm_eps, q.norm_clip = norm_clip;
        q.ready = static_cast<unsigned int*>(workspace) + 32;  // bytes 128..135 of the reserved 256-byte header
        if (px) q.px = *px;
    }
#define PPO_TILE(D)                                                                    \
    (ew == 16 ? launch_gae_tile_variant<D, 16>(q, want_stats, grid, smem, s)           \
              : launch_gae_tile_variant<D, 32>(q, want_stats, grid, smem, s))
    if (done_kind == PPO_DONE_U8) return PPO_TILE(PPO_DONE_U8);
    if (done_kind == PPO_DONE_F32) return PPO_TILE(PPO_DONE_F32);
    return PPO_TILE(PPO_DONE_NONE);
#undef PPO_TILE
}

}  // namespace ppo

using namespace ppo;

extern "C" int64_t ppo_gae_workspace_bytes(int64_t envs) {
    int64_t blocks = scan_grid(envs > 0 ? envs : 0);
    if (blocks < kMomentsMaxBlocks) blocks = kMomentsMaxBlocks;
    return 256 + blocks * PPO_STATS_LEN * static_cast<int64_t>(sizeof(double));
}

extern "C" int ppo_gae(const float* rewards, int64_t reward_row_stride, const float* values,
                       int64_t value_row_stride, int64_t value_env_stride, const float* final_value,
                       int64_t final_env_stride, const void* dones, int done_kind, int64_t done_row_stride,
                       int64_t steps, int64_t envs, double gamma, double lamb_adv, double lamb_ret,
                       float* adv_out, float* ret_out, double* stats, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_gae: negative shape [%lld, %lld]", (long long)steps, (long long)envs);
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32,
                "ppo_gae: unknown done_kind %d", done_kind);
    if (steps == 0 || envs == 0) {
        if (stats) PPO_CUDA(cudaMemsetAsync(stats, 0, PPO_STATS_LEN * sizeof(double), (cudaStream_t)stream));
        return PPO_OK;
    }
    PPO_REQUIRE(rewards && values && final_value, "ppo_gae: null input pointer");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || dones != nullptr, "ppo_gae: dones is null but done_kind=%d", done_kind);
    PPO_REQUIRE(envs <= (int64_t)INT32_MAX * 32, "ppo_gae: too many envs");

    GaeParams p{};
    p.rewards = rewards, p.r_rs = reward_row_stride;
    p.values = values, p.v_rs = value_row_stride, p.v_es = value_env_stride;
    p.final_value = final_value, p.f_es = final_env_stride;
    p.dones = dones, p.d_rs = done_row_stride;
    p.steps = steps, p.envs = envs;
    p.gl_adv = gamma * lamb_adv, p.gl_ret = gamma * lamb_ret;
    p.gamma32 = static_cast<float>(gamma);
    p.adv = adv_out, p.ret = ret_out;
    p.chunk_len = static_cast<int>((steps + kScanWarps - 1) / kScanWarps);
    const int64_t blocks = scan_grid(envs);
    const bool want_stats = stats != nullptr;
    if (want_stats) {
        int rc = split_workspace(workspace, workspace_bytes, blocks, &p.io);
        if (rc != PPO_OK) return rc;
        p.io.stats = stats;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (launch_gae_tile(p, done_kind, want_stats, workspace, workspace_bytes, s) == PPO_OK) return PPO_OK;
    const bool cached = p.chunk_len <= kCachedLen;
    dim3 grid(static_cast<unsigned>(blocks)), block(kScanThreads);

#define PPO_LAUNCH_GAE(D)                                                           \
    do {                                                                            \
        if (cached && want_stats) gae_kernel<D, true, true><<<grid, block, 0, s>>>(p);       \
        else if (cached) gae_kernel<D, true, false><<<grid, block, 0, s>>>(p);       \
        else if (want_stats) gae_kernel<D, false, true><<<grid, block, 0, s>>>(p);   \
        else gae_kernel<D, false, false><<<grid, block, 0, s>>>(p);                  \
    } while (0)
    if (done_kind == PPO_DONE_U8) PPO_LAUNCH_GAE(PPO_DONE_U8);
    else if (done_kind == PPO_DONE_F32) PPO_LAUNCH_GAE(PPO_DONE_F32);
    else PPO_LAUNCH_GAE(PPO_DONE_NONE);
#undef PPO_LAUNCH_GAE
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_gae_normalized(const float* rewards, int64_t reward_row_stride, const float* values,
                                  int64_t value_row_stride, int64_t value_env_stride, const float* final_value,
                                  int64_t final_env_stride, const void* dones, int done_kind,
                                  int64_t done_row_stride, int64_t steps, int64_t envs, double gamma, double lamb_adv,
                                  double lamb_ret, double eps, double clip, float* adv_out, float* ret_out,
                                  float* norm_out, double* stats, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    PPO_REQUIRE(stats && norm_out && adv_out, "ppo_gae_normalized: stats, adv_out and norm_out are required");
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_gae_normalized: negative shape");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32,
                "ppo_gae_normalized: unknown done_kind %d", done_kind);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (steps > 0 && envs > 0 && rewards && values && final_value && (done_kind == PPO_DONE_NONE || dones)) {
        GaeParams p{};
        p.rewards = rewards, p.r_rs = reward_row_stride;
        p.values = values, p.v_rs = value_row_stride, p.v_es = value_env_stride;
        p.final_value = final_value, p.f_es = final_env_stride;
        p.dones = dones, p.d_rs = done_row_stride;
        p.steps = steps, p.envs = envs;
        p.gl_adv = gamma * lamb_adv, p.gl_ret = gamma * lamb_ret;
        p.gamma32 = static_cast<float>(gamma);
        p.adv = adv_out, p.ret = ret_out;
        p.io.stats = stats;
        // one cooperative launch when the whole grid is resident; PPO_ERR_UNSUPPORTED otherwise
        if (launch_gae_tile(p, done_kind, true, workspace, workspace_bytes, s, norm_out, static_cast<float>(eps),
                            clip > 0 ? static_cast<float>(clip) : 0.f) == PPO_OK)
            return 1;  // fused: one launch
    }
    int rc = ppo_gae(rewards, reward_row_stride, values, value_row_stride, value_env_stride, final_value,
                     final_env_stride, dones, done_kind, done_row_stride, steps, envs, gamma, lamb_adv, lamb_ret,
                     adv_out, ret_out, stats, workspace, workspace_bytes, stream);
    if (rc != PPO_OK) return rc;
    return ppo_adv_normalize(adv_out, steps * envs, stats, eps, clip, norm_out, stream);
}

extern "C" int ppo_gae_normalized_fits(int64_t steps, int64_t envs, int done_kind, int64_t value_env_stride) {
    if (steps <= 0 || envs <= 0) return 0;
    GaeParams p{};
    alignas(16) static float dummy[4];
    p.rewards = p.values = p.final_value = dummy, p.dones = dummy;  // never dereferenced: nothing is launched
    p.r_rs = p.v_rs = p.d_rs = envs, p.v_es = value_env_stride, p.f_es = 1;
    p.steps = steps, p.envs = envs, p.gl_adv = p.gl_ret = 0.5, p.gamma32 = 0.5f;
    p.adv = p.ret = dummy;
    double stats_dummy[PPO_STATS_LEN];
    p.io.stats = stats_dummy;
    t_tile_dry_run = true;
    const int rc = launch_gae_tile(p, done_kind, true, dummy, ppo_gae_workspace_bytes(envs), nullptr, dummy, 1e-8f, 0.f);
    t_tile_dry_run = false;
    return rc == PPO_OK ? 1 : 0;
}

extern "C" int ppo_gae_normalized_peers(const float* rewards, int64_t reward_row_stride, const float* values,
                                        int64_t value_row_stride, int64_t value_env_stride, const float* final_value,
                                        int64_t final_env_stride, const void* dones, int done_kind,
                                        int64_t done_row_stride, int64_t steps, int64_t envs, double gamma,
                                        double lamb_adv, double lamb_ret, double eps, double clip, float* adv_out,
                                        float* ret_out, float* norm_out, double* stats, double* global_stats,
                                        void* workspace, int64_t workspace_bytes, int rank, int world,
                                        double* const* peer_mail, unsigned long long* const* peer_flags,
                                        unsigned long long epoch, unsigned long long* status, double timeout_ms,
                                        void* stream) {
    PPO_REQUIRE(stats && norm_out && adv_out, "ppo_gae_normalized_peers: stats, adv_out and norm_out are required");
    PPO_REQUIRE(world >= 2 && world <= PPO_MAX_PEERS && rank >= 0 && rank < world,
                "ppo_gae_normalized_peers: rank %d of %d (at most %d peers)", rank, world, PPO_MAX_PEERS);
    PPO_REQUIRE(peer_mail && peer_flags && epoch > 0, "ppo_gae_normalized_peers: mailboxes and a positive epoch are required");
    PPO_REQUIRE(steps > 0 && envs > 0 && rewards && values && final_value, "ppo_gae_normalized_peers: empty shard");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || ((done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32) && dones),
                "ppo_gae_normalized_peers: bad done_kind %d", done_kind);
    GaeParams p{};
    p.rewards = rewards, p.r_rs = reward_row_stride;
    p.values = values, p.v_rs = value_row_stride, p.v_es = value_env_stride;
    p.final_value = final_value, p.f_es = final_env_stride;
    p.dones = dones, p.d_rs = done_row_stride;
    p.steps = steps, p.envs = envs;
    p.gl_adv = gamma * lamb_adv, p.gl_ret = gamma * lamb_ret;
    p.gamma32 = static_cast<float>(gamma);
    p.adv = adv_out, p.ret = ret_out;
    p.io.stats = stats;
    PeerExchange px{};
    px.world = world, px.rank = rank, px.epoch = epoch, px.global_stats = global_stats;
    px.status = status;
    px.timeout_ns = static_cast<unsigned long long>((timeout_ms > 0 ? timeout_ms : 4000.0) * 1e6);
    for (int r = 0; r < world; ++r) {
        PPO_REQUIRE(peer_mail[r] && peer_flags[r], "ppo_gae_normalized_peers: null mailbox of rank %d", r);
        px.mail[r] = peer_mail[r], px.flags[r] = peer_flags[r];
    }
    // Every rank must take the same decision (a rank that does not launch never raises its flags): the fused launch
    // needs the tile kernel with its whole grid resident, which depends on the shard shape only.
    const int rc = launch_gae_tile(p, done_kind, true, workspace, workspace_bytes, static_cast<cudaStream_t>(stream),
                                   norm_out, static_cast<float>(eps), clip > 0 ? static_cast<float>(clip) : 0.f, &px);
    return rc == PPO_OK ? 1 : 0;  // 0: not launched (shape does not fit), use ppo_gae + all-reduce + ppo_adv_normalize
}

extern "C" int ppo_discounted_returns(const float* rewards, int64_t reward_row_stride, const void* dones,
                                      int done_kind, int64_t done_row_stride, const float* final_value,
                                      int64_t final_env_stride, double gamma, const float* gamma_array,
                                      int64_t gamma_row_stride, int64_t steps, int64_t envs, float* out,
                                      void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_discounted_returns: negative shape");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32,
                "ppo_discounted_returns: unknown done_kind %d", done_kind);
    if (steps == 0 || envs == 0) return PPO_OK;
    PPO_REQUIRE(rewards && final_value && out, "ppo_discounted_returns: null pointer");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || dones != nullptr, "ppo_discounted_returns: dones is null");
    ReturnsParams p{};
    p.rewards = rewards, p.r_rs = reward_row_stride;
    p.dones = dones, p.d_rs = done_row_stride;
    p.final_value = final_value, p.f_es = final_env_stride;
    p.gamma_array = gamma_array, p.g_rs = gamma_row_stride;
    p.gamma32 = static_cast<float>(gamma);
    p.steps = steps, p.envs = envs, p.out = out;
    p.chunk_len = static_cast<int>((steps + kScanWarps - 1) / kScanWarps);
    dim3 grid(static_cast<unsigned>(scan_grid(envs))), block(kScanThreads);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (done_kind == PPO_DONE_U8) returns_kernel<PPO_DONE_U8><<<grid, block, 0, s>>>(p);
    else if (done_kind == PPO_DONE_F32) returns_kernel<PPO_DONE_F32><<<grid, block, 0, s>>>(p);
    else returns_kernel<PPO_DONE_NONE><<<grid, block, 0, s>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

namespace ppo {
// Records of disjoint env blocks (or ranks) -> one record: sums add in record order, min / max reduce.
__global__ void merge_stats_kernel(const double* __restrict__ records, int64_t n, double* __restrict__ out) {
    const int j = threadIdx.x;
    if (j >= PPO_STATS_LEN) return;
    const bool is_min = (j == PPO_STAT_ADV_MIN || j == PPO_STAT_RET_MIN);
    const bool is_max = (j == PPO_STAT_ADV_MAX || j == PPO_STAT_RET_MAX);
    double acc = is_min ? INFINITY : is_max ? -INFINITY : 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double x = records[i * PPO_STATS_LEN + j];
        acc = is_min ? fmin(acc, x) : is_max ? fmax(acc, x) : acc + x;
    }
    out[j] = j <= PPO_STAT_RET_ZEROS ? acc : 0.0;
}
}  // namespace ppo

extern "C" int ppo_merge_stats(const double* records, int64_t n_records, double* out, void* stream) {
    PPO_REQUIRE(n_records >= 1 && records && out, "ppo_merge_stats: bad arguments");
    merge_stats_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(records, n_records, out);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_moments(const float* x, int64_t count, double* stats, void* workspace,
                           int64_t workspace_bytes, void* stream) {
    PPO_REQUIRE(count >= 0 && stats != nullptr, "ppo_moments: bad arguments");
    PPO_REQUIRE(count == 0 || x != nullptr, "ppo_moments: null input");
    int64_t blocks = (count + 512 * 8 - 1) / (512 * 8);
    if (blocks < 1) blocks = 1;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 2;
    if (blocks > cap) blocks = cap;
    if (blocks > kMomentsMaxBlocks) blocks = kMomentsMaxBlocks;
    MomentsParams p{};
    p.x = x, p.count = count;
    int rc = split_workspace(workspace, workspace_bytes, blocks, &p.io);
    if (rc != PPO_OK) return rc;
    p.io.stats = stats;
    moments_kernel<<<static_cast<unsigned>(blocks), 512, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_adv_combine(const float* ext_adv, const float* int_adv, double scale, int64_t count, float* out,
                               double* stats, void* workspace, int64_t workspace_bytes, void* stream) {
    PPO_REQUIRE(count >= 0 && stats != nullptr, "ppo_adv_combine: bad arguments");
    PPO_REQUIRE(count == 0 || (ext_adv && int_adv && out), "ppo_adv_combine: null pointer");
    int64_t blocks = (count + 512 * 8 - 1) / (512 * 8);
    if (blocks < 1) blocks = 1;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 2;
    if (blocks > cap) blocks = cap;
    if (blocks > kMomentsMaxBlocks) blocks = kMomentsMaxBlocks;
    CombineParams p{};
    p.ext = ext_adv, p.intr = int_adv, p.scale = static_cast<float>(scale), p.out = out, p.count = count;
    int rc = split_workspace(workspace, workspace_bytes, blocks, &p.io);
    if (rc != PPO_OK) return rc;
    p.io.stats = stats;
    adv_combine_kernel<<<static_cast<unsigned>(blocks), 512, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_adv_normalize(const float* adv, int64_t count, const double* stats, double eps, double clip,
                                 float* out, void* stream) {
    PPO_REQUIRE(count >= 0, "ppo_adv_normalize: negative count");
    if (count == 0) return PPO_OK;
    PPO_REQUIRE(adv && out && stats, "ppo_adv_normalize: null pointer");
    NormParams p{};
    p.x = adv, p.out = out, p.count = count, p.stats = stats;
    p.eps = static_cast<float>(eps);
    p.clip = clip > 0 ? static_cast<float>(clip) : 0.f;
    int64_t blocks = (count / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    normalize_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
