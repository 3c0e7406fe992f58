// TVF multi-horizon truncated-return targets: entry point and the generic (any-shape) kernel.
//
// Replaces rl/returns_truncated.py:623-693 (+ :558-620, :142-174) of the reference.
//   out[t,a,k] = 1/C * sum_c ( S_n(t,a) + D_n(t,a) * I(V[t+n,a,:], h_k - n) ),  n = min(sample[k,c], h_k, N - t)
//
// Decomposition used by both kernels (DESIGN.md "TVF targets"):
//  * D_n(t,a) = G[n] when no terminal occurs in steps t..t+n-1, else 0, where G[n] is the reference's
//    float32 running product of gamma -- identical for every (t, a), so it is a host table and the
//    only per-(t,a) state is nd(t,a), the distance to the first terminal at or after t;
//  * S_n(t,a) is a prefix sum over i < n of float32(r[t+i] * D_i(t,a)) accumulated in float32 in
//    forward order exactly like rl/returns_truncated.py:668; one row P[t][0..nmax] per (t,a) is built
//    once and shared by all K horizons;
//  * per (t,a,k) the (up to C) distinct clamped n-steps n' = min(sample, h_k) come from the host
//    entry table with their slots / weights; entries with n' > min(N - t, nd) collapse into a single
//    shared term (the episode ended: no bootstrap; or the rollout ended: bootstrap from row N).
#include <stdlib.h>
#include "tvf_common.cuh"

namespace ppo {

int launch_tvf_tslab(const TvfParams& p, double* stats, double* partials, unsigned int* ticket, int64_t max_partials,
                     const float* host_value_weights, float* value_out, int64_t value_row_stride, cudaStream_t stream,
                     bool* stats_fused, const TsGae* gae, bool* gae_fused);  // tvf_tslab.cu
int launch_tvf_qslab(const TvfParams& p, double* stats, double* partials, unsigned int* ticket, int64_t max_partials,
                     const float* host_value_weights, float* value_out, int64_t value_row_stride, cudaStream_t stream,
                     bool* stats_fused);  // tvf_qslab.cu
int launch_tvf_cslab(const TvfParams& p, double* stats, double* partials, unsigned int* ticket, int64_t max_partials,
                     const float* host_value_weights, float* value_out, int64_t value_row_stride, cudaStream_t stream,
                     bool* stats_fused);
int launch_value_dot(const float* v, int64_t rows_t, int64_t envs, int64_t rs_t, int64_t rs_e, int k,
                     const float* host_weights, float* out, int64_t os_t, cudaStream_t s);  // value_estimate.cu
int launch_tvf_ring(const TvfParams& p, const float* tail_cwlo, const float* tail_cwhi, cudaStream_t stream,
                    bool probe_only);  // tvf_ring.cu

// ------------------------------------------------------------------------------------------
// Generic kernel: one CTA per (env, chunk of TC time rows); value rows are gathered from global
// memory (L1/L2 cached).  Works for any N, A, K, V, C; used when the ring kernel's shape
// constraints do not hold.
constexpr int kGenThreads = 256;

__global__ void __launch_bounds__(kGenThreads) tvf_generic_kernel(const TvfParams p, int tc, int n_chunks) {
    extern __shared__ float smem[];
    const int nmax = p.tab.nmax;
    const int pstride = nmax + 1;
    float* P = smem;                               // [tc][nmax + 1]
    float* rs = P + static_cast<size_t>(tc) * pstride;  // [tc + nmax]
    int* nd = reinterpret_cast<int*>(rs + tc + nmax);   // [tc]
    unsigned char* ds = reinterpret_cast<unsigned char*>(nd + tc);  // [tc + nmax]

    const int64_t env = blockIdx.x / n_chunks;
    const int t0 = static_cast<int>(blockIdx.x % n_chunks) * tc;
    const int N = p.steps;
    const int rows = min(tc, N - t0);
    const int span = min(rows + nmax, N - t0);  // reward rows t0 .. t0 + span - 1

    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        rs[i] = __ldg(p.rewards + static_cast<int64_t>(t0 + i) * p.r_rs + env);
        ds[i] = done_at(p.dones, p.done_kind, static_cast<int64_t>(t0 + i) * p.d_rs + env) ? 1 : 0;
    }
    __syncthreads();

    // prefix rows, one thread per time row (sequential float32 accumulation, reference order)
    if (threadIdx.x < rows) {
        const int tl = threadIdx.x;
        const int lim = min(nmax, N - (t0 + tl));
        float* row = P + static_cast<size_t>(tl) * pstride;
        float s = 0.f;
        bool alive = true;
        int first_done = kNoDone;
        row[0] = 0.f;
        for (int i = 0; i < lim; ++i) {
            const float disc = alive ? __ldg(p.tab.gpow + i) : 0.f;
            s = __fadd_rn(s, __fmul_rn(rs[tl + i], disc));
            row[i + 1] = s;
            if (alive && ds[tl + i]) {
                alive = false;
                first_done = i;
            }
        }
        nd[tl] = first_done;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int K = p.n_req;
    const int E = p.tab.n_entries;
    for (int tl = warp; tl < rows; tl += nwarps) {
        const int t = t0 + tl;
        const int m = N - t;                 // steps left in the rollout
        const int ndt = nd[tl];
        const int lim = min(m, ndt);         // largest n that is alive and inside the rollout
        const float* row = P + static_cast<size_t>(tl) * pstride;
        const float* vbase = p.vs + static_cast<int64_t>(t) * p.vs_rs + env * p.vs_es;
        for (int k = lane; k < K; k += 32) {
            float acc = 0.f, rest = 0.f;
            for (int e = 0; e < E; ++e) {
                const float cnt = __ldg(p.tab.ent_cnt + e * K + k);
                if (cnt == 0.f) break;  // entries are packed to the front
                const int n = __ldg(p.tab.ent_n + e * K + k);
                if (n <= lim) {
                    const int idx = __ldg(p.tab.ent_idx + e * K + k);
                    const float* vr = vbase + static_cast<int64_t>(n) * p.vs_rs;
                    const float boot = lerp2(__ldg(vr + (idx & 0xffff)), __ldg(vr + (idx >> 16)),
                                             __ldg(p.tab.ent_wlo + e * K + k), __ldg(p.tab.ent_whi + e * K + k));
                    const float term = __fadd_rn(row[n], __fmul_rn(boot, __ldg(p.tab.ent_g + e * K + k)));
                    acc = fmaf(cnt, term, acc);
                } else {
                    rest += cnt;
                }
            }
            if (rest > 0.f) {
                float term;
                if (ndt < m) {
                    term = row[ndt + 1];  // episode ended inside the window: D = 0, S stops after the terminal step
                } else {
                    // rollout ended first: n = N - t, bootstrap from the final row (rl/returns_truncated.py:610-614)
                    const int idx = __ldg(p.tab.tail_idx + m * K + k);
                    const float* vr = p.vs + static_cast<int64_t>(N) * p.vs_rs + env * p.vs_es;
                    const float boot = lerp2(__ldg(vr + (idx & 0xffff)), __ldg(vr + (idx >> 16)),
                                             __ldg(p.tab.tail_wlo + m * K + k), __ldg(p.tab.tail_whi + m * K + k));
                    term = __fadd_rn(row[m], __fmul_rn(boot, __ldg(p.tab.gpow + m)));
                }
                acc = fmaf(rest, term, acc);
            }
            p.out[(static_cast<int64_t>(t) * p.envs + env) * K + k] = acc * p.tab.inv_c;
        }
    }
}

static int launch_generic(const TvfParams& p, cudaStream_t stream) {
    const int nmax = p.tab.nmax;
    int max_smem = 0, dev = 0;
    PPO_CUDA(cudaGetDevice(&dev));
    PPO_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    auto bytes_for = [&](int tc) {
        return static_cast<size_t>(tc) * (nmax + 1) * 4 + static_cast<size_t>(tc + nmax) * 5 + tc * 4 + 16;
    };
    int tc = 32;
    const size_t budget = static_cast<size_t>(max_smem) < (96u << 10) ? max_smem : (96u << 10);
    while (tc > 1 && bytes_for(tc) > budget) tc >>= 1;
    if (bytes_for(tc) > static_cast<size_t>(max_smem)) {
        set_error("ppo_tvf_targets: n-step %d needs %zu bytes of shared memory per CTA", nmax, bytes_for(tc));
        return PPO_ERR_UNSUPPORTED;
    }
    if (tc > p.steps) tc = p.steps;
    const int n_chunks = (p.steps + tc - 1) / tc;
    const int64_t blocks = p.envs * n_chunks;
    if (blocks > INT32_MAX) {
        set_error("ppo_tvf_targets: grid of %lld CTAs is too large", (long long)blocks);
        return PPO_ERR_UNSUPPORTED;
    }
    const size_t smem = bytes_for(tc);
    PPO_CUDA(cudaFuncSetAttribute(tvf_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    tvf_generic_kernel<<<static_cast<unsigned>(blocks), kGenThreads, smem, stream>>>(p, tc, n_chunks);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

// last-horizon statistics (rl/rollout.py:1270): strided read of out[:, :, K-1]
struct LastHorizonParams {
    const float* out;
    int64_t rows;  // N * A
    int64_t K;
    double* stats;
    double* partials;
    unsigned int* ticket;
};

__global__ void __launch_bounds__(512) tvf_last_horizon_stats_kernel(const LastHorizonParams p) {
    __shared__ double red[16][4];  // n, sum, sumsq, zeros
    __shared__ bool is_last;
    double n = 0, s = 0, q = 0, z = 0;
    float mn = INFINITY, mx = -INFINITY;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.rows; i += stride) {
        const float x = __ldg(p.out + i * p.K + (p.K - 1));
        n += 1.0, s += x, q += static_cast<double>(x) * x, z += x == 0.f ? 1.0 : 0.0;
        mn = fminf(mn, x), mx = fmaxf(mx, x);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    n = warp_sum_f64(n), s = warp_sum_f64(s), q = warp_sum_f64(q), z = warp_sum_f64(z);
    mn = warp_min_f32(mn), mx = warp_max_f32(mx);
    __shared__ float redm[16][2];
    if (lane == 0) red[w][0] = n, red[w][1] = s, red[w][2] = q, red[w][3] = z, redm[w][0] = mn, redm[w][1] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 16; ++i) {
            n += red[i][0], s += red[i][1], q += red[i][2], z += red[i][3];
            mn = fminf(mn, redm[i][0]), mx = fmaxf(mx, redm[i][1]);
        }
        double* rec = p.partials + static_cast<int64_t>(blockIdx.x) * 8;
        rec[0] = n, rec[1] = s, rec[2] = q, rec[3] = mn, rec[4] = mx, rec[5] = z;
        __threadfence();
        is_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (w == 0) {
        double an = 0, as = 0, aq = 0, az = 0, amn = INFINITY, amx = -INFINITY;
        for (unsigned int b = lane; b < gridDim.x; b += 32) {
            const double* rec = p.partials + static_cast<int64_t>(b) * 8;
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
            for (int j = 0; j < PPO_STATS_LEN; ++j) p.stats[j] = 0.0;
            p.stats[PPO_STAT_COUNT] = an;
            p.stats[PPO_STAT_ADV_SUM] = as;
            p.stats[PPO_STAT_ADV_SUMSQ] = aq;
            p.stats[PPO_STAT_ADV_MIN] = amn;
            p.stats[PPO_STAT_ADV_MAX] = amx;
            p.stats[PPO_STAT_ADV_ZEROS] = az;
            *p.ticket = 0u;
        }
    }
}

constexpr int kStatsBlocks = 296;

}  // namespace ppo

using namespace ppo;

static thread_local int g_last_path = 0;

extern "C" int ppo_tvf_last_path(void) { return g_last_path; }

// PPO_TVF_AUTO takes the quad slab kernel at 256 steps and up to 4096 envs per launch (x 128 horizons on 128 value
// horizons: the launcher checks the rest) -- 0.488 against 0.510 ms per bench pass, bit-identical targets; 125 / 228 /
// 448 us against 127 / 238 / 453 us for the targets of 1024 / 2048 / 4096 envs.  Beyond that it falls off a cliff
// (4608 envs: 553 against 513 us, 8192: 1195 against 891 us, 16384: 2398 against 1757 us): its 128-row x 32-byte stores
// and 128-row x 128-byte loads visit a page four times as often per byte as the transposed slab kernel's 32-row boxes,
// and past ~1.1 GB of buffers the address translation no longer keeps up.  At 128 steps x 1024 envs the transposed slab kernel is the faster
// one as well.  PPO_TVF_QSLAB=0 / 1 in the environment forces the choice (read per call: A/B runs).
static bool prefer_qslab(int64_t steps, int64_t envs) {
    const char* e = getenv("PPO_TVF_QSLAB");
    if (e != nullptr) return atoi(e) != 0;
    return steps == 256 && envs <= 4096;
}

extern "C" int64_t ppo_tvf_workspace_bytes(int64_t steps, int64_t envs) {
    (void)steps;
    // ticket + per-block records of the statistics kernels, or one record per env (column slab kernel)
    const int64_t records = envs > kStatsBlocks ? envs : kStatsBlocks;
    return 256 + records * 8 * static_cast<int64_t>(sizeof(double));
}

extern "C" int ppo_tvf_targets(const float* rewards, int64_t reward_row_stride, const void* dones, int done_kind,
                               int64_t done_row_stride, const float* value_samples, int64_t vs_row_stride,
                               int64_t vs_env_stride, int64_t steps, int64_t envs, int64_t n_req, int64_t n_val,
                               const ppo_tvf_tables* tables, int path, float* out, double* stats,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    return ppo_tvf_targets_ex(rewards, reward_row_stride, dones, done_kind, done_row_stride, value_samples,
                              vs_row_stride, vs_env_stride, steps, envs, n_req, n_val, tables, path, out, stats,
                              workspace, workspace_bytes, nullptr, stream);
}

extern "C" int ppo_tvf_targets_ex(const float* rewards, int64_t reward_row_stride, const void* dones, int done_kind,
                                  int64_t done_row_stride, const float* value_samples, int64_t vs_row_stride,
                                  int64_t vs_env_stride, int64_t steps, int64_t envs, int64_t n_req, int64_t n_val,
                                  const ppo_tvf_tables* tables, int path, float* out, double* stats,
                                  void* workspace, int64_t workspace_bytes, ppo_tvf_extras* extras, void* stream) {
    if (extras) extras->gae_done = 0;
    const float* vweights = extras ? extras->value_weights : nullptr;
    float* vout = extras ? extras->value_out : nullptr;
    const int64_t vout_rs = extras && extras->value_out_row_stride > 0 ? extras->value_out_row_stride : envs;
    PPO_REQUIRE(steps >= 0 && envs >= 0 && n_req >= 0 && n_val >= 1, "ppo_tvf_targets: bad shape");
    PPO_REQUIRE(steps < (1 << 24) && n_req < (1 << 20) && n_val < (1 << 16), "ppo_tvf_targets: shape too large");
    PPO_REQUIRE(done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32 || done_kind == PPO_DONE_NONE,
                "ppo_tvf_targets: unknown done_kind %d", done_kind);
    PPO_REQUIRE(path == PPO_TVF_AUTO || path == PPO_TVF_GENERIC || path == PPO_TVF_RING || path == PPO_TVF_TSLAB ||
                    path == PPO_TVF_CSLAB || path == PPO_TVF_QSLAB,
                "ppo_tvf_targets: unknown path %d", path);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (steps == 0 || envs == 0 || n_req == 0) {
        if (stats) PPO_CUDA(cudaMemsetAsync(stats, 0, PPO_STATS_LEN * sizeof(double), s));
        return PPO_OK;
    }
    PPO_REQUIRE(rewards && value_samples && out && tables, "ppo_tvf_targets: null pointer");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || dones, "ppo_tvf_targets: dones is null");
    PPO_REQUIRE(tables->n_entries >= 1 && tables->nmax >= 1 && tables->nmax <= steps,
                "ppo_tvf_targets: tables describe n-steps up to %d for %lld steps", tables->nmax, (long long)steps);
    // the generic section is all there or all absent (absent: tables built for the slab kernel only)
    const bool generic_tables = tables->ent_n && tables->ent_cnt && tables->ent_idx && tables->ent_wlo && tables->ent_whi &&
                                tables->ent_g && tables->tail_idx && tables->tail_wlo && tables->tail_whi && tables->gpow;
    PPO_REQUIRE(generic_tables || (!tables->ent_n && tables->host_blob && tables->dev_blob && tables->tslab_offset_words > 0),
                "ppo_tvf_targets: null table pointer");

    TvfParams p{};
    p.rewards = rewards, p.r_rs = reward_row_stride;
    p.dones = dones, p.d_rs = done_row_stride, p.done_kind = done_kind;
    p.vs = value_samples, p.vs_rs = vs_row_stride, p.vs_es = vs_env_stride;
    p.steps = static_cast<int>(steps), p.envs = envs;
    p.n_req = static_cast<int>(n_req), p.n_val = static_cast<int>(n_val);
    p.tab = *tables;
    p.out = out;

    int rc = PPO_ERR_UNSUPPORTED;
    bool stats_fused = false;
    int taken = PPO_TVF_GENERIC;
    // Several n-steps per horizon: gather-bound, column slab kernel (128 / 256 steps) or transposed slab kernel.
    // One or two: HBM-bound, ring kernel.
    double* partials = nullptr;
    unsigned int* ticket = nullptr;
    if (stats && workspace && workspace_bytes >= ppo_tvf_workspace_bytes(0, 0)) {
        ticket = static_cast<unsigned int*>(workspace);
        partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    }
    static const bool prefer_cslab = [] {
        const char* e = getenv("PPO_TVF_PREFER_CSLAB");
        return e != nullptr && atoi(e) != 0;
    }();
    if (path == PPO_TVF_CSLAB || (path == PPO_TVF_AUTO && tables->n_entries >= 3 && prefer_cslab)) {
        rc = launch_tvf_cslab(p, stats, partials, ticket, partials ? (workspace_bytes - 256) / 8 : 0, vweights, vout, vout_rs, s,
                              &stats_fused);
        if (path == PPO_TVF_CSLAB && rc != PPO_OK) return rc;
        taken = PPO_TVF_CSLAB;
        if (rc == PPO_OK) vout = nullptr;  // the value estimate was written by the same launch
    }
    // (a caller that asks for the fused GAE stage gets the transposed slab kernel, the only one that has it)
    const bool wants_gae = extras && extras->gae_enable;
    if (path == PPO_TVF_QSLAB ||
        (path == PPO_TVF_AUTO && tables->n_entries >= 3 && rc == PPO_ERR_UNSUPPORTED && !wants_gae && prefer_qslab(steps, envs))) {
        rc = launch_tvf_qslab(p, stats, partials, ticket, kStatsBlocks, vweights, vout, vout_rs, s, &stats_fused);
        if (path == PPO_TVF_QSLAB && rc != PPO_OK) return rc;
        if (rc != PPO_OK && rc != PPO_ERR_UNSUPPORTED) return rc;
        if (rc == PPO_OK) {
            taken = PPO_TVF_QSLAB;
            vout = nullptr;  // the value estimate was written by the same launch
        }
    }
    if (path == PPO_TVF_TSLAB || (path == PPO_TVF_AUTO && tables->n_entries >= 3 && rc == PPO_ERR_UNSUPPORTED)) {
        TsGae gae{};
        const bool ask_gae = extras && extras->gae_enable && extras->gae_adv_out && extras->gae_ret_out && extras->gae_stats &&
                             extras->gae_workspace && extras->gae_workspace_bytes >= 256 + kStatsBlocks * PPO_STATS_LEN * 8;
        if (ask_gae) {
            gae.gamma = extras->gae_gamma, gae.lambda_adv = extras->gae_lambda_adv, gae.lambda_ret = extras->gae_lambda_ret;
            gae.adv = extras->gae_adv_out, gae.ret = extras->gae_ret_out, gae.stats = extras->gae_stats;
            gae.ticket = static_cast<unsigned int*>(extras->gae_workspace);
            gae.partials = reinterpret_cast<double*>(static_cast<char*>(extras->gae_workspace) + 256);
            gae.max_partials = kStatsBlocks;
        }
        bool gae_fused = false;
        rc = launch_tvf_tslab(p, stats, partials, ticket, kStatsBlocks, vweights, vout, vout_rs, s, &stats_fused,
                              ask_gae ? &gae : nullptr, &gae_fused);
        if (rc == PPO_OK && gae_fused) extras->gae_done = 1;
        if (path == PPO_TVF_TSLAB && rc != PPO_OK) return rc;
        taken = PPO_TVF_TSLAB;
        if (rc == PPO_OK) vout = nullptr;  // the value estimate was written by the same launch
    }
    if (rc == PPO_ERR_UNSUPPORTED && !generic_tables) {
        // (tables built with ppo_tvf_tables_build_ex for the slab kernel only, and the slab kernel declined)
        set_error("ppo_tvf_targets: the tables carry no generic section (ppo_tvf_tables_build_ex sections) and the slab "
                  "kernel cannot take this call; rebuild them with PPO_TVF_TABLES_ALL");
        return PPO_ERR_UNSUPPORTED;
    }
    if (rc == PPO_ERR_UNSUPPORTED && path != PPO_TVF_GENERIC) {
        rc = launch_tvf_ring(p, tables->tail_cwlo, tables->tail_cwhi, s, false);
        if (path == PPO_TVF_RING && rc != PPO_OK) return rc;
        taken = PPO_TVF_RING;
    }
    if (rc == PPO_ERR_UNSUPPORTED) rc = launch_generic(p, s), taken = PPO_TVF_GENERIC;
    if (rc != PPO_OK) return rc;
    if (vout) {  // not fused: one more pass over the value rows (weighted) or a strided column copy (last horizon)
        rc = launch_value_dot(value_samples, steps + 1, envs, vs_row_stride, vs_env_stride, static_cast<int>(n_val),
                              vweights, vout, vout_rs, s);
        if (rc != PPO_OK) return rc;
    }
    g_last_path = taken | (stats_fused ? 0x100 : 0) | (extras && extras->value_out && !vout ? 0x200 : 0);

    if (stats && !stats_fused)
        return ppo_tvf_last_horizon_stats(out, steps * envs, n_req, stats, workspace, workspace_bytes, stream);
    return PPO_OK;
}

extern "C" int ppo_tvf_last_horizon_stats(const float* targets, int64_t rows, int64_t n_req, double* stats,
                                          void* workspace, int64_t workspace_bytes, void* stream) {
    PPO_REQUIRE(rows >= 0 && n_req >= 1 && stats, "ppo_tvf_last_horizon_stats: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (rows == 0) {
        PPO_CUDA(cudaMemsetAsync(stats, 0, PPO_STATS_LEN * sizeof(double), s));
        return PPO_OK;
    }
    PPO_REQUIRE(targets, "ppo_tvf_last_horizon_stats: null targets");
    const int64_t need = ppo_tvf_workspace_bytes(0, 0);
    if (!workspace || workspace_bytes < need) {
        set_error("ppo_tvf_last_horizon_stats: workspace of %lld bytes required", (long long)need);
        return PPO_ERR_WORKSPACE;
    }
    LastHorizonParams q{};
    q.out = targets, q.rows = rows, q.K = n_req, q.stats = stats;
    q.ticket = static_cast<unsigned int*>(workspace);
    q.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    int64_t blocks = (q.rows + 511) / 512;
    if (blocks > kStatsBlocks) blocks = kStatsBlocks;
    tvf_last_horizon_stats_kernel<<<static_cast<unsigned>(blocks), 512, 0, s>>>(q);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
