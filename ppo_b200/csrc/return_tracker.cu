// Forward return trackers and running moments: the reward / intrinsic-return normalisation around the hot path.
//
//   rl/rollout.py:1147-1165  (Runner.calculate_intrinsic_returns)   ems_t = (1 - done_t) * gamma * ems_{t-1} + r_t
//   rl/wrappers.py:847       (VecNormalizeRewardWrapper.step)        ret_t = r_t + gamma * ret_{t-1} * (1 - done_t)
//   rl/utils.py:379-443      (RunningMeanStd / update_mean_var_count_from_moments)
//
// Both trackers are the same linear recurrence run FORWARDS in time with a float64 carry that survives from one
// rollout to the next; after every step the batch mean / variance of the tracker over the envs is merged into a
// running (mean, var, count) with the parallel-variance update.  The reference does this with N NumPy passes over
// [A] vectors; here:
//   * tracker_scan_kernel   one thread per (env, time chunk): the chunk is composed into x_out = B + C * x_in from
//                           registers, the carries are chained through shared memory, the chunk is replayed and the
//                           tracker rows are written as float64 [N, A] (env axis contiguous, coalesced);
//   * row_sum_kernel        one CTA per time row: sum of the row, or the sum of squared deviations from the row mean
//                           (two passes like np.var, the second one re-reads the row from L2);
//   * rms_fold_kernel       the N sequential merges of rl/utils.py:379-396 in float64, one thread, also recording the
//                           running variance after every step (what the reward wrapper divides by);
//   * scale_rewards_kernel  r / scale, optional clip, optional centring.
// Env-sharded runs all-reduce the two [N] row vectors between the launches (host side), nothing else.
#include <math.h>

#include "common.cuh"

namespace ppo {

constexpr int kTrkEnvs = 32;     // envs per CTA (one lane each)
constexpr int kTrkChunks = 8;    // time chunks per CTA
constexpr int kTrkLen = 32;      // steps per chunk and pass (registers)
constexpr int kTrkPass = kTrkChunks * kTrkLen;

struct TrackerParams {
    const float* rewards;
    int64_t r_rs;
    const void* dones;
    int done_kind;
    int64_t d_rs;
    int64_t steps, envs;
    double gamma;
    double* carry;  // [A] in / out
    double* x;      // [N, A]
};

template <int DONE>
__global__ void __launch_bounds__(kTrkEnvs* kTrkChunks) tracker_scan_kernel(const TrackerParams p) {
    __shared__ double sB[kTrkChunks][kTrkEnvs], sC[kTrkChunks][kTrkEnvs];
    const int lane = threadIdx.x % kTrkEnvs, w = threadIdx.x / kTrkEnvs;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * kTrkEnvs + lane;
    const bool live = env < p.envs;
    double x_in = live ? p.carry[env] : 0.0;
    for (int64_t t0 = 0; t0 < p.steps; t0 += kTrkPass) {
        const int64_t lo = t0 + static_cast<int64_t>(w) * kTrkLen;
        float r[kTrkLen];
        uint32_t alive = 0;  // bit i: the carry survives step lo + i
#pragma unroll
        for (int i = 0; i < kTrkLen; ++i) {
            const int64_t t = lo + i;
            r[i] = 0.f;
            if (live && t < p.steps) {
                r[i] = __ldg(p.rewards + t * p.r_rs + env);
                bool d = false;
                if (DONE == PPO_DONE_U8) d = __ldg(static_cast<const unsigned char*>(p.dones) + t * p.d_rs + env) != 0;
                else if (DONE == PPO_DONE_F32) d = __ldg(static_cast<const float*>(p.dones) + t * p.d_rs + env) != 0.f;
                alive |= d ? 0u : (1u << i);
            }
        }
        // phase 1: the chunk as an affine map of its carry-in
        double B = 0.0, C = 1.0;
#pragma unroll
        for (int i = 0; i < kTrkLen; ++i) {
            if (lo + i < p.steps) {
                const double g = (alive >> i) & 1u ? p.gamma : 0.0;
                B = __dadd_rn(__dmul_rn(g, B), static_cast<double>(r[i]));
                C = g * C;
            }
        }
        sB[w][lane] = B, sC[w][lane] = C;
        __syncthreads();
        // phase 2: carry-in of this chunk = the earlier chunks applied to the carry of the pass
        double x = x_in;
        for (int j = 0; j < w; ++j) x = sB[j][lane] + sC[j][lane] * x;
        double x_end = x;
        for (int j = w; j < kTrkChunks; ++j) x_end = sB[j][lane] + sC[j][lane] * x_end;
        // phase 3: replay, one float64 row element per step
#pragma unroll
        for (int i = 0; i < kTrkLen; ++i) {
            const int64_t t = lo + i;
            if (live && t < p.steps) {
                const double g = (alive >> i) & 1u ? p.gamma : 0.0;
                x = __dadd_rn(__dmul_rn(g, x), static_cast<double>(r[i]));
                p.x[t * p.envs + env] = x;
            }
        }
        x_in = x_end;
        __syncthreads();
    }
    if (live && w == 0) p.carry[env] = x_in;
}

// out[t] = sum_a x[t, a]                      (center == nullptr)
//        = sum_a (x[t, a] - center[t] / n)^2  (center = row sums, n = total env count)
__global__ void __launch_bounds__(256) row_sum_kernel(const double* __restrict__ x, int64_t envs, const double* center,
                                                      double n, double* __restrict__ out) {
    __shared__ double red[8];
    const int64_t t = blockIdx.x;
    const double* row = x + t * envs;
    const double c = center ? center[t] / n : 0.0;
    double acc = 0.0;
    for (int64_t a = threadIdx.x; a < envs; a += blockDim.x) {
        const double v = __ldcg(row + a) - c;
        acc += center ? v * v : v;
    }
    acc = warp_sum_f64(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        out[t] = s;
    }
}

// rl/utils.py:379-396 applied N times in order; rms = {mean, var, count}
// ema_var (optional, in / out): the wrapper's mode "ema" (rl/wrappers.py:878-881) -- after every merge
//   alpha = 1 - n / min(count, ema_horizon);  ema = alpha * ema + (1 - alpha) * var_over_envs(x_t)
// and row_var then records the EMA instead of the running variance.
struct RmsFold {
    double mean, var, count, ema;
    // one merge of rl/utils.py:379-396 with the batch moments (bm, bv) of n samples; returns the variance to record
    __device__ __forceinline__ double step(double bm, double bv, double n, bool use_ema, double ema_horizon) {
        const double delta = __dsub_rn(bm, mean);
        const double tot = __dadd_rn(count, n);
        const double new_mean = __dadd_rn(mean, __ddiv_rn(__dmul_rn(delta, n), tot));
        const double m_a = __dmul_rn(var, count), m_b = __dmul_rn(bv, n);
        const double cross = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), count), n), tot);
        const double m2 = __dadd_rn(__dadd_rn(m_a, m_b), cross);
        mean = new_mean, var = __ddiv_rn(m2, tot), count = tot;
        if (use_ema) {
            const double alpha = 1.0 - n / fmin(count, ema_horizon);
            ema = alpha * ema + (1.0 - alpha) * bv;
        }
        return use_ema ? ema : var;
    }
};

__global__ void rms_fold_kernel(const double* __restrict__ row_sum, const double* __restrict__ row_sq, int64_t steps,
                                double n, double* rms, double* row_var, double* ema_var, double ema_horizon) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    RmsFold f{rms[0], rms[1], rms[2], ema_var ? *ema_var : 0.0};
    for (int64_t t = 0; t < steps; ++t) {
        const double v = f.step(row_sum[t] / n, row_sq[t] / n, n, ema_var != nullptr, ema_horizon);
        if (row_var) row_var[t] = v;
    }
    rms[0] = f.mean, rms[1] = f.var, rms[2] = f.count;
    if (ema_var) *ema_var = f.ema;
}

// One GPU (no exchange of the row vectors): both passes of a row in one CTA -- the row sum, then the squared deviations
// from the row mean out of L1 / L2 -- and the N sequential merges by the last CTA to finish (ticket), which first
// stages the row vectors in shared memory (a dependent L2 load per merge would triple the chain).  Same reduction
// order as row_sum_kernel, so the result is bit-identical to the three launches.
constexpr int kFoldStage = 512;
__global__ void __launch_bounds__(256) row_moments_fold_kernel(const double* __restrict__ x, int64_t steps, int64_t envs,
                                                               double* row_sum, double* row_sq, double* rms,
                                                               double* row_var, double* ema_var, double ema_horizon,
                                                               unsigned int* ticket) {
    __shared__ double red[8];
    __shared__ double s_c;
    __shared__ bool s_last;
    __shared__ double stage[2][kFoldStage];
    const int64_t t = blockIdx.x;
    const double* row = x + t * envs;
    const double n = static_cast<double>(envs);
    double total[2];
    for (int pass = 0; pass < 2; ++pass) {
        const double c = pass ? s_c : 0.0;
        double acc = 0.0;
        for (int64_t a = threadIdx.x; a < envs; a += blockDim.x) {
            const double v = __ldcg(row + a) - c;
            acc += pass ? v * v : v;
        }
        acc = warp_sum_f64(acc);
        __syncthreads();  // red[] of the previous pass has been read
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = 0; i < 8; ++i) s += red[i];
            total[pass] = s;
            if (pass == 0) s_c = s / n;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        row_sum[t] = total[0], row_sq[t] = total[1];
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    RmsFold f{rms[0], rms[1], rms[2], ema_var ? *ema_var : 0.0};
    for (int64_t t0 = 0; t0 < steps; t0 += kFoldStage) {
        const int m = static_cast<int>(min(static_cast<int64_t>(kFoldStage), steps - t0));
        for (int i = threadIdx.x; i < m; i += blockDim.x)
            stage[0][i] = __ldcg(row_sum + t0 + i), stage[1][i] = __ldcg(row_sq + t0 + i);
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < m; ++i) {
                const double v = f.step(stage[0][i] / n, stage[1][i] / n, n, ema_var != nullptr, ema_horizon);
                if (row_var) row_var[t0 + i] = v;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rms[0] = f.mean, rms[1] = f.var, rms[2] = f.count;
        if (ema_var) *ema_var = f.ema;
        *ticket = 0u;
    }
}

struct ScaleParams {
    const float* rewards;
    int64_t r_rs, steps, envs;
    const double* var;  // one running variance (row_var == 0) or one per row
    int per_row;
    int mode;           // PPO_SCALE_INTRINSIC: eps + sqrt(var);  PPO_SCALE_WRAPPER: sqrt(var + eps)
    double eps, clip, scale;
    const double* center_stats;  // ppo_moments record of the raw rewards (centre on their mean), or nullptr
    float* out;
    unsigned long long* clips;
};

__global__ void __launch_bounds__(256) scale_rewards_kernel(const ScaleParams p) {
    const int64_t total = p.steps * p.envs;
    unsigned int my_clips = 0;
    double shift = 0.0;
    if (p.center_stats) shift = p.center_stats[PPO_STAT_ADV_SUM] / fmax(p.center_stats[PPO_STAT_COUNT], 1.0);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t t = i / p.envs, a = i - t * p.envs;
        const double v = p.var[p.per_row ? t : 0];
        const double s = p.mode == PPO_SCALE_WRAPPER ? sqrt(v + p.eps) : p.eps + sqrt(v);
        double y = (static_cast<double>(__ldg(p.rewards + t * p.r_rs + a)) - shift) / s;
        if (p.clip >= 0.0) {
            const double z = fmin(fmax(y, -p.clip), p.clip);
            my_clips += z != y;
            y = z;
        }
        p.out[i] = static_cast<float>(y * p.scale);
    }
    if (p.clips) {
        my_clips = __reduce_add_sync(0xffffffffu, my_clips);
        if ((threadIdx.x & 31) == 0 && my_clips) atomicAdd(p.clips, static_cast<unsigned long long>(my_clips));
    }
}

// ---- ReturnTracker of rl/wrappers.py:737-755 (the multi-env reward normaliser, :946-971): the same forward
// recurrence, but only the envs of a per-step MASK move, the state is float32 and is rounded after every step, and
// the running moments take the batch moments of the masked envs only (a step without masked envs changes nothing).
// The reference's arithmetic, step by step (NumPy, rewards float32):  gamma * cur -> float32;  * (1 - dones) -> float64
// for boolean dones (int64 mask), float32 for float32 dones;  rewards + that -> float64 / float32;  stored as float32.
struct MaskedTrackerParams {
    const float* rewards;
    int64_t r_rs;
    const void* dones;
    int done_kind;
    int64_t d_rs;
    const unsigned char* mask;
    int64_t m_rs;
    int64_t steps, envs;
    float gamma32;
    float* carry;  // [A] float32 in / out (ReturnTracker.current_returns)
    double* x;     // [N, A] the state after every step (unchanged where the mask is off)
};

__global__ void __launch_bounds__(128) masked_tracker_scan_kernel(const MaskedTrackerParams p) {
    const int64_t env = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (env >= p.envs) return;
    float cur = p.carry[env];
    for (int64_t t = 0; t < p.steps; ++t) {
        if (__ldg(p.mask + t * p.m_rs + env)) {
            const float r = __ldg(p.rewards + t * p.r_rs + env);
            const float gc = __fmul_rn(p.gamma32, cur);
            if (p.done_kind == PPO_DONE_F32) {
                const float d = __ldg(static_cast<const float*>(p.dones) + t * p.d_rs + env);
                cur = __fadd_rn(r, __fmul_rn(gc, __fsub_rn(1.0f, d)));
            } else {
                const bool d = p.done_kind == PPO_DONE_U8 && __ldg(static_cast<const unsigned char*>(p.dones) + t * p.d_rs + env) != 0;
                cur = static_cast<float>(__dadd_rn(static_cast<double>(r), d ? 0.0 : static_cast<double>(gc)));
            }
        }
        p.x[t * p.envs + env] = static_cast<double>(cur);
    }
    p.carry[env] = cur;
}

// One CTA per step: count, mean and variance of the masked envs (np.mean / np.var of a float32 vector: the results are
// rounded to float32 like NumPy's), then the last CTA to finish merges the rows in order (rl/utils.py:379-396), skipping
// the rows without a masked env (rl/wrappers.py:750-751); row_var records the running variance after every step.
__global__ void __launch_bounds__(256) masked_row_moments_fold_kernel(const double* __restrict__ x,
                                                                      const unsigned char* __restrict__ mask, int64_t m_rs,
                                                                      int64_t steps, int64_t envs, double* row_rec,
                                                                      double* rms, double* row_var, unsigned int* ticket) {
    __shared__ double red[8], redn[8];
    __shared__ double s_c;
    __shared__ bool s_last;
    const int64_t t = blockIdx.x;
    const double* row = x + t * envs;
    const unsigned char* mrow = mask + t * m_rs;
    double total[2] = {0.0, 0.0}, count = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        const double c = pass ? s_c : 0.0;
        double acc = 0.0, n = 0.0;
        for (int64_t a = threadIdx.x; a < envs; a += blockDim.x) {
            if (__ldg(mrow + a)) {
                const double v = __ldcg(row + a) - c;
                acc += pass ? v * v : v;
                n += 1.0;
            }
        }
        acc = warp_sum_f64(acc), n = warp_sum_f64(n);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc, redn[threadIdx.x >> 5] = n;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0, m = 0.0;
            for (int i = 0; i < 8; ++i) s += red[i], m += redn[i];
            total[pass] = s, count = m;
            if (pass == 0) s_c = m > 0.0 ? static_cast<double>(static_cast<float>(s / m)) : 0.0;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        row_rec[3 * t] = count;
        row_rec[3 * t + 1] = s_c;                                                                    // float32 mean
        row_rec[3 * t + 2] = count > 0.0 ? static_cast<double>(static_cast<float>(total[1] / count)) : 0.0;  // float32 var
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    RmsFold f{rms[0], rms[1], rms[2], 0.0};
    for (int64_t i = 0; i < steps; ++i) {
        const double n = __ldcg(row_rec + 3 * i);
        if (n > 0.0) f.step(__ldcg(row_rec + 3 * i + 1), __ldcg(row_rec + 3 * i + 2), n, false, 1.0);
        if (row_var) row_var[i] = f.var;
    }
    rms[0] = f.mean, rms[1] = f.var, rms[2] = f.count;
    *ticket = 0u;
}

}  // namespace ppo

using namespace ppo;

extern "C" int ppo_return_tracker_scan(const float* rewards, int64_t reward_row_stride, const void* dones,
                                       int done_kind, int64_t done_row_stride, int64_t steps, int64_t envs,
                                       double gamma, double* carry, double* tracker_out, void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_return_tracker_scan: negative shape [%lld, %lld]", (long long)steps,
                (long long)envs);
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32,
                "ppo_return_tracker_scan: unknown done_kind %d", done_kind);
    if (steps == 0 || envs == 0) return PPO_OK;
    PPO_REQUIRE(rewards && carry && tracker_out, "ppo_return_tracker_scan: null pointer");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || dones, "ppo_return_tracker_scan: dones is null but done_kind=%d", done_kind);
    TrackerParams p{rewards, reward_row_stride, dones, done_kind, done_row_stride, steps, envs, gamma, carry, tracker_out};
    const unsigned grid = static_cast<unsigned>((envs + kTrkEnvs - 1) / kTrkEnvs);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (done_kind == PPO_DONE_U8) tracker_scan_kernel<PPO_DONE_U8><<<grid, kTrkEnvs * kTrkChunks, 0, s>>>(p);
    else if (done_kind == PPO_DONE_F32) tracker_scan_kernel<PPO_DONE_F32><<<grid, kTrkEnvs * kTrkChunks, 0, s>>>(p);
    else tracker_scan_kernel<PPO_DONE_NONE><<<grid, kTrkEnvs * kTrkChunks, 0, s>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_row_sums(const double* x, int64_t steps, int64_t envs, const double* row_sums, double total_envs,
                            double* out, void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_row_sums: negative shape");
    if (steps == 0) return PPO_OK;
    PPO_REQUIRE(x && out, "ppo_row_sums: null pointer");
    PPO_REQUIRE(!row_sums || total_envs > 0.0, "ppo_row_sums: total_envs must be positive");
    row_sum_kernel<<<static_cast<unsigned>(steps), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, envs, row_sums,
                                                                                                total_envs, out);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_rms_fold(const double* row_sums, const double* row_sq, int64_t steps, double total_envs,
                            double* rms, double* row_var, void* stream) {
    PPO_REQUIRE(steps >= 0 && total_envs > 0.0, "ppo_rms_fold: bad shape");
    PPO_REQUIRE(rms && (steps == 0 || (row_sums && row_sq)), "ppo_rms_fold: null pointer");
    rms_fold_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(row_sums, row_sq, steps, total_envs, rms, row_var,
                                                                     nullptr, 0.0);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_row_moments_fold(const double* x, int64_t steps, int64_t envs, double* row_sums, double* row_sq,
                                    double* rms, double* ema_var, double ema_horizon, double* row_var,
                                    unsigned int* ticket, void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 1, "ppo_row_moments_fold: bad shape");
    if (steps == 0) return PPO_OK;
    PPO_REQUIRE(x && row_sums && row_sq && rms && ticket, "ppo_row_moments_fold: null pointer");
    PPO_REQUIRE(!ema_var || ema_horizon > 0.0, "ppo_row_moments_fold: ema_horizon must be positive");
    row_moments_fold_kernel<<<static_cast<unsigned>(steps), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, steps, envs, row_sums, row_sq, rms, row_var, ema_var, ema_horizon, ticket);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_rms_fold_ema(const double* row_sums, const double* row_sq, int64_t steps, double total_envs,
                                double* rms, double* ema_var, double ema_horizon, double* row_var, void* stream) {
    PPO_REQUIRE(steps >= 0 && total_envs > 0.0 && ema_horizon > 0.0, "ppo_rms_fold_ema: bad shape");
    PPO_REQUIRE(rms && ema_var && (steps == 0 || (row_sums && row_sq)), "ppo_rms_fold_ema: null pointer");
    rms_fold_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(row_sums, row_sq, steps, total_envs, rms, row_var,
                                                                     ema_var, ema_horizon);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_scale_rewards(const float* rewards, int64_t reward_row_stride, int64_t steps, int64_t envs,
                                 const double* var, int per_row, int mode, double eps, double clip, double scale,
                                 const double* center_stats, float* out, unsigned long long* clips, void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_scale_rewards: negative shape");
    PPO_REQUIRE(mode == PPO_SCALE_INTRINSIC || mode == PPO_SCALE_WRAPPER, "ppo_scale_rewards: unknown mode %d", mode);
    if (steps == 0 || envs == 0) return PPO_OK;
    PPO_REQUIRE(rewards && var && out, "ppo_scale_rewards: null pointer");
    ScaleParams p{rewards, reward_row_stride, steps, envs, var, per_row, mode, eps, clip, scale, center_stats, out, clips};
    int64_t blocks = (steps * envs + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    scale_rewards_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_return_tracker_masked(const float* rewards, int64_t reward_row_stride, const void* dones,
                                         int done_kind, int64_t done_row_stride, const unsigned char* mask,
                                         int64_t mask_row_stride, int64_t steps, int64_t envs, double gamma,
                                         float* carry, double* tracker_out, double* row_records, double* rms,
                                         double* row_var, unsigned int* ticket, void* stream) {
    PPO_REQUIRE(steps >= 0 && envs >= 0, "ppo_return_tracker_masked: negative shape [%lld, %lld]", (long long)steps,
                (long long)envs);
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || done_kind == PPO_DONE_U8 || done_kind == PPO_DONE_F32,
                "ppo_return_tracker_masked: unknown done_kind %d", done_kind);
    if (steps == 0 || envs == 0) return PPO_OK;
    PPO_REQUIRE(rewards && mask && carry && tracker_out && row_records && rms && ticket,
                "ppo_return_tracker_masked: null pointer");
    PPO_REQUIRE(done_kind == PPO_DONE_NONE || dones, "ppo_return_tracker_masked: dones is null but done_kind=%d", done_kind);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MaskedTrackerParams p{rewards, reward_row_stride, dones, done_kind, done_row_stride, mask, mask_row_stride, steps, envs,
                          static_cast<float>(gamma), carry, tracker_out};
    masked_tracker_scan_kernel<<<static_cast<unsigned>((envs + 127) / 128), 128, 0, s>>>(p);
    PPO_CUDA(cudaGetLastError());
    masked_row_moments_fold_kernel<<<static_cast<unsigned>(steps), 256, 0, s>>>(tracker_out, mask, mask_row_stride, steps,
                                                                              envs, row_records, rms, row_var, ticket);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
