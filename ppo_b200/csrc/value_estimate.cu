// Value estimates read off the truncated value curves: rediscounting to another gamma.
//
// Replaces rl/tvf.py:388-433 `get_rediscounted_value_estimate` and rl/tvf.py:362-380 `rediscount_horizons`:
//     est[b]    = sum_{i=1..K-1} (V[b,i] - V[b,i-1]) * rho_i,   rho_i = min((g_new / g_old)^mid_i, clip)
//     pref[b,k] = sum_{i=1..k}   (V[b,i] - V[b,i-1]) * rho_i
// The ratios are host floats (Python computes them exactly like the reference, float64 then float32).  `est` is
// evaluated as one dot product per row with w_0 = -rho_1, w_i = rho_i - rho_{i+1}, w_{K-1} = rho_{K-1} (the same sum
// re-associated; float32 accumulation like the reference's torch float32 tensors) so that it can also ride on the
// value rows the TVF target kernel already streams (ppo_tvf_targets_ex); this file is the standalone form.
// Both kernels are one warp per row, coalesced 128-byte row segments, HBM-bound: 4 K + 4 bytes per row.
#include "common.cuh"

namespace ppo {

constexpr int kMaxValueHorizons = 1024;

struct DotParams {
    const float* v;   // element (r, a, j) at v[r * rs_t + a * rs_e + j]
    int64_t rows_t, envs, rs_t, rs_e, os_t;
    int k;
    int last_only;    // 1: out = the last column (no arithmetic)
    float* out;       // (r, a) at out[r * os_t + a]
    float w[kMaxValueHorizons];
};

__global__ void __launch_bounds__(256) value_dot_kernel(const __grid_constant__ DotParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t total = p.rows_t * p.envs;
    if (p.last_only) {
        const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int64_t r = i / p.envs, a = i - r * p.envs;
            p.out[r * p.os_t + a] = __ldcs(p.v + r * p.rs_t + a * p.rs_e + (p.k - 1));
        }
        return;
    }
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp; row < total; row += nwarps) {
        const int64_t r = row / p.envs, a = row - r * p.envs;
        const float* src = p.v + r * p.rs_t + a * p.rs_e;
        float acc = 0.f;
        for (int j = lane; j < p.k; j += 32) acc = fmaf(__ldcs(src + j), p.w[j], acc);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) p.out[r * p.os_t + a] = acc;
    }
}

struct PrefixParams {
    const float* v;
    int64_t rows, rs, os;
    int k;
    float* out;
    float rho[kMaxValueHorizons];  // rho[0] unused
};

// lane l owns columns [l * seg, (l + 1) * seg): local running sum, exclusive warp scan of the segment totals
__global__ void __launch_bounds__(256) value_prefix_kernel(const __grid_constant__ PrefixParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int seg = (p.k + 31) / 32;
    for (int64_t row = warp; row < p.rows; row += nwarps) {
        const float* src = p.v + row * p.rs;
        float* dst = p.out + row * p.os;
        const int lo = lane * seg, hi = min(p.k, lo + seg);
        float total = 0.f;
        for (int j = lo; j < hi; ++j)
            if (j >= 1) total += __fmul_rn(__fsub_rn(src[j], src[j - 1]), p.rho[j]);
        float incl = total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        float run = incl - total;  // sum of the segments before this lane
        for (int j = lo; j < hi; ++j) {
            if (j >= 1) run += __fmul_rn(__fsub_rn(src[j], src[j - 1]), p.rho[j]);
            dst[j] = run;
        }
    }
}

// Per-row linear interpolation of a value curve at a per-row target horizon (rl/tvf.py:527-573 `horizon_interpolate`):
// clip the target to [h_0, h_{K-1}], post = searchsorted(h, x, left), pre = max(post - 1, 0),
// f = (x - h_pre) / (h_post - h_pre) (1 where the width is 0), out = v_pre (1 - f) + v_post f in float64, 0 where
// post == 0.  The horizon axis arrives as doubles (the reference also calls this on a log-scaled axis,
// rl/tvf.py:143-152); one thread per row, the axis in shared memory, two value reads per row.
struct InterpParams {
    const float* v;
    int64_t rows, rs;
    int k;
    const double* targets;
    double* out;
    double h[kMaxValueHorizons];
};

__global__ void __launch_bounds__(256) horizon_interpolate_kernel(const __grid_constant__ InterpParams p) {
    __shared__ double sh[kMaxValueHorizons];
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) sh[j] = p.h[j];
    __syncthreads();
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < p.rows; row += stride) {
        double x = p.targets[row];
        x = fmin(fmax(x, sh[0]), sh[p.k - 1]);
        int lo = 0, hi = p.k;  // first index with h[index] >= x
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sh[mid] < x) lo = mid + 1;
            else hi = mid;
        }
        const int post = lo, pre = post > 0 ? post - 1 : 0;
        double width = sh[post] - sh[pre];
        if (width == 0.0) width = 1.0;
        const double f = (x - sh[pre]) / width;
        const double a = static_cast<double>(p.v[row * p.rs + pre]), b = static_cast<double>(p.v[row * p.rs + post]);
        p.out[row] = post == 0 ? 0.0 : __dadd_rn(__dmul_rn(a, 1.0 - f), __dmul_rn(b, f));
    }
}

// Horizon trimming of one rollout step (rl/tvf.py:91-208): value estimates of horizons beyond the time left are pulled
// back to the horizon that is still reachable.  Everything that is an index (the trimming point `searchsorted(h, ttt)`,
// the interpolation slot on the log-scaled axis, the random horizon picks) is resolved on the host; this kernel does
// the arithmetic on the value estimates.  One CTA per env, one thread per (horizon, head) where that parallelises and
// one thread per head for the sequential running mean of the `average` mode (float32 accumulation in the
// reference's order).
struct TrimParams {
    const float* v;        // [A, K, VH]
    int64_t envs;
    int k, vh, mode;       // mode: 0 interpolate, 1 average, 2 substitute, 3 random
    const int32_t* tk;     // [A] trimming point
    const int32_t* pre;    // [A] interpolate: lower slot, upper slot, factor
    const int32_t* post;
    const double* frac;
    const int32_t* picks;  // [A, K] random: source horizon of every horizon
    float clip;            // < 0: off
    float* out;            // [A, K, VH] ([A, K, 1] for substitute)
    float* final_value;    // [A]
};

__global__ void __launch_bounds__(256) trim_horizons_kernel(const TrimParams p) {
    const int64_t a = blockIdx.x;
    const float* src = p.v + a * p.k * p.vh;
    const int K = p.k, VH = p.vh, tk = p.tk[a];
    const int ovh = p.mode == 2 ? 1 : VH;
    float* dst = p.out + a * K * ovh;
    auto old_at = [&](int k, int h) { return k == 0 ? 0.f : src[k * VH + h]; };  // h = 0 is zero by definition (:111-112)
    auto clipped = [&](float nv, float ov) {
        return p.clip >= 0.f ? __fadd_rn(ov, fminf(fmaxf(__fsub_rn(nv, ov), -p.clip), p.clip)) : nv;  // :199-201
    };
    if (p.mode == 1) {
        // running mean from the trimming point on, every head (:156-171)
        for (int h = threadIdx.x; h < VH; h += blockDim.x) {
            float acc = 0.f;
            int cnt = 0;
            for (int k = 0; k < K; ++k) {
                const float ov = old_at(k, h);
                float nv = ov;
                if (tk < K - 1 && k >= tk) {
                    acc = __fadd_rn(acc, ov);
                    ++cnt;
                    nv = __fdiv_rn(acc, static_cast<float>(cnt));
                }
                dst[k * VH + h] = clipped(nv, ov);
            }
        }
    } else {
        float iv = 0.f;
        if (p.mode == 0) {  // horizon_interpolate on the log axis, float64 like the reference, rounded on assignment
            const int lo = p.pre[a], hi = p.post[a];
            const double f = p.frac[a];
            const double x = __dadd_rn(__dmul_rn(static_cast<double>(old_at(lo, 0)), 1.0 - f),
                                       __dmul_rn(static_cast<double>(old_at(hi, 0)), f));
            iv = hi == 0 ? 0.f : static_cast<float>(x);
        }
        for (int i = threadIdx.x; i < K * ovh; i += blockDim.x) {
            const int k = i / ovh, h = i % ovh;
            const float ov = old_at(k, h);
            float nv = ov;
            if (p.mode == 0) {
                if (h == 0 && k >= tk) nv = iv;                                    // :151-155
            } else if (p.mode == 2) {
                nv = old_at(min(tk, k), 0);                                        // :172-178
            } else {
                nv = old_at(p.picks[a * K + k], h);                                // :179-189
            }
            dst[i] = clipped(nv, ov);
        }
    }
    // estimate used for the advantages: mean of the untrimmed ext estimates that are still valid (:193-196)
    __shared__ double part[8];
    const int k0 = min(tk, K - 1);
    double s = 0.0;
    for (int k = k0 + threadIdx.x; k < K; k += blockDim.x) s += static_cast<double>(old_at(k, 0));
    s = warp_sum_f64(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) tot += part[w];
        p.final_value[a] = static_cast<float>(tot / static_cast<double>(K - k0));
    }
}

static int64_t warp_grid(int64_t rows) {
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}

// host_weights == nullptr: the last column
int launch_value_dot(const float* v, int64_t rows_t, int64_t envs, int64_t rs_t, int64_t rs_e, int k,
                     const float* host_weights, float* out, int64_t os_t, cudaStream_t s) {
    if (k > kMaxValueHorizons) {
        set_error("value estimate: more than %d value horizons", kMaxValueHorizons);
        return PPO_ERR_UNSUPPORTED;
    }
    DotParams p{};
    p.v = v, p.rows_t = rows_t, p.envs = envs, p.rs_t = rs_t, p.rs_e = rs_e, p.os_t = os_t, p.k = k, p.out = out;
    p.last_only = host_weights ? 0 : 1;
    if (host_weights)
        for (int j = 0; j < k; ++j) p.w[j] = host_weights[j];
    const int64_t units = p.last_only ? (rows_t * envs + 31) / 32 : rows_t * envs;
    value_dot_kernel<<<static_cast<unsigned>(warp_grid(units)), 256, 0, s>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

}  // namespace ppo

using namespace ppo;

extern "C" int ppo_value_estimate(const float* values, int64_t rows, int64_t row_stride, int64_t n_val,
                                  const float* host_weights, float* out, void* stream) {
    PPO_REQUIRE(rows >= 0 && n_val >= 1 && row_stride >= n_val, "ppo_value_estimate: bad shape");
    if (rows == 0) return PPO_OK;
    PPO_REQUIRE(values && host_weights && out, "ppo_value_estimate: null pointer");
    return launch_value_dot(values, 1, rows, 0, row_stride, static_cast<int>(n_val), host_weights, out, 0,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int ppo_rediscount_horizons(const float* values, int64_t rows, int64_t row_stride, int64_t n_val,
                                       const float* host_ratios, float* out, int64_t out_row_stride, void* stream) {
    PPO_REQUIRE(rows >= 0 && n_val >= 1 && row_stride >= n_val && out_row_stride >= n_val,
                "ppo_rediscount_horizons: bad shape");
    if (rows == 0) return PPO_OK;
    PPO_REQUIRE(values && host_ratios && out, "ppo_rediscount_horizons: null pointer");
    if (n_val > kMaxValueHorizons) {
        set_error("ppo_rediscount_horizons: more than %d value horizons", kMaxValueHorizons);
        return PPO_ERR_UNSUPPORTED;
    }
    PrefixParams p{};
    p.v = values, p.rows = rows, p.rs = row_stride, p.os = out_row_stride, p.k = static_cast<int>(n_val), p.out = out;
    for (int j = 0; j < p.k; ++j) p.rho[j] = host_ratios[j];
    value_prefix_kernel<<<static_cast<unsigned>(warp_grid(rows)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_horizon_interpolate(const double* host_horizons, int64_t n_val, const float* values, int64_t rows,
                                       int64_t row_stride, const double* targets, double* out, void* stream) {
    PPO_REQUIRE(rows >= 0 && n_val >= 1 && row_stride >= n_val, "ppo_horizon_interpolate: bad shape");
    if (rows == 0) return PPO_OK;
    PPO_REQUIRE(host_horizons && values && targets && out, "ppo_horizon_interpolate: null pointer");
    if (n_val > kMaxValueHorizons) {
        set_error("ppo_horizon_interpolate: more than %d value horizons", kMaxValueHorizons);
        return PPO_ERR_UNSUPPORTED;
    }
    InterpParams p{};
    p.v = values, p.rows = rows, p.rs = row_stride, p.k = static_cast<int>(n_val), p.targets = targets, p.out = out;
    for (int j = 0; j < p.k; ++j) p.h[j] = host_horizons[j];
    int64_t blocks = (rows + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    horizon_interpolate_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_tvf_trim_horizons(const float* values, int64_t envs, int64_t n_val, int64_t n_heads, int mode,
                                     const int32_t* trim_index, const int32_t* interp_pre, const int32_t* interp_post,
                                     const double* interp_frac, const int32_t* picks, double clip, float* out,
                                     float* final_value, void* stream) {
    PPO_REQUIRE(envs >= 0 && n_val >= 1 && n_heads >= 1 && mode >= 0 && mode <= 3, "ppo_tvf_trim_horizons: bad arguments");
    if (envs == 0) return PPO_OK;
    PPO_REQUIRE(values && trim_index && out && final_value, "ppo_tvf_trim_horizons: null pointer");
    PPO_REQUIRE(mode != 0 || (interp_pre && interp_post && interp_frac), "ppo_tvf_trim_horizons: interpolation slots missing");
    PPO_REQUIRE(mode != 3 || picks, "ppo_tvf_trim_horizons: random picks missing");
    TrimParams p{};
    p.v = values, p.envs = envs, p.k = static_cast<int>(n_val), p.vh = static_cast<int>(n_heads), p.mode = mode;
    p.tk = trim_index, p.pre = interp_pre, p.post = interp_post, p.frac = interp_frac, p.picks = picks;
    p.clip = clip >= 0 ? static_cast<float>(clip) : -1.f;
    p.out = out, p.final_value = final_value;
    trim_horizons_kernel<<<static_cast<unsigned>(envs), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
