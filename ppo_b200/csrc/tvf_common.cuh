// Pieces shared by the TVF target kernels (generic gather kernel and TMA ring kernel).
#pragma once

#include "common.cuh"

namespace ppo {

constexpr int kNoDone = 1 << 30;  // "no terminal before the end of the window"

struct TvfParams {
    const float* rewards;
    int64_t r_rs;
    const void* dones;
    int64_t d_rs;
    int done_kind;
    const float* vs;  // [N+1, A, V]
    int64_t vs_rs, vs_es;
    int steps, n_req, n_val;
    int64_t envs;
    ppo_tvf_tables tab;
    float* out;  // [N, A, K]
};

__device__ __forceinline__ bool done_at(const void* dones, int kind, int64_t idx) {
    if (kind == PPO_DONE_U8) return static_cast<const unsigned char*>(dones)[idx] != 0;
    if (kind == PPO_DONE_F32) return static_cast<const float*>(dones)[idx] != 0.f;
    return false;
}

// lerp of rl/returns_truncated.py:174 in float32: v_lo * (1 - f) + v_hi * f with both weights
// resolved on the host (zero weights encode "horizon <= 0 -> 0", :151-153).
__device__ __forceinline__ float lerp2(float v_lo, float v_hi, float w_lo, float w_hi) {
    return fmaf(v_hi, w_hi, v_lo * w_lo);
}

}  // namespace ppo
