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

// Launch constants of the transposed slab kernel (tvf_tslab.cu), resolved on the host by ppo_tvf_tables_build.
constexpr int kTsMaxK = 128;
constexpr int kTsMaxEntries = 1024;
constexpr int kTsConstMagic = 0x54534c42;  // header word of the constants section: magic, n_ent (-1 = none), pitch, V
struct __align__(16) TsEntry {
    uint32_t off;    // byte offset of W^T[col][n - 1] in the slab
    float wlo, whi;  // multiplicity * gamma^n * lerp weights / samples per horizon
    int32_t n;
};
struct __align__(16) TsHorizon {
    int32_t beg, cnt;  // entries [beg, beg + cnt)
    float rscale;      // 1 for a live horizon, 0 for h = 0 (all-zero targets)
    int32_t maxn;      // largest n-step of the horizon (tail rows: N - t < maxn)
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
