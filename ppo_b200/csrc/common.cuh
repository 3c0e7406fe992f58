// Shared helpers for the ppo_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ppo_b200.h"

namespace ppo {

// ---- per-thread error string -------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);

#define PPO_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            ::ppo::set_error(__VA_ARGS__); \
            return PPO_ERR_BAD_ARG;       \
        }                                 \
    } while (0)

#define PPO_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return ::ppo::cuda_fail(e__, #call); \
    } while (0)

// Device properties of the current device, cached per device.
int sm_count();

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ double shfl_down_f64(double v, int delta) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(0xffffffffu, lo, delta);
    hi = __shfl_down_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += shfl_down_f64(v, d);
    return v;  // valid in lane 0
}

__device__ __forceinline__ float warp_min_f32(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}

__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, d));
    return v;
}

// (1 - done) as the reference computes it, widened to double.  The bool case is exact 0/1;
// the float32 case keeps the float32 subtraction of `1.0 - terminal` (rl/returns.py:26).
template <int KIND>
__device__ __forceinline__ double alive_of(const void* dones, int64_t idx) {
    if (KIND == PPO_DONE_U8) {
        return static_cast<const unsigned char*>(dones)[idx] ? 0.0 : 1.0;
    } else if (KIND == PPO_DONE_F32) {
        return static_cast<double>(1.0f - static_cast<const float*>(dones)[idx]);
    } else {
        return 1.0;
    }
}

// (1 - done) as a float32 mask with a runtime kind (rl/vtrace.py:78 `np.asarray(1 - dones, dtype=np.float32)`)
__device__ __forceinline__ float alive_f32(const void* dones, int kind, int64_t idx) {
    if (kind == PPO_DONE_U8) return static_cast<const unsigned char*>(dones)[idx] ? 0.f : 1.f;
    if (kind == PPO_DONE_F32) return 1.0f - static_cast<const float*>(dones)[idx];
    return 1.f;
}

// streaming store: outputs are written once and not re-read by the producing kernel
__device__ __forceinline__ void store_stream(float* p, float v) { __stcs(p, v); }

}  // namespace ppo
