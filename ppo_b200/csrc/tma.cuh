// Thin wrappers over the sm_100a async-copy machinery: mbarrier, TMA tiled loads, tensor-map
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace ppo {

// ---- host: tensor maps ---------------------------------------------------------------------
// Encode a tiled tensor map of up to 3 dimensions (dims[0] is the contiguous one).  strides are in
// BYTES for dims 1.. (dim 0 is packed).  Returns PPO_OK / PPO_ERR_CUDA / PPO_ERR_UNSUPPORTED.
int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      bool swizzle_128b = false);
// the same with the shared-memory swizzle span in bytes (0 = none, 32, 64, 128)
int encode_tensor_map_swz(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base,
                          const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// ---- device: mbarrier ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the
// hint expires) instead of burning issue slots on a polling loop.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}

// for warps that are far ahead of their consumers (producer, scan): back off between polls
// (every poll is a shared-memory access: warps that wait for microseconds should poll rarely, or they take a
// measurable share of the shared-memory pipe from the warps that do the work)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns = 64) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---- device: TMA loads (global -> shared, completion on an mbarrier) ------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---- device: 4-byte asynchronous copies (global -> shared) tracked by an mbarrier -------------
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// Arrive on `bar` once all cp.async copies issued so far by this thread have landed.  `.noinc`: the
// arrival is one of the barrier's expected arrivals (init count = number of copying threads).
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- device: TMA stores (shared -> global, bulk-group completion) ------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

}  // namespace ppo
