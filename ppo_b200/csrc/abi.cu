// Library-wide pieces of the C ABI: version, per-thread error text, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ppo {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
    set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), where);
    return PPO_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

}  // namespace ppo

extern "C" int ppo_version(void) { return 100; }

extern "C" const char* ppo_last_error(void) { return ppo::g_error; }

extern "C" int ppo_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return ppo::cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

extern "C" int ppo_memcpy_2d_async(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                                   int64_t width_bytes, int64_t rows, void* stream) {
    PPO_REQUIRE(width_bytes >= 0 && rows >= 0 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
                "ppo_memcpy_2d_async: bad geometry");
    if (width_bytes == 0 || rows == 0) return PPO_OK;
    PPO_REQUIRE(dst && src, "ppo_memcpy_2d_async: null pointer");
    PPO_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(dst_pitch_bytes), src, static_cast<size_t>(src_pitch_bytes),
                               static_cast<size_t>(width_bytes), static_cast<size_t>(rows), cudaMemcpyDefault,
                               static_cast<cudaStream_t>(stream)));
    return PPO_OK;
}

// Peer-visible device memory for the mailboxes of ppo_gae_normalized_peers: a plain cudaMalloc allocation (zeroed) with
// its CUDA IPC handle; other processes of the node open the handle while THEIR device is current, which maps the memory
// for access from that device over NVLink (cudaIpcMemLazyEnablePeerAccess).
extern "C" int ppo_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
    PPO_REQUIRE(bytes > 0 && ptr && handle64, "ppo_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    PPO_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes)));
    PPO_CUDA(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
    PPO_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    PPO_CUDA(cudaIpcGetMemHandle(&h, *ptr));
    memcpy(handle64, &h, 64);
    return PPO_OK;
}

extern "C" int ppo_peer_open(const unsigned char* handle64, void** ptr) {
    PPO_REQUIRE(handle64 && ptr, "ppo_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PPO_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PPO_OK;
}

extern "C" int ppo_peer_close(void* ptr) {
    if (ptr) PPO_CUDA(cudaIpcCloseMemHandle(ptr));
    return PPO_OK;
}

extern "C" int ppo_peer_free(void* ptr) {
    if (ptr) PPO_CUDA(cudaFree(ptr));
    return PPO_OK;
}
