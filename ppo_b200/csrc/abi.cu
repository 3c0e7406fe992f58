// Library-wide pieces of the C ABI: version, per-thread error text, device queries.
#include <stdarg.h>

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
