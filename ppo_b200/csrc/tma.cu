// Host side of tma.cuh: cuTensorMapEncodeTiled resolved through the runtime's driver entry point.
#include <mutex>

#include "tma.cuh"

namespace ppo {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        }
    });
    return fn;
}

int encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool swizzle_128b) {
    return encode_tensor_map_swz(map, dtype, rank, base, dims, strides_bytes, box, swizzle_128b ? 128 : 0);
}

int encode_tensor_map_swz(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base,
                          const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = resolve_encode();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return PPO_ERR_CUDA;
    }
    cuuint64_t gdims[5], gstrides[5];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstrides[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = fn(map, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstrides, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        return PPO_ERR_UNSUPPORTED;
    }
    return PPO_OK;
}

}  // namespace ppo
