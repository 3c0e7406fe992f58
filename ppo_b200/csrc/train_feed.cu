// What consumes the targets once they stay in HBM: the minibatch row gather of the trainer's
// batch loop (rl/rollout.py:2349-2372, `data = var_value[sample]` + upload) and the TVF value loss with its gradient
// (rl/tvf.py:32-77, TVFRunnerModule.on_train_value_minibatch).  Both are pure HBM streams.
//
//   gather_rows_kernel      out[i, :] = src[index[i], :] for rows of any byte width; 16-byte chunks when every pitch
//                           allows it, a warp (or more) per row so that a row is one coalesced burst
//   tvf_value_loss_kernel   one warp per sample: loss[b] = sqrt(K) * mean_k(coef * 0.5 * (target - pred)^2 * w_k * mask)
//                           and d loss[b] / d pred[b, k] in the same pass over the two [B, K] arrays
#include <math.h>

#include "common.cuh"

namespace ppo {

struct GatherParams {
    const unsigned char* src;
    int64_t src_pitch, row_bytes, n_src;
    const int64_t* index;
    int64_t n_index;
    unsigned char* out;
    int64_t out_pitch;
    int* bad;  // set to 1 when an index is out of range (the row is zero-filled)
};

template <typename CHUNK>
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherParams p) {
    const int64_t chunks = p.row_bytes / static_cast<int64_t>(sizeof(CHUNK));
    // threads walk (row, chunk) pairs with the chunk index fastest: consecutive threads read consecutive bytes
    const int64_t total = p.n_index * chunks;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / chunks, c = i - row * chunks;
        int64_t s = __ldg(p.index + row);
        if (s < 0) s += p.n_src;  // NumPy semantics for negative indices
        CHUNK v{};
        if (s >= 0 && s < p.n_src) v = __ldg(reinterpret_cast<const CHUNK*>(p.src + s * p.src_pitch) + c);
        else if (c == 0) *p.bad = 1;
        __stcs(reinterpret_cast<CHUNK*>(p.out + row * p.out_pitch) + c, v);
    }
}

struct LossParams {
    const float* targets;
    int64_t t_rs;
    const int* head_index;
    const float* pred;
    int64_t p_rs, p_hs;
    const float* weights;
    const float* mask;
    int64_t rows, heads;
    float half_coef, root_k, inv_k;
    float* loss;
    float* grad;
};

__global__ void __launch_bounds__(256) tvf_value_loss_kernel(const LossParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t b = warp; b < p.rows; b += nwarps) {
        float acc = 0.f;
        for (int64_t k = lane; k < p.heads; k += 32) {
            const int64_t tk = p.head_index ? p.head_index[k] : k;
            const float t = __ldg(p.targets + b * p.t_rs + tk);
            const float q = __ldg(p.pred + b * p.p_rs + k * p.p_hs);
            float w = __ldg(p.weights + k);
            if (p.mask) w *= __ldg(p.mask + b * p.heads + k);
            const float d = t - q;
            // rl/tvf.py:49 `coef * 0.5 * square(targets - predictions)`, then the weights (:53-69)
            acc += (p.half_coef * (d * d)) * w;
            // d/dq [sqrt(K) / K * coef * 0.5 * (t - q)^2 * w] = -sqrt(K) / K * coef * (t - q) * w
            if (p.grad) __stcs(p.grad + b * p.heads + k, -(p.root_k * p.inv_k) * (2.f * p.half_coef) * d * w);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) p.loss[b] = p.root_k * (acc * p.inv_k);  // :73 `K ** 0.5 * mean(dim=-1)`
    }
}

// out[b, k] = g[b, k] * row_scale[b]: the chain rule step of the loss (d total / d pred = d total / d loss[b] * grad[b, k])
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ g, const float* __restrict__ row_scale,
                                                         int64_t rows, int64_t width, float* __restrict__ out) {
    const int64_t total = rows * width;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = __ldg(g + i) * __ldg(row_scale + i / width);
}

}  // namespace ppo

using namespace ppo;

extern "C" int ppo_scale_rows(const float* grad, const float* row_scale, int64_t rows, int64_t width, float* out,
                              void* stream) {
    PPO_REQUIRE(rows >= 0 && width >= 0, "ppo_scale_rows: negative shape");
    if (rows == 0 || width == 0) return PPO_OK;
    PPO_REQUIRE(grad && row_scale && out, "ppo_scale_rows: null pointer");
    int64_t blocks = (rows * width + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    scale_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, row_scale, rows,
                                                                                                  width, out);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_gather_rows(const void* src, int64_t src_pitch_bytes, int64_t row_bytes, int64_t n_src_rows,
                               const int64_t* index, int64_t n_index, void* out, int64_t out_pitch_bytes,
                               int* out_of_range, void* stream) {
    PPO_REQUIRE(row_bytes >= 0 && n_index >= 0 && n_src_rows >= 0, "ppo_gather_rows: negative size");
    PPO_REQUIRE(src_pitch_bytes >= row_bytes && out_pitch_bytes >= row_bytes, "ppo_gather_rows: pitch smaller than a row");
    if (n_index == 0 || row_bytes == 0) return PPO_OK;
    PPO_REQUIRE(src && index && out && out_of_range, "ppo_gather_rows: null pointer");
    GatherParams p{static_cast<const unsigned char*>(src), src_pitch_bytes, row_bytes, n_src_rows, index, n_index,
                   static_cast<unsigned char*>(out), out_pitch_bytes, out_of_range};
    auto aligned = [&](int64_t a) {
        return (reinterpret_cast<uintptr_t>(src) % a == 0) && (reinterpret_cast<uintptr_t>(out) % a == 0) &&
               src_pitch_bytes % a == 0 && out_pitch_bytes % a == 0 && row_bytes % a == 0;
    };
    const int64_t width = aligned(16) ? 16 : aligned(4) ? 4 : 1;
    int64_t blocks = (n_index * (row_bytes / width) + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (width == 16) gather_rows_kernel<uint4><<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
    else if (width == 4) gather_rows_kernel<uint32_t><<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
    else gather_rows_kernel<unsigned char><<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}

extern "C" int ppo_tvf_value_loss(const float* targets, int64_t target_row_stride, const int32_t* head_index,
                                  const float* predictions, int64_t pred_row_stride, int64_t pred_head_stride,
                                  const float* head_weights, const float* mask, int64_t rows, int64_t heads,
                                  double coef, float* loss_out, float* grad_out, void* stream) {
    PPO_REQUIRE(rows >= 0 && heads >= 1, "ppo_tvf_value_loss: bad shape [%lld, %lld]", (long long)rows, (long long)heads);
    if (rows == 0) return PPO_OK;
    PPO_REQUIRE(targets && predictions && head_weights && loss_out, "ppo_tvf_value_loss: null pointer");
    LossParams p{};
    p.targets = targets, p.t_rs = target_row_stride, p.head_index = head_index;
    p.pred = predictions, p.p_rs = pred_row_stride, p.p_hs = pred_head_stride;
    p.weights = head_weights, p.mask = mask, p.rows = rows, p.heads = heads;
    p.half_coef = static_cast<float>(coef * 0.5);
    p.root_k = static_cast<float>(sqrt(static_cast<double>(heads)));
    p.inv_k = 1.0f / static_cast<float>(heads);
    p.loss = loss_out, p.grad = grad_out;
    int64_t blocks = (rows + 7) / 8;  // eight warps per CTA, one sample per warp
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (blocks > cap) blocks = cap;
    tvf_value_loss_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PPO_CUDA(cudaGetLastError());
    return PPO_OK;
}
