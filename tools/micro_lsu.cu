// Development micro-benchmark: what do warp-uniform constants cost next to shared-memory gathers?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/micro_lsu tools/micro_lsu.cu
// One CTA of 16 warps per SM; every test issues `iters` rounds of a fixed instruction mix per warp and reports
// SM cycles per round (clock64 on one SM, all SMs run the same thing).  The gather is the tslab pattern: 32
// consecutive words at a warp-uniform offset (one wavefront).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__constant__ uint4 ctab[1024];

__device__ __forceinline__ float lds(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// mode: 0 gathers only (4 per round)            1 LDS.128 broadcast + 4 gathers (the kernel today)
//       2 LDS.64 broadcast + 4 gathers          3 LDS.32 broadcast + 4 gathers
//       4 3 x SHFL + 4 gathers                  5 3 x SHFL only
//       6 LDS.128 broadcast only                7 ld.const.v4 (per-warp walk of a 13 KB table) + 4 gathers
//       8 ld.const.v4 only                      9 LDS.64 broadcast only
//      10 two LDS.64 broadcasts + 4 gathers    11 LDS.128 broadcast + 8 gathers (four row sets)
//      12 LDS.128 with two addresses per warp (half-warps) only
template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, float* out, long long* cyc, int n_ent) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    // entries: [n_ent] x 16 B at the start, then the slab
    for (int i = threadIdx.x; i < n_ent * 4; i += blockDim.x) {
        const int e = i >> 2, f = i & 3;
        uint32_t v;
        if (f == 0) v = ((e * 37) % 2000) * 4;  // gather offset in bytes
        else if (f == 3) v = e & 63;
        else v = __float_as_uint(1.0f / (1 + e + f));
        w[i] = v;
    }
    const uint32_t slab_off = n_ent * 16;
    for (int i = threadIdx.x; i < 40000; i += blockDim.x) w[slab_off / 4 + i] = __float_as_uint(i * 1e-3f);
    __syncthreads();
    const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t base = sa + slab_off + lane * 4;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    uint32_t keep = 0;
    int e = (warp * 53) % n_ent;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t off = ((e * 37) % 2000) * 4;
        float wlo = 0.5f, whi = 0.25f;
        if (MODE == 1 || MODE == 6 || MODE == 11) {
            const uint4 c = lds128(sa + e * 16);
            off = c.x, wlo = __uint_as_float(c.y), whi = __uint_as_float(c.z), keep += c.w;
        } else if (MODE == 12) {
            const uint4 c = lds128(sa + ((e + (lane >> 4) * 7) % n_ent) * 16);
            off = c.x, wlo = __uint_as_float(c.y), whi = __uint_as_float(c.z), keep += c.w;
        } else if (MODE == 2 || MODE == 9) {
            const uint2 c = lds64(sa + e * 16);
            off = c.x, wlo = __uint_as_float(c.y);
        } else if (MODE == 10) {
            const uint2 c = lds64(sa + e * 16);
            const uint2 d = lds64(sa + e * 16 + 8);
            off = c.x, wlo = __uint_as_float(c.y), whi = __uint_as_float(d.x), keep += d.y;
        } else if (MODE == 3) {
            off = lds32u(sa + e * 16);
        } else if (MODE == 4 || MODE == 5) {
            // lane j of the warp holds "its" entry; broadcast three words from lane e & 31
            const uint32_t mine_off = ((lane * 37 + it) % 2000) * 4;
            const float mine_lo = 1.0f / (1 + lane), mine_hi = 0.5f / (1 + lane);
            off = __shfl_sync(0xffffffffu, mine_off, e & 31);
            wlo = __shfl_sync(0xffffffffu, mine_lo, e & 31);
            whi = __shfl_sync(0xffffffffu, mine_hi, e & 31);
        } else if (MODE == 7 || MODE == 8) {
            uint4 c;
            const uint4* cp = ctab + e;
            asm volatile("ld.const.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(cp));
            off = c.x, wlo = __uint_as_float(c.y), whi = __uint_as_float(c.z), keep += c.w;
        }
        if (MODE != 5 && MODE != 6 && MODE != 8 && MODE != 9 && MODE != 12) {
            const uint32_t a = base + off;
            const float l0 = lds(a), h0 = lds(a + 257 * 4), l1 = lds(a + 128), h1 = lds(a + 257 * 4 + 128);
            acc0 = fmaf(l0, wlo, acc0), acc0 = fmaf(h0, whi, acc0);
            acc1 = fmaf(l1, wlo, acc1), acc1 = fmaf(h1, whi, acc1);
            if (MODE == 11) {
                const float l2 = lds(a + 256), h2 = lds(a + 257 * 4 + 256), l3 = lds(a + 384), h3 = lds(a + 257 * 4 + 384);
                acc2 = fmaf(l2, wlo, acc2), acc2 = fmaf(h2, whi, acc2);
                acc3 = fmaf(l3, wlo, acc3), acc3 = fmaf(h3, whi, acc3);
            }
        } else {
            acc0 += wlo, acc1 += whi, keep += off;
        }
        if (++e == n_ent) e = 0;
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3 + keep;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
static void run(const char* name, int iters, float* out, long long* cyc, int n_ent) {
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<MODE><<<148, 512, smem>>>(iters, out, cyc, n_ent);
    bench<MODE><<<148, 512, smem>>>(iters, out, cyc, n_ent);
    cudaError_t rc = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    // 16 warps each issue `iters` rounds: SM cycles per warp-round
    printf("%-48s %8.2f cycles per warp-round   (%s)\n", name, (double)c / iters / 16.0, cudaGetErrorString(rc));
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 8);
    uint4* h = (uint4*)malloc(1024 * 16);
    for (int e = 0; e < 1024; ++e) h[e] = make_uint4(((e * 37) % 2000) * 4, 0x3f000000u, 0x3e800000u, e & 63);
    cudaMemcpyToSymbol(ctab, h, 1024 * 16);
    const int iters = 20000, n_ent = 838;
    run<0>("4 gathers", iters, out, cyc, n_ent);
    run<1>("LDS.128 bcast + 4 gathers (today)", iters, out, cyc, n_ent);
    run<2>("LDS.64 bcast + 4 gathers", iters, out, cyc, n_ent);
    run<3>("LDS.32 bcast + 4 gathers", iters, out, cyc, n_ent);
    run<10>("2 x LDS.64 bcast + 4 gathers", iters, out, cyc, n_ent);
    run<4>("3 x SHFL + 4 gathers", iters, out, cyc, n_ent);
    run<5>("3 x SHFL only", iters, out, cyc, n_ent);
    run<6>("LDS.128 bcast only", iters, out, cyc, n_ent);
    run<9>("LDS.64 bcast only", iters, out, cyc, n_ent);
    run<12>("LDS.128, two addresses per warp, only", iters, out, cyc, n_ent);
    run<7>("ld.const.v4 + 4 gathers", iters, out, cyc, n_ent);
    run<8>("ld.const.v4 only", iters, out, cyc, n_ent);
    run<11>("LDS.128 bcast + 8 gathers (four row sets)", iters, out, cyc, n_ent);
    return 0;
}
