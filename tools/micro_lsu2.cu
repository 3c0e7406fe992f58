// Development micro-benchmark for the column-slab TVF kernel (tvf_cslab.cu): what does one (horizon, n-step) entry cost
// when a thread owns 2 x 4 consecutive time rows and gathers with LDS.128?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/micro_lsu2 tools/micro_lsu2.cu
// One CTA of 16 warps per SM; every warp issues `iters` rounds; SM cycles per warp-round are reported (divide the
// CTA's cycles by 16 warps: with the LSU pipe as the limit this is the wavefront count of a round).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__constant__ uint4 ctab[1024];

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
__device__ __forceinline__ void fma_if(float& acc, float lo, float hi, float wlo, float whi, uint32_t w, uint32_t key) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %5, %6;\n\t"
        "@p fma.rn.f32 %0, %1, %3, %0;\n\t"
        "@p fma.rn.f32 %0, %2, %4, %0;\n\t}"
        : "+f"(acc)
        : "f"(lo), "f"(hi), "f"(wlo), "f"(whi), "r"(w), "r"(key));
}

// MODE 0: 4 x LDS.128 gathers only (sum)            1: constants from shared memory (LDS.32 + LDS.64) + 4 gathers + 8 setp + 16 fma
//      2: like 1, constants computed (no loads)     3: like 1, constants from the constant bank, fetched one entry ahead
//      4: like 1 plus 6 SHFL every 7th round        5: like 1 but only every second round loads (shared aligned group)
//      6: LDS.128 constants (16-byte entry)          7: 8 x LDS.32 gathers (two rows per thread, the round-1 pattern) per half round
template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, float* out, long long* cyc, int n_ent) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    // SoA constants: words [n_ent], weights [n_ent] float2, then 16-byte copies, then the slab
    for (int e = threadIdx.x; e < n_ent; e += blockDim.x) {
        const uint32_t off = ((e * 37) % 2000) * 16, n = 1 + (e % 60);
        w[e] = off | (n << 18);
        w[1024 + 2 * e] = __float_as_uint(1.0f / (1 + e)), w[1024 + 2 * e + 1] = __float_as_uint(0.5f / (2 + e));
        w[4096 + 4 * e] = off | (n << 18), w[4096 + 4 * e + 1] = w[1024 + 2 * e], w[4096 + 4 * e + 2] = w[1024 + 2 * e + 1];
    }
    const uint32_t slab_off = 8192 * 4;
    for (int i = threadIdx.x; i < 40000; i += blockDim.x) w[slab_off / 4 + i] = __float_as_uint(i * 1e-3f);
    __syncthreads();
    const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t base = sa + slab_off + lane * 16;
    float acc[2][7];
    uint32_t key[2][7];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[g][j] = 0.f, key[g][j] = static_cast<uint32_t>(20 + ((lane * 7 + j * 3 + g) % 50)) << 18;
    uint32_t keep = 0;
    int e = (warp * 53) % n_ent;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (MODE == 3) nxt = ctab[e];
    uint4 lo0 = make_uint4(0, 0, 0, 0), hi0 = lo0, lo1 = lo0, hi1 = lo0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t wd;
        float wlo, whi;
        if (MODE == 1 || MODE == 4 || MODE == 5 || MODE == 7) {
            wd = lds32u(sa + e * 4);
            const uint2 c = lds64(sa + 4096 + e * 8);
            wlo = __uint_as_float(c.x), whi = __uint_as_float(c.y);
        } else if (MODE == 6) {
            const uint4 c = lds128(sa + 16384 + e * 16);
            wd = c.x, wlo = __uint_as_float(c.y), whi = __uint_as_float(c.z);
        } else if (MODE == 3) {
            wd = nxt.x, wlo = __uint_as_float(nxt.y), whi = __uint_as_float(nxt.z);
            const int en = e + 1 == n_ent ? 0 : e + 1;
            const uint4* cp = ctab + en;
            asm volatile("ld.const.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(nxt.x), "=r"(nxt.y), "=r"(nxt.z), "=r"(nxt.w) : "l"(cp));
        } else {
            wd = (((e * 37) % 2000) * 16) | ((1 + (e % 60)) << 18);
            wlo = 0.5f, whi = 0.25f;
        }
        const uint32_t a = base + (wd & 0x3fff0u);
        if (MODE == 7) {
            // the round-1 pattern: scalar gathers, two rows per thread and step, four steps
            float s = 0.f;
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                float l0, h0, l1, h1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l0) : "r"(a + st * 256));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(h0) : "r"(a + st * 256 + 1028));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l1) : "r"(a + st * 256 + 128));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(h1) : "r"(a + st * 256 + 1028 + 128));
                fma_if(acc[0][st], l0, h0, wlo, whi, wd, key[0][st]);
                fma_if(acc[1][st], l1, h1, wlo, whi, wd, key[1][st]);
            }
            keep += __float_as_uint(s);
        } else {
            if (MODE != 5 || (it & 1) == 0) {
                lo0 = lds128(a), hi0 = lds128(a + 1024), lo1 = lds128(a + 512), hi1 = lds128(a + 1024 + 512);
            }
            if (MODE == 0) {
                keep += lo0.x ^ hi0.y ^ lo1.z ^ hi1.w;
            } else {
                fma_if(acc[0][1], __uint_as_float(lo0.x), __uint_as_float(hi0.x), wlo, whi, wd, key[0][1]);
                fma_if(acc[0][2], __uint_as_float(lo0.y), __uint_as_float(hi0.y), wlo, whi, wd, key[0][2]);
                fma_if(acc[0][3], __uint_as_float(lo0.z), __uint_as_float(hi0.z), wlo, whi, wd, key[0][3]);
                fma_if(acc[0][4], __uint_as_float(lo0.w), __uint_as_float(hi0.w), wlo, whi, wd, key[0][4]);
                fma_if(acc[1][1], __uint_as_float(lo1.x), __uint_as_float(hi1.x), wlo, whi, wd, key[1][1]);
                fma_if(acc[1][2], __uint_as_float(lo1.y), __uint_as_float(hi1.y), wlo, whi, wd, key[1][2]);
                fma_if(acc[1][3], __uint_as_float(lo1.z), __uint_as_float(hi1.z), wlo, whi, wd, key[1][3]);
                fma_if(acc[1][4], __uint_as_float(lo1.w), __uint_as_float(hi1.w), wlo, whi, wd, key[1][4]);
            }
        }
        if (MODE == 4 && it % 7 == 6) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[g][j] += __shfl_sync(0xffffffffu, acc[g][4 + j], (lane + 31) & 31);
        }
        if (++e == n_ent) e = 0;
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int j = 0; j < 7; ++j) s += acc[g][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + keep;
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
    printf("%-64s %8.2f cycles per warp-round   (%s)\n", name, (double)c / iters / 16.0, cudaGetErrorString(rc));
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 8);
    uint4* h = (uint4*)malloc(1024 * 16);
    for (int e = 0; e < 1024; ++e) h[e] = make_uint4((((e * 37) % 2000) * 16) | ((1 + e % 60) << 18), 0x3f000000u, 0x3e800000u, 0);
    cudaMemcpyToSymbol(ctab, h, 1024 * 16);
    const int iters = 20000, n_ent = 838;
    run<0>("4 x LDS.128 gathers (8 rows x lo/hi)", iters, out, cyc, n_ent);
    run<2>("4 x LDS.128 + 8 setp + 16 fma, computed constants", iters, out, cyc, n_ent);
    run<1>("LDS.32 + LDS.64 constants + 4 x LDS.128 + 8 setp + 16 fma", iters, out, cyc, n_ent);
    run<6>("LDS.128 constants + 4 x LDS.128 + 8 setp + 16 fma", iters, out, cyc, n_ent);
    run<3>("ld.const.v4 one ahead + 4 x LDS.128 + 8 setp + 16 fma", iters, out, cyc, n_ent);
    run<4>("mode 1 + 6 SHFL every 7th round", iters, out, cyc, n_ent);
    run<5>("mode 1, gathers every second round (shared aligned group)", iters, out, cyc, n_ent);
    run<7>("round-1 pattern: constants + 16 x LDS.32 + 8 setp + 16 fma", iters, out, cyc, n_ent);
    return 0;
}
