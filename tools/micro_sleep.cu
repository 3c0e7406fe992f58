// Development micro-benchmark: how long does __nanosleep(t) really suspend a warp on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned ns, int iters, long long* out, int mode) {
    const long long t0 = clock64();
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    if (mode == 0 || threadIdx.x == 0)
        for (int i = 0; i < iters; ++i) __nanosleep(ns);
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0, out[1] = (long long)(g1 - g0);
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    long long h[2];
    for (int mode = 0; mode < 2; ++mode)
        for (unsigned ns : {20u, 100u, 200u, 400u, 800u, 2000u, 8000u}) {
            k<<<1, 32>>>(ns, 1000, d, mode); cudaDeviceSynchronize();
            k<<<1, 32>>>(ns, 1000, d, mode); cudaDeviceSynchronize();
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mode %d (%s) nanosleep(%u) x1000: %.1f clk, %.1f ns per call\n", mode, mode ? "lane 0 only" : "all lanes", ns, h[0] / 1000.0, h[1] / 1000.0);
        }
    return 0;
}
