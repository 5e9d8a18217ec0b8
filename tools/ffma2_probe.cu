// FFMA vs FFMA2 (fma.rn.f32x2) throughput probe on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a = threadIdx.x * 1e-3f + 1.0f, b = 0.999f;
    long long t0 = clock64();
    if (MODE == 0) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], b, a);
        }
        float s = 0; for (int i = 0; i < 32; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        unsigned long long acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = pack2((float)i, (float)i + 0.5f);
        const unsigned long long bb = pack2(b, b), aa = pack2(a, a);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { unsigned long long d = aa; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(acc[i]), "l"(bb), "l"(aa)); acc[i] = d; }
        }
        unsigned long long s = 0; for (int i = 0; i < 16; ++i) s ^= acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int threads : {128, 256, 512, 1024}) {
        long long c0, c1;
        k<0><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        k<1><<<148, threads>>>(out, iters, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c1, cyc, 8, cudaMemcpyDeviceToHost);
        double fma_per_clk_sm0 = (double)threads * 32 * iters / c0, fma_per_clk_sm1 = (double)threads * 32 * iters / c1;
        printf("threads/SM=%4d: FFMA %lld cyc -> %.1f FMA/clk/SM | FFMA2 %lld cyc -> %.1f FMA/clk/SM\n", threads, c0, fma_per_clk_sm0, c1, fma_per_clk_sm1);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
