// tcgen05 probe 4: kind::f16 (bf16 inputs, fp32 accumulate) with the A operand in TENSOR MEMORY.  How are the two bf16 of a
// 32-bit TMEM column ordered along K?  M128 x N32 x K16, integer-valued data (exact), checks hypothesis "low half = even k".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/tc_probe4.cu -o tools/_build/tc_probe4
#include <cstdio>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "tc05.cuh"
namespace b200rl { int set_error(const char*, ...) { return 1; } int check_launch(const char*) { return 0; } }
using namespace b200rl;

__device__ __forceinline__ unsigned short f2bf(float f) { return (unsigned short)(__float_as_uint(f) >> 16); }   // exact for the small integers used here

__global__ void probe4(const float* A, const float* Bm, float* out, int N) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* Bs = sm;                                  // N x 16 bf16, K-major: (n%8)*16 + (n/8)*256 + (k%8)*2 + (k/8)*128
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < N * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        *reinterpret_cast<unsigned short*>(Bs + (n % 8) * 16 + (n / 8) * 256 + (k % 8) * 2 + (k / 8) * 128) = f2bf(Bm[n * 16 + k]);
    }
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 128);
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = *slot;
    {   // A row m = tid (128 threads): 16 bf16 -> 8 columns at 64.., column c = (k = 2c low half, k = 2c+1 high half)
        float v[16];
        for (int c = 0; c < 8; ++c) v[c] = __uint_as_float((unsigned)f2bf(A[tid * 16 + 2 * c]) | ((unsigned)f2bf(A[tid * 16 + 2 * c + 1]) << 16));
        for (int c = 8; c < 16; ++c) v[c] = 0.f;
        tc::tmem_st16(tc::tmem_addr(tmem, 32 * (warp & 3), 64), v);
    }
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 acc, BF16 x BF16, K-major
        const tc::SmemDesc db = tc::make_desc(Bs, 128, 256);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                     ::"r"(tmem), "r"(tmem + 64), "l"(db), "r"(idesc), "r"(0u) : "memory");
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, 0);
    tc::fence_after_sync();
    float r[32];
    tc::tmem_ld32(tc::tmem_addr(tmem, 32 * (warp & 3), 0), r);
    for (int i = 0; i < 32; ++i) out[(size_t)tid * N + i] = r[i];
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

int main() {
    const int M = 128, N = 32, K = 16;
    std::vector<float> A(M * K), Bh(N * K);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = (float)((m * 7 + k * 3) % 11 - 5);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bh[n * K + k] = (float)((n * 5 + k * k) % 13 - 6);
    float *dA, *dB, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bh.size() * 4); cudaMalloc(&dO, M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bh.data(), Bh.size() * 4, cudaMemcpyHostToDevice);
    probe4<<<1, 128, 8192>>>(dA, dB, dO, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> O(M * N);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    int bad1 = 0, bad2 = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double h1 = 0, h2 = 0;
        for (int k = 0; k < K; ++k) { h1 += (double)A[m * K + k] * Bh[n * K + k]; h2 += (double)A[m * K + (k ^ 1)] * Bh[n * K + k]; }
        bad1 += O[m * N + n] != (float)h1; bad2 += O[m * N + n] != (float)h2;
    }
    printf("kind::f16 bf16, A from TMEM: mismatches vs 'low half = even k': %d ; vs 'low half = odd k': %d ; O[0][0..3] = %g %g %g %g\n",
           bad1, bad2, O[0], O[1], O[2], O[3]);
    return 0;
}
