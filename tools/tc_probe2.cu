// tcgen05 probe 2: K-major SWIZZLE_128B correctness (K=32, 4 K-steps) and MMA timing vs SWIZZLE_NONE.
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include "common.cuh"
#include "tc05.cuh"
namespace b200rl { int set_error(const char*, ...) { return 1; } int check_launch(const char*) { return 0; } }
using namespace b200rl;

// mode 0: no swizzle (LBO=128 between K core matrices, SBO = (K/4)*128) ; mode 1: SW128 atoms (8 rows x 128 B), SBO=1024
__global__ void probe2(const unsigned char* imgA, const unsigned char* imgB, int bytesA, int bytesB, int mode, int N, int reps,
                       float* out, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* A = sm; unsigned char* Bm = sm + 32768;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32768 + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < bytesA; i += blockDim.x) A[i] = imgA[i];
    for (int i = tid; i < bytesB; i += blockDim.x) Bm[i] = imgB[i];
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 256);
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = *slot;
    const uint32_t idesc = tc::make_idesc_tf32(128, N, false, false);
    if (tid == 0) {
        long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
            for (int s = 0; s < 4; ++s) {
                uint64_t da, db;
                if (mode == 0) {
                    da = tc::make_desc(A + s * 256, 128, 1024);
                    db = tc::make_desc(Bm + s * 256, 128, 1024);
                } else {
                    da = tc::make_desc(A + s * 32, 16, 1024) | ((uint64_t)2 << 61);
                    db = tc::make_desc(Bm + s * 32, 16, 1024) | ((uint64_t)2 << 61);
                }
                tc::mma_tf32(tmem, da, db, idesc, (rep > 0 || s > 0) ? 1u : 0u);
            }
        tc::mma_commit(bar);
        long long t1 = clock64();
        cycles[0] = t1 - t0;
    }
    long long t2 = clock64();
    tc::mbar_wait(bar, 0);
    long long t3 = clock64();
    if (tid == 0) cycles[1] = t3 - t2;
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float r[32];
        tc::tmem_ld32(tc::tmem_addr(tmem, 32 * (warp & 3), c0), r);
        for (int i = 0; i < 32; ++i) out[(size_t)(32 * (warp & 3) + lane) * N + c0 + i] = r[i];
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

static size_t off0(int mn, int k) { return (size_t)(mn % 8) * 16 + (size_t)(mn / 8) * 1024 + (size_t)(k % 4) * 4 + (size_t)(k / 4) * 128; }
static size_t off1(int mn, int k) { return (size_t)(mn / 8) * 1024 + (size_t)(mn % 8) * 128 + (size_t)(((k / 4) ^ (mn % 8))) * 16 + (size_t)(k % 4) * 4; }

int run(int mode, int N, int reps) {
    const int M = 128, K = 32;
    std::vector<float> Ah(M * K), Bh(N * K);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ah[m * K + k] = (float)((m * 7 + k * 3) % 11 - 5);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bh[n * K + k] = (float)((n * 5 + k * 2) % 13 - 6);
    std::vector<unsigned char> ia(32768, 0), ib(65536, 0);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { size_t o = mode ? off1(m, k) : off0(m, k); memcpy(&ia[o], &Ah[m * K + k], 4); }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { size_t o = mode ? off1(n, k) : off0(n, k); memcpy(&ib[o], &Bh[n * K + k], 4); }
    unsigned char *dA, *dB; float* dO; long long* dC;
    cudaMalloc(&dA, ia.size()); cudaMalloc(&dB, ib.size()); cudaMalloc(&dO, sizeof(float) * M * N); cudaMalloc(&dC, 16);
    cudaMemcpy(dA, ia.data(), ia.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, ib.data(), ib.size(), cudaMemcpyHostToDevice);
    const int smem = 32768 + 65536 + 64;
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe2<<<1, 128, smem>>>(dA, dB, (int)ia.size(), (int)ib.size(), mode, N, reps, dO, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d N %d: CUDA error %s\n", mode, N, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(M * N); long long cyc[2];
    cudaMemcpy(O.data(), dO, sizeof(float) * M * N, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
    int nbad = 0; double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float ex = 0; for (int k = 0; k < K; ++k) ex += Ah[m * K + k] * Bh[n * K + k];
        double d = fabs((double)O[m * N + n] - (double)ex * reps); if (!(d <= 1e-2 * reps)) ++nbad; if (d > maxerr) maxerr = d;
    }
    printf("mode=%s N=%3d reps=%3d (%4d MMAs): nbad=%d maxerr=%g | issue %lld cyc, wait-after-issue %lld cyc -> %.1f cyc/MMA\n",
           mode ? "SW128 " : "NOSWZ ", N, reps, reps * 4, nbad, maxerr, cyc[0], cyc[1], (double)(cyc[0] + cyc[1]) / (reps * 4));
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
    return nbad != 0;
}

int main() {
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {64, 256}) { run(mode, N, 1); run(mode, N, 64); }
    return 0;
}
