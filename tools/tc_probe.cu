// Standalone tcgen05 probe: one kind::tf32 MMA (M=128, N, K=8) on host-built smem images.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I<csrc> -I<include> tools/tc_probe.cu -o tools/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "common.cuh"
#include "tc05.cuh"
namespace b200rl { int set_error(const char*, ...) { return 1; } int check_launch(const char*) { return 0; } }
using namespace b200rl;

__global__ void probe_kernel(const unsigned char* imgA, int bytesA, const unsigned char* imgB, int bytesB,
                             uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB, uint32_t idesc, int N, uint32_t ltype,
                             float* out /*128 x N*/, uint32_t* info) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* A = sm; unsigned char* Bm = sm + 16384;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < bytesA; i += blockDim.x) A[i] = imgA[i];
    for (int i = tid; i < bytesB; i += blockDim.x) Bm[i] = imgB[i];
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        info[0] = tmem; info[1] = tc::smem_u32(A); info[2] = tc::smem_u32(Bm);
        uint64_t da = tc::make_desc(A, lboA, sboA) | ((uint64_t)ltype << 61);
        uint64_t db = tc::make_desc(Bm, lboB, sboB) | ((uint64_t)ltype << 61);
        tc::mma_tf32(tmem, da, db, idesc, 0u);
        tc::mma_commit(bar);
    }
    tc::mbar_wait(bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float r[32];
        tc::tmem_ld32(tc::tmem_addr(tmem, 32 * (warp & 3), c0), r);
        for (int i = 0; i < 32; ++i) out[(size_t)(32 * (warp & 3) + lane) * N + c0 + i] = r[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

static size_t off_mn(int mn, int k, uint32_t lbo, uint32_t sbo) { return (size_t)(mn % 4) * 4 + (size_t)(mn / 4) * sbo + (size_t)(k % 8) * 16 + (size_t)(k / 8) * lbo; }
static size_t off_k(int mn, int k, uint32_t lbo, uint32_t sbo) { return (size_t)(mn % 8) * 16 + (size_t)(mn / 8) * sbo + (size_t)(k % 4) * 4 + (size_t)(k / 4) * lbo; }

static size_t off_mn128(int mn, int k, uint32_t atom_stride, uint32_t kgrp_stride) {
    return (size_t)(mn / 32) * atom_stride + (size_t)(k / 8) * kgrp_stride + (size_t)(k % 8) * 128 + (size_t)((((mn % 32) / 4) ^ (k % 8))) * 16 + (size_t)(mn % 4) * 4;
}
int run_case(const char* name, bool a_mn, bool b_mn, int N, uint32_t lboA, uint32_t sboA, uint32_t lboB, uint32_t sboB,
             int img = -1, uint32_t ltype = 0, uint32_t i1A = 0, uint32_t i2A = 0, uint32_t i1B = 0, uint32_t i2B = 0) {
    const int M = 128, K = 8;
    std::vector<float> Ah(M * K), Bh(N * K);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ah[m * K + k] = (float)((m * 7 + k * 3) % 11 - 5);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bh[n * K + k] = (float)((n * 5 + k * 2) % 13 - 6);
    std::vector<unsigned char> ia(16384, 0), ib(32768, 0);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { size_t o = img == 1 ? off_mn(m, k, i1A, i2A) : img == 2 ? off_mn128(m, k, i1A, i2A) : a_mn ? off_mn(m, k, lboA, sboA) : off_k(m, k, lboA, sboA); if (o + 4 > ia.size()) { printf("%s: A image overflow\n", name); return 1; } memcpy(&ia[o], &Ah[m * K + k], 4); }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { size_t o = img == 1 ? off_mn(n, k, i1B, i2B) : img == 2 ? off_mn128(n, k, i1B, i2B) : b_mn ? off_mn(n, k, lboB, sboB) : off_k(n, k, lboB, sboB); if (o + 4 > ib.size()) { printf("%s: B image overflow\n", name); return 1; } memcpy(&ib[o], &Bh[n * K + k], 4); }
    unsigned char *dA, *dB; float* dO; uint32_t* dI;
    cudaMalloc(&dA, ia.size()); cudaMalloc(&dB, ib.size()); cudaMalloc(&dO, sizeof(float) * M * N); cudaMalloc(&dI, 16);
    cudaMemcpy(dA, ia.data(), ia.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, ib.data(), ib.size(), cudaMemcpyHostToDevice);
    cudaMemset(dO, 0xff, sizeof(float) * M * N);
    const uint32_t idesc = tc::make_idesc_tf32(M, N, a_mn, b_mn);
    const int smem = 16384 + 32768 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(dA, (int)ia.size(), dB, (int)ib.size(), lboA, sboA, lboB, sboB, idesc, N, ltype, dO, dI);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(M * N); uint32_t info[4];
    cudaMemcpy(O.data(), dO, sizeof(float) * M * N, cudaMemcpyDeviceToHost); cudaMemcpy(info, dI, 16, cudaMemcpyDeviceToHost);
    double maxerr = 0; int nbad = 0, nnz = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float ex = 0; for (int k = 0; k < K; ++k) ex += Ah[m * K + k] * Bh[n * K + k];
        float got = O[m * N + n]; if (got != 0) ++nnz;
        double d = fabs((double)got - ex); if (!(d <= 1e-3)) { ++nbad; } if (d > maxerr || d != d) maxerr = d;
    }
    printf("%-34s idesc=%08x tmem=%08x smemA=%05x: nbad=%d/%d nonzero=%d maxerr=%g  O[0][0..3]=%g %g %g %g  O[1][0]=%g O[5][7]=%g\n",
           name, idesc, info[0], info[1], nbad, M * N, nnz, maxerr, O[0], O[1], O[2], O[3], O[N], O[5 * N + 7]);
    // expected samples
    float e00 = 0, e01 = 0; for (int k = 0; k < K; ++k) { e00 += Ah[k] * Bh[k]; e01 += Ah[k] * Bh[K + k]; }
    printf("%-34s expected O[0][0]=%g O[0][1]=%g\n", "", e00, e01);
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dI);
    return nbad != 0;
}

int main() {
    int bad = 0;
    // K-major, canonical: core matrices along K adjacent (LBO=128), 8-row groups SBO=256
    bad += run_case("K-major A,B  N=64  lbo128 sbo256", false, false, 64, 128, 256, 128, 256);
    bad += run_case("K-major A,B  N=256 lbo128 sbo256", false, false, 256, 128, 256, 128, 256);
    // K-major with the roles of LBO/SBO swapped in the descriptor
    bad += run_case("K-major swapped lbo256 sbo128", false, false, 64, 256, 128, 256, 128);
    // MN-major as used by wgrad_tc: SBO = 128 between 4-wide MN groups, LBO = next 8 K rows
    bad += run_case("MN-major A,B N=256 lbo4096/8192 sbo128", true, true, 256, 4096, 128, 8192, 128);
    bad += run_case("MN-major A,B N=64", true, true, 64, 4096, 128, 2048, 128);
    // MN-major no swizzle, image (kgrp stride 4096, mn-group stride 128) but descriptor fields swapped
    bad += run_case("MN none, desc swapped lbo128 sbo4096", true, true, 64, 128, 4096, 128, 2048, 1, 0, 4096, 128, 2048, 128);
    // MN-major SWIZZLE_128B: atoms of 32 mn x 8 k (1024 B), atom stride 1024 along MN
    bad += run_case("MN sw128 lbo=atom(1024) sbo=kgrp", true, true, 64, 1024, 4096, 1024, 2048, 2, 2, 1024, 4096, 1024, 2048);
    bad += run_case("MN sw128 swapped lbo=kgrp sbo=atom", true, true, 64, 4096, 1024, 2048, 1024, 2, 2, 1024, 4096, 1024, 2048);
    bad += run_case("MN sw128 N=256", true, true, 256, 1024, 4096, 1024, 8192, 2, 2, 1024, 4096, 1024, 8192);
    printf("failures: %d\n", bad);
    return 0;
}
