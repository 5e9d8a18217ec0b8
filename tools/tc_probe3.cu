// tcgen05 probe 3: how fast do small MMAs go?  reps x (M128 x N x K8 tf32) issued by NI threads (one per warp, own accumulator),
// A from tensor memory (ts) or shared memory (ss).  Prints cycles per MMA (per issuer and aggregate).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/tc_probe3.cu -o tools/_build/tc_probe3
#include <cstdio>
#include "common.cuh"
#include "tc05.cuh"
namespace b200rl { int set_error(const char*, ...) { return 1; } int check_launch(const char*) { return 0; } }
using namespace b200rl;

__global__ void probe3(int NI, int N, int reps, int ts, int same_acc, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* A = sm;                   // 128 x 64 k, K-major, LBO 128, SBO 2048  (32 KB)
    unsigned char* Bm = sm + 32768;          // N x 64 k                                 (<= 64 KB)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32768 + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (tid == 0) { for (int i = 0; i < 8; ++i) tc::mbar_init(&bar[i], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = *slot;
    const uint32_t idesc = tc::make_idesc_tf32(128, N, false, false);
    __syncthreads();
    if (warp < NI && lane == 0) {
        const uint32_t D = tmem + (same_acc ? 0 : warp * N), At = tmem + 256;        // A tile: columns 256..319
        long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const tc::SmemDesc db = tc::make_desc(Bm + s * 256, 128, 2048);
                const uint32_t Dd = same_acc >= 2 ? tmem + (s % same_acc) * N : D;      // one thread, round-robin over `same_acc` accumulators
                if (ts) tc::mma_tf32_ts(Dd, At + s * 8, db, idesc, 1u);
                else tc::mma_tf32(Dd, tc::make_desc(A + s * 256, 128, 2048), db, idesc, 1u);
            }
        tc::mma_commit(&bar[warp]);
        long long t1 = clock64();
        tc::mbar_wait(&bar[warp], 0);
        long long t2 = clock64();
        cycles[warp * 2] = t1 - t0; cycles[warp * 2 + 1] = t2 - t0;
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long* dC; cudaMalloc(&dC, 128);
    const int smem = 32768 + 65536 + 128;
    cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 32;
    for (int ts = 1; ts >= 0; --ts)
        for (int N : {16, 32, 64, 128})
            for (int NI : {1, 2, 4}) {
                if (NI * N > 256) continue;
                probe3<<<1, 128, smem>>>(NI, N, reps, ts, 0, dC);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                long long c[8]; cudaMemcpy(c, dC, 64, cudaMemcpyDeviceToHost);
                long long done = 0, issue = 0;
                for (int i = 0; i < NI; ++i) { if (c[2 * i + 1] > done) done = c[2 * i + 1]; if (c[2 * i] > issue) issue = c[2 * i]; }
                printf("A from %s, N=%3d, %d issuer(s): issue loop %.1f cycles/MMA per issuer, completion %.1f cycles per MMA aggregate\n",
                       ts ? "TMEM" : "smem", N, NI, (double)issue / (reps * 8), (double)done / (reps * 8 * NI));
            }
    for (int acc : {2, 4, 8}) {
        probe3<<<1, 128, smem>>>(1, 32, reps, 1, acc, dC);
        cudaDeviceSynchronize();
        long long c[8]; cudaMemcpy(c, dC, 64, cudaMemcpyDeviceToHost);
        printf("A from TMEM, N= 32, ONE issuer round-robin over %d accumulators: issue loop %.1f cycles/MMA, completion %.1f cycles per MMA\n",
               acc, (double)c[0] / (reps * 8), (double)c[1] / (reps * 8));
    }
    for (int NI : {2, 4}) {
        probe3<<<1, 32 * NI, smem>>>(NI, 32, reps, 1, 2, dC);      // NI issuers, each alternating over the SAME two accumulators (hazard across threads)
        cudaDeviceSynchronize();
        long long c[8]; cudaMemcpy(c, dC, 64, cudaMemcpyDeviceToHost);
        printf("A from TMEM, N= 32, %d issuers all alternating over the same 2 accumulators: completion %.1f cycles per MMA aggregate\n", NI, (double)c[1] / (reps * 8 * NI));
    }
    return 0;
}
