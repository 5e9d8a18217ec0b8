// Where the time of lstm_wgrad_tc_kernel goes (dev tool): the kernel compiled with parts switched off.
//   for m in 0 1 2 3; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200RL_WGPROBE=$m \
//     -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/wgrad_prof.cu -o tools/_build/wgrad_prof$m; done
//   0: the kernel as shipped ; 1: HBM loads only ; 2: loads + split + shared-memory stores, no MMAs ; 3: no HBM loads
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
namespace b200rl {
int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return 1; }
int check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e)); return 0; }
}
#include "wgrad_tc.cu"
using namespace b200rl;
int main(int argc, char** argv) {
    const int B = 4096, T = 128, Fin = 4;
    const int nparts = argc > 1 ? atoi(argv[1]) : 296;
    const long long rows = (long long)B * T, P = 33796;
    float *dpre, *h, *hx0, *u, *Wt, *part;
    cudaMalloc(&dpre, rows * 256 * 4); cudaMalloc(&h, rows * 64 * 4); cudaMalloc(&hx0, B * 64 * 4); cudaMalloc(&u, rows * Fin * 4);
    cudaMalloc(&Wt, 64 * (Fin + 1) * 4); cudaMalloc(&part, nparts * P * 4);
    cudaMemset(dpre, 0, rows * 256 * 4); cudaMemset(h, 0, rows * 64 * 4); cudaMemset(hx0, 0, B * 64 * 4); cudaMemset(u, 0, rows * Fin * 4);
    cudaMemset(Wt, 0, 64 * (Fin + 1) * 4);
    auto run = [&]() { return launch_wgrad_tc(dpre, h, hx0, u, Wt, Wt + 64 * Fin, part, nparts, P, 256, 256 + 16384, 33000, 33256, B, T, Fin, 0); };
    for (int it = 0; it < 3; ++it) run();
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) run();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lstm_wgrad_tc_kernel<false>, wtc::NT, wgrad_tc_smem_bytes(Fin));
    printf("CTAs per SM: %d  ", occ);
    printf("wgrad_tc probe %d (bf16 cross terms %d), %d CTAs: %.3f ms  (%.0f GB/s of dpre + h + u)  %s\n", B200RL_WGPROBE, B200RL_WG_BF16X, nparts, ms / 10,
           (double)rows * (256 + 64 + Fin) * 4 / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
