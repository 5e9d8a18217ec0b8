// Phase timeline of lstm_fwdt_kernel (dev tool): cycles per timestep seen by compute thread 0 and by the MMA issuer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200RL_FWDPROF \
//     -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/fwd_prof.cu -o tools/_build/fwd_prof
#include <cstdarg>
#include <cstdio>
#include "common.cuh"
namespace b200rl {
int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return 1; }
int check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e)); return 0; }
}
#include "fwd_tc.cu"
using namespace b200rl;
int main() {
    const int B = 4096, T = 128;
    const long long rows = (long long)B * T;
    float *xg, *hx0, *cx0, *W, *c, *h;
    cudaMalloc(&xg, rows * 256 * 4); cudaMalloc(&hx0, B * 64 * 4); cudaMalloc(&cx0, B * 64 * 4); cudaMalloc(&W, 256 * 64 * 4);
    cudaMalloc(&c, rows * 64 * 4); cudaMalloc(&h, rows * 64 * 4);
    cudaMemset(xg, 0, rows * 256 * 4); cudaMemset(hx0, 0, B * 64 * 4); cudaMemset(cx0, 0, B * 64 * 4); cudaMemset(W, 0, 256 * 64 * 4);
    for (int it = 0; it < 2; ++it) launch_fwd_tc(xg, hx0, cx0, W, c, h, B, T, 148, 0);
    cudaDeviceSynchronize();
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_fwdprof, z, sizeof(z));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch_fwd_tc(xg, hx0, cx0, W, c, h, B, T, 148, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpyFromSymbol(z, g_fwdprof, sizeof(z));
    const double steps = (B / 32.0) * T;
    printf("fwd_tc %.3f ms; cycles/step compute thread 0: loop top %.0f | wait acc_done %.0f | tmem ld %.0f | send raw %.0f | swap barrier %.0f | 8 cells %.0f | fence+arrive %.0f\n",
           ms, z[0] / steps, z[1] / steps, z[2] / steps, z[3] / steps, z[4] / steps, z[5] / steps, z[6] / steps);
    printf("   issuer: loop top %.0f | wait b_full %.0f | issue 48 MMAs + commit %.0f\n", z[8] / steps, z[9] / steps, z[10] / steps);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
