// Phase timeline of lstm_dx_tc_kernel (dev tool): cycles thread 0 spends between the instrumented points,
// averaged per tile.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200RL_DXPROF \
//   -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/dx_prof.cu -o gpurun_out/dx_prof
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "common.cuh"
namespace b200rl {
int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return 1; }
int check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e)); return 0; }
}
#include "dx_tc.cu"
using namespace b200rl;
int main(int argc, char** argv) {
    const int B = 4096, T = 128, Fin = 4, NH = 3, nparts = 148;
    const long long rows = (long long)B * T, P = 40000;
    float *dpre, *u, *W_ih, *Wt, *bt, *h, *dhead, *partials, *du;
    cudaMalloc(&dpre, rows * 256 * 4); cudaMalloc(&u, rows * Fin * 4); cudaMalloc(&W_ih, 256 * 64 * 4);
    cudaMalloc(&Wt, 64 * Fin * 4); cudaMalloc(&bt, 64 * 4); cudaMalloc(&h, rows * 64 * 4); cudaMalloc(&dhead, rows * NH * 4);
    cudaMalloc(&partials, nparts * P * 4); cudaMalloc(&du, rows * Fin * 4);
    cudaMemset(dpre, 0, rows * 256 * 4); cudaMemset(u, 0, rows * Fin * 4); cudaMemset(W_ih, 0, 256 * 64 * 4);
    cudaMemset(Wt, 0, 64 * Fin * 4); cudaMemset(bt, 0, 64 * 4); cudaMemset(h, 0, rows * 64 * 4); cudaMemset(dhead, 0, rows * NH * 4);
    for (int with_du = 0; with_du < 2; ++with_du) {
        for (int it = 0; it < 3; ++it) launch_dx_tc(dpre, u, W_ih, Wt, bt, h, dhead, partials, nparts, P, 0, 256, 1000, 2000, with_du ? du : nullptr, B, T, Fin, NH, 0);
        cudaDeviceSynchronize();
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_dxprof, z, sizeof(z));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        launch_dx_tc(dpre, u, W_ih, Wt, bt, h, dhead, partials, nparts, P, 0, 256, 1000, 2000, with_du ? du : nullptr, B, T, Fin, NH, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpyFromSymbol(z, g_dxprof, sizeof(z));
        const double ntiles = rows / 128.0;
        printf("du=%d  dx_tc+head_grad %.3f ms; cycles/tile: fill %.0f | mma tail %.0f | u load %.0f | mask->S %.0f | bar %.0f | reduce %.0f | du %.0f | bar %.0f || per tile: load wait %.0f | empty wait %.0f | split+store+arrive %.0f\n",
               with_du, ms, z[0] / ntiles, z[1] / ntiles, z[2] / ntiles, z[3] / ntiles, z[4] / ntiles, z[5] / ntiles, z[6] / ntiles, z[7] / ntiles, z[8] / ntiles, z[9] / ntiles, z[10] / ntiles);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
