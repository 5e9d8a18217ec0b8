// Phase timeline of lstm_fwd_fused_kernel (dev tool): cycles per timestep seen by compute thread 0, a recurrent issuer and an
// input-projection issuer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200RL_FWD3PROF \
//     -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/fwd3_prof.cu -o tools/_build/fwd3_prof
#include <cstdarg>
#include <cstdio>
#include "common.cuh"
namespace b200rl {
int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return 1; }
int check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e)); return 0; }
}
#include "fwd_fused.cu"
int main(int argc, char** argv) {
    const int B = 4096, T = 128, NH = 3, Fin = 4;
    const bool heads = argc < 2 || argv[1][0] != '0';
    const long long rows = (long long)B * T;
    float *u, *g, *hx0, *cx0, *W, *c, *h, *hp, *ho;
    cudaMalloc(&u, rows * Fin * 4); cudaMalloc(&g, rows * 256 * 4); cudaMalloc(&hx0, B * 64 * 4); cudaMalloc(&cx0, B * 64 * 4); cudaMalloc(&W, 4 * 256 * 64 * 4);
    cudaMalloc(&c, rows * 64 * 4); cudaMalloc(&h, rows * 64 * 4); cudaMalloc(&hp, rows * 2 * NH * 4); cudaMalloc(&ho, rows * NH * 4);
    cudaMemset(u, 0, rows * Fin * 4); cudaMemset(hx0, 0, B * 64 * 4); cudaMemset(cx0, 0, B * 64 * 4); cudaMemset(W, 0, 4 * 256 * 64 * 4);
    auto run = [&]() { return launch_fwd_fused(u, hx0, cx0, W, W, W, W + 16384, W, W, W, W, g, c, h, heads ? hp : nullptr, ho, B, T, Fin, NH, 148, 0); };
    for (int it = 0; it < 2; ++it) run();
    cudaDeviceSynchronize();
    unsigned long long z[16] = {0};
#ifdef B200RL_FWD3PROF
    cudaMemcpyToSymbol(g_fwd3prof, z, sizeof(z));
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    run();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
#ifdef B200RL_FWD3PROF
    cudaMemcpyFromSymbol(z, g_fwd3prof, sizeof(z));
    const double steps = (B / 32.0) * T;
    printf("fwd_fused (heads %d, sep %d) %.3f ms incl. head_combine; cycles/step compute thread 0: loop top %.0f | wait acc_done %.0f | tmem ld + arrive acc_free %.0f | send raw + swap barrier %.0f | 4 cells + operand %.0f | fence + arrive b_full %.0f | stores + heads %.0f\n",
           (int)heads, B200RL_FWD3_SEP, ms, z[0] / steps, z[1] / steps, z[2] / steps, z[3] / steps, z[4] / steps, z[5] / steps, z[6] / steps);
    printf("   recurrent issuer: loop top %.0f | wait b_full %.0f | wait x_done %.0f | issue 10 MMAs + commit %.0f\n", z[7] / steps, z[8] / steps, z[9] / steps, z[10] / steps);
    printf("   input-projection issuer: loop top %.0f | wait x_full %.0f | wait acc_free %.0f | issue 10 MMAs + commit %.0f\n", z[11] / steps, z[12] / steps, z[13] / steps, z[14] / steps);
#else
    printf("fwd_fused (no probes, heads %d, sep %d) %.3f ms incl. head_combine\n", (int)heads, B200RL_FWD3_SEP, ms);
#endif
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
