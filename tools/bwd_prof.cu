// Phase timeline of lstm_bwdt_kernel (dev tool): cycles per timestep seen by a critical thread, a helper thread and an MMA issuer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200RL_BWDPROF \
//     -I pytorch-distributed-reinforcement-learning_b200/csrc -I include tools/bwd_prof.cu -o tools/_build/bwd_prof
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
namespace b200rl {
int set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return 1; }
int check_launch(const char* what) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e)); return 0; }
bool recurrent_tile16(int B, int sms) { const char* e = getenv("B200RL_TILE16"); return e ? atoi(e) != 0 : (B + 31) / 32 <= sms / 2; }   // (the library's lives in fwd_fused.cu)
}
#include "bwd_tc.cu"
using namespace b200rl;
int main(int argc, char** argv) {
    const int B = 4096, T = 128, NH = 3;
    const int mode = argc > 1 ? atoi(argv[1]) : 0;      // 0: no fused gradients ; 1: head gradients fused (dx stored) ; 2: head + torso fused
    const long long rows = (long long)B * T;
    float *gates, *c, *cx0, *Whh, *Wih, *Whead, *dhead, *dx;
    cudaMalloc(&gates, rows * 256 * 4); cudaMalloc(&c, rows * 64 * 4); cudaMalloc(&cx0, B * 64 * 4); cudaMalloc(&Whh, 256 * 64 * 4);
    cudaMalloc(&Wih, 256 * 64 * 4); cudaMalloc(&Whead, NH * 64 * 4); cudaMalloc(&dhead, rows * NH * 4); cudaMalloc(&dx, rows * 64 * 4);
    cudaMemset(gates, 0, rows * 256 * 4); cudaMemset(c, 0, rows * 64 * 4); cudaMemset(cx0, 0, B * 64 * 4); cudaMemset(Whh, 0, 256 * 64 * 4);
    cudaMemset(Wih, 0, 256 * 64 * 4); cudaMemset(Whead, 0, NH * 64 * 4); cudaMemset(dhead, 0, rows * NH * 4);
    float *u, *Wt, *part;
    const long long P = 33796;
    cudaMalloc(&u, rows * 4 * 4); cudaMalloc(&Wt, 64 * 5 * 4); cudaMalloc(&part, 148 * P * 4);
    cudaMemset(u, 0, rows * 4 * 4); cudaMemset(Wt, 0, 64 * 5 * 4);
    auto run = [&]() {
        if (mode == 0) return launch_bwd_tc(dhead, Whead, Whh, Wih, gates, c, cx0, dx, B, T, NH, 148, 0);
        return launch_bwd_tc_fused(dhead, Whead, Whh, Wih, gates, c, cx0, mode == 1 ? dx : nullptr, u, Wt, Wt + 256, part, P, 0, 256, 33000, 33200,
                                   B, T, 4, NH, 148, 0);
    };
    for (int it = 0; it < 2; ++it) run();
    cudaDeviceSynchronize();
    unsigned long long z[24] = {0};
#ifdef B200RL_BWDPROF
    cudaMemcpyToSymbol(g_bwdprof, z, sizeof(z));
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    run();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
#ifdef B200RL_BWDPROF
    cudaMemcpyFromSymbol(z, g_bwdprof, sizeof(z));
#else
    printf("bwd_tc (no probes, NI=%d, fused mode %d) %.3f ms\n", (mode == 2 ? B200RL_BWD_NI_TORSO : B200RL_BWD_NI), mode, ms);
    return 0;
#endif
    const double steps = (B / 32.0) * T;
    printf("bwd_tc (fused mode %d) %.3f ms; cycles/step A thread: loads issue %.0f | wait acc_done %.0f | tmem ld %.0f | publish + pair barrier %.0f | 4 cells + operand %.0f | fence + arrive %.0f | stores + coefs(t-1) %.0f\n",
           mode, ms, z[0] / steps, z[1] / steps, z[2] / steps, z[3] / steps, z[4] / steps, z[5] / steps, z[6] / steps);
    printf("   B thread: loads issue %.0f | wait acc_done %.0f | tmem ld %.0f | pair barrier + read dh %.0f | 4 cells + operand %.0f | fence + arrive %.0f | stores + dx + coefs(t-1) %.0f\n",
           z[8] / steps, z[9] / steps, z[10] / steps, z[11] / steps, z[12] / steps, z[13] / steps, z[14] / steps);
    printf("   issuer: loop top %.0f | wait b_full %.0f | issue 40 MMAs + commit %.0f\n", z[16] / steps, z[17] / steps, z[18] / steps);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
