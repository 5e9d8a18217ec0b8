// TEST-ONLY host emulation of the small CUDA subset the b200rl kernels use.
// Every CUDA thread of a block runs as a std::thread; blocks run one after another.
// Purpose: debug kernel indexing / barriers in the GPU-less build container.
// Not part of the product; never loaded by the package.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include <math.h>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
struct alignas(8) uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emul {
struct WarpCtx {
    std::barrier<> bar{32};
    unsigned long long buf[32];
};
struct BlockCtx {
    std::unique_ptr<std::barrier<>> bar;
    std::mutex named_mu;
    std::unique_ptr<std::barrier<>> named[16];
    std::vector<std::unique_ptr<WarpCtx>> warps;
    unsigned char* dyn = nullptr;
};
inline BlockCtx*& cur_block() { static BlockCtx* b = nullptr; return b; }
inline dim3& g_blockDim() { static dim3 d; return d; }
inline dim3& g_gridDim() { static dim3 d; return d; }
inline std::mutex& atomic_mu() { static std::mutex m; return m; }
inline void named_barrier(int id, int nthreads);
}  // namespace emul

inline thread_local dim3 threadIdx;
inline thread_local dim3 blockIdx;
#define blockDim (emul::g_blockDim())
#define gridDim (emul::g_gridDim())

static inline void __syncthreads() { emul::cur_block()->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::cur_block()->warps[threadIdx.x >> 5]->bar.arrive_and_wait(); }

template <typename T>
static inline T emul_shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    emul::WarpCtx& w = *emul::cur_block()->warps[threadIdx.x >> 5];
    int lane = threadIdx.x & 31;
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.buf[lane] = raw;
    w.bar.arrive_and_wait();
    unsigned long long got = w.buf[src_lane & 31];
    w.bar.arrive_and_wait();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_shfl(v, (int)(threadIdx.x & 31) ^ m); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_shfl(v, src); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    int lane = threadIdx.x & 31;
    return emul_shfl(v, lane + d < 32 ? lane + d : lane);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    int lane = threadIdx.x & 31;
    return emul_shfl(v, lane - d >= 0 ? lane - d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (emul_shfl(pred ? 1 : 0, i) ? 1u : 0u) << i;
    return r;
}

template <typename T> static inline T atomicAdd(T* p, T v) {
    std::lock_guard<std::mutex> g(emul::atomic_mu());
    T old = *p;
    *p = old + v;
    return old;
}
static inline unsigned atomicInc(unsigned* p, unsigned wrap) {
    std::lock_guard<std::mutex> g(emul::atomic_mu());
    unsigned old = *p;
    *p = old >= wrap ? 0 : old + 1;
    return old;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }

namespace emul {
inline void named_barrier(int id, int nthreads) {
    BlockCtx* b = cur_block();
    std::barrier<>* bar;
    {
        std::lock_guard<std::mutex> g(b->named_mu);
        if (!b->named[id]) b->named[id].reset(new std::barrier<>(nthreads));
        bar = b->named[id].get();
    }
    bar->arrive_and_wait();
}
}  // namespace emul

#define B200RL_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emul::cur_block()->dyn)

namespace emul {
template <typename F>
void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
    g_blockDim() = block;
    g_gridDim() = grid;
    unsigned nthreads = block.x * block.y * block.z;
    if (nthreads % 32 != 0) { fprintf(stderr, "emul: block size must be a multiple of 32\n"); abort(); }
    std::vector<unsigned char> dyn(smem + 256);
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            BlockCtx ctx;
            ctx.bar.reset(new std::barrier<>(nthreads));
            for (unsigned w = 0; w < nthreads / 32; ++w) ctx.warps.emplace_back(new WarpCtx());
            ctx.dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 127) & ~uintptr_t(127));
            cur_block() = &ctx;
            std::vector<std::thread> ts;
            ts.reserve(nthreads);
            for (unsigned t = 0; t < nthreads; ++t)
                ts.emplace_back([&, t, bx, by] {
                    threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    blockIdx = dim3(bx, by, 0);
                    body();
                });
            for (auto& th : ts) th.join();
            cur_block() = nullptr;
        }
}
}  // namespace emul

#define B200RL_LAUNCH(kern, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), (smem), [&] { kern(__VA_ARGS__); })

static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
