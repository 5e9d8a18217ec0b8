// Data-parallel gradient exchange fused with the optimizer step, over NVLink / NVSwitch peer memory (sm_100a).
//
// Replaces, for a learner whose batch is sharded over the GPUs of one box,
//     dist.all_reduce(grad) ; clip_grad_norm_(params, max_norm) ; optimizer.step()
// (the reference has no multi-GPU learner; the single-GPU tail is agents/learner_module/ppo/learning.py:93-106 with the
// optimizers of agents/learner.py:70,360-367) by ONE kernel per step and NO library collective:
//
//   every rank keeps its flat gradient (+ the 8 fp64 loss sums) in a buffer the other ranks have mapped through CUDA IPC;
//   1. announce : one release-store of the step number into each peer's flag array ("my gradient of step e is complete" --
//                 it was written by the preceding kernels of this stream);
//   2. wait     : acquire-poll my own flag array until every peer has announced step e (bounded: a peer that never
//                 arrives raises a status word instead of hanging the GPU);
//   3. one-shot all-reduce: every CTA sums ITS slice of the gradient over all ranks, reading the peers' buffers directly
//                 (16-byte volatile loads over NVLink; 135 KB per peer for the CartPole actor-critic -- latency-, not
//                 bandwidth-bound, which is why a ring / tree collective with its extra hops loses here).  The sum runs in
//                 rank order 0..W-1 on every rank, so all ranks hold bit-identical gradients and stay in lock-step;
//   4. global L2 norm: per-CTA fp64 partials, a grid barrier, every CTA adds the partials in the same order;
//   5. clip coefficient min(1, max_norm / (norm + 1e-6)), then RMSprop or Adam on the slice still held in registers.
// The buffers are double-buffered by step parity: a rank may run at most one step ahead of the slowest peer (it cannot
// pass step e+1's wait before every peer has finished reading step e), so buffer e & 1 is never overwritten while read.
//
// With world == 1 the kernel degenerates to the fused clip + optimizer step on more than one CTA.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace b200rl {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 256;
constexpr int kPeerMaxGrid = 64;
#ifdef B200RL_EMUL
constexpr int kPeerVecPerThread = 96;            // (the host emulator runs ONE CTA: it must hold the whole vector)
#else
constexpr int kPeerVecPerThread = 4;             // float4 chunks a thread keeps in registers between the reduce and the update
#endif
constexpr int kFlagStride = 32;                  // one 128-byte line per flag

struct PeerStepArgs {
    const float* peer_grad[kPeerMaxWorld];       // each rank's gradient buffer of this step (index = rank; [rank] is local)
    const double* peer_scal[kPeerMaxWorld];      // each rank's 8 loss sums of this step (may be null)
    unsigned* peer_flag[kPeerMaxWorld];          // &flags_of_rank_r[my_rank * kFlagStride]
    unsigned* my_flags;                          // flags[world * kFlagStride], written by the peers
    unsigned* gsync;                             // local grid-barrier counter (monotonic)
    double* part;                                // local [kPeerMaxGrid] per-CTA partial sums of squares
    float* param; float* s1; float* s2;          // RMSprop: s1 = square average ; Adam: s1 = m, s2 = v
    const float* mask;                           // optional 0/1 gradient mask (SAC-continuous critic torso)
    float* grad_out;                             // local: reduced, scaled and clipped gradient (what `.grad` holds after the step)
    double* scal_out;                            // local: reduced loss sums
    float* norm_out;
    int* status;                                 // set to 1 if a peer never announced (timeout)
    long long n;                                 // floats, multiple of 4
    int rank, world;
    unsigned epoch;                              // step number, >= 1, +1 per launch of this communicator
    unsigned gsync_target;                       // value the grid-barrier counter reaches when every CTA of this launch arrived
    float grad_scale, max_norm, lr, a1, a2, eps; // RMSprop: a1 = alpha ; Adam: a1 = beta1, a2 = beta2
    float adam_step_size, adam_rsq_bc2;
    long long timeout_cycles;
};

#ifdef B200RL_EMUL
static inline void st_release_sys(unsigned* p, unsigned v) { std::atomic_thread_fence(std::memory_order_seq_cst); *p = v; }
static inline unsigned ld_acquire_sys(const unsigned* p) { unsigned v = *(volatile const unsigned*)p; std::atomic_thread_fence(std::memory_order_seq_cst); return v; }
static inline float4 ld_peer4(const float* p) { return *reinterpret_cast<const float4*>(p); }
static inline long long clock64() { return 0; }
#else
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// peer (or local) gradient data: never served from a stale L1 line
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
#endif

template <int OPT>   // 0 = RMSprop, 1 = Adam
__global__ void __launch_bounds__(kPeerThreads)
peer_step_kernel(PeerStepArgs a) {
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    __shared__ double scratch[32];
    __shared__ float coef_s;
    __shared__ int timed_out;
    if (tid == 0) timed_out = 0;
    // 1. announce (CTA 0), 2. wait (every CTA polls for itself: no second grid barrier needed)
    if (a.world > 1) {
        if (cta == 0 && tid < a.world && tid != a.rank) {
            __threadfence_system();
            st_release_sys(a.peer_flag[tid], a.epoch);
        }
        __syncthreads();
        if (tid < a.world && tid != a.rank) {
            const unsigned* f = a.my_flags + tid * kFlagStride;
            const long long t0 = clock64();
            // (signed distance: the counter may wrap after 2^32 steps)
            while ((int)(ld_acquire_sys(f) - a.epoch) < 0) {
                if (clock64() - t0 > a.timeout_cycles) { timed_out = 1; break; }
            }
        }
        __syncthreads();
        if (timed_out && tid == 0) *a.status = 1;
    }
    // 3. this CTA's slice: float4 chunk c = (k * G + cta) * kPeerThreads + tid
    const long long nvec = a.n >> 2;
    float4 g[kPeerVecPerThread];
    double ss = 0.0;
#pragma unroll
    for (int k = 0; k < kPeerVecPerThread; ++k) {
        const long long c = ((long long)k * G + cta) * kPeerThreads + tid;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < nvec) {
            float4 v[kPeerMaxWorld];
#pragma unroll
            for (int r = 0; r < kPeerMaxWorld; ++r) if (r < a.world) v[r] = ld_peer4(a.peer_grad[r] + 4 * c);     // all loads in flight
#pragma unroll
            for (int r = 0; r < kPeerMaxWorld; ++r) if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
            s.x *= a.grad_scale; s.y *= a.grad_scale; s.z *= a.grad_scale; s.w *= a.grad_scale;
            if (a.mask) { const float4 m = ld4(a.mask + 4 * c); s.x *= m.x; s.y *= m.y; s.z *= m.z; s.w *= m.w; }
            ss += (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z + (double)s.w * s.w;
        }
        g[k] = s;
    }
    // 4. global norm
    const double mine = block_sum(ss, scratch);
    if (tid == 0) {
        a.part[cta] = mine;
        if (G > 1) {
            __threadfence();
            atomicAdd(a.gsync, 1u);
#ifndef B200RL_EMUL
            while ((int)(*(volatile unsigned*)a.gsync - a.gsync_target) < 0) { }
#endif
            __threadfence();
        }
        double tot = 0.0;
        for (int i = 0; i < G; ++i) tot += ((volatile double*)a.part)[i];
        const float norm = (float)sqrt(tot);
        coef_s = a.max_norm > 0.0f ? fminf(a.max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
        if (cta == 0 && a.norm_out) *a.norm_out = norm;
    }
    __syncthreads();
    const float coef = coef_s;
    // 5. optimizer step on the slice
#pragma unroll
    for (int k = 0; k < kPeerVecPerThread; ++k) {
        const long long c = ((long long)k * G + cta) * kPeerThreads + tid;
        if (c >= nvec) continue;
        const float gg[4] = {g[k].x * coef, g[k].y * coef, g[k].z * coef, g[k].w * coef};
        const float4 w = ld4(a.param + 4 * c), q1 = ld4(a.s1 + 4 * c);
        float wv[4] = {w.x, w.y, w.z, w.w}, q1v[4] = {q1.x, q1.y, q1.z, q1.w};
        if (OPT == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v = q1v[e] * a.a1 + (1.0f - a.a1) * gg[e] * gg[e];
                q1v[e] = v;
                wv[e] = wv[e] - a.lr * (gg[e] / (sqrtf(v) + a.eps));
            }
        } else {
            const float4 q2 = ld4(a.s2 + 4 * c);
            float q2v[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float mi = q1v[e] * a.a1 + (1.0f - a.a1) * gg[e];
                const float vi = q2v[e] * a.a2 + (1.0f - a.a2) * gg[e] * gg[e];
                q1v[e] = mi; q2v[e] = vi;
                wv[e] = wv[e] - a.adam_step_size * (mi / (sqrtf(vi) * a.adam_rsq_bc2 + a.eps));
            }
            st4(a.s2 + 4 * c, make_float4(q2v[0], q2v[1], q2v[2], q2v[3]));
        }
        st4(a.s1 + 4 * c, make_float4(q1v[0], q1v[1], q1v[2], q1v[3]));
        st4(a.param + 4 * c, make_float4(wv[0], wv[1], wv[2], wv[3]));
        st4(a.grad_out + 4 * c, make_float4(gg[0], gg[1], gg[2], gg[3]));
    }
    // loss sums for logging: [0..3] sums, [4] min, [5] max over ranks, [6..7] taken from this rank
    if (cta == 0 && tid < 8 && a.scal_out && a.peer_scal[a.rank]) {
        double v = 0.0;
        if (tid < 4) { for (int r = 0; r < a.world; ++r) v += ((volatile const double*)a.peer_scal[r])[tid]; }
        else if (tid == 4) { v = INFINITY; for (int r = 0; r < a.world; ++r) v = fmin(v, ((volatile const double*)a.peer_scal[r])[tid]); }
        else if (tid == 5) { v = -INFINITY; for (int r = 0; r < a.world; ++r) v = fmax(v, ((volatile const double*)a.peer_scal[r])[tid]); }
        else v = a.peer_scal[a.rank][tid];
        a.scal_out[tid] = v;
    }
}

static long long g_peer_timeout_ms = 30000;

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_peer_set_timeout_ms(long long ms) { g_peer_timeout_ms = ms > 0 ? ms : 1; return 0; }

#ifndef B200RL_EMUL
extern "C" int b200rl_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out) {
    B200RL_REQUIRE(bytes > 0 && ptr_out && handle64_out, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return set_error("peer_alloc: cudaMalloc(%lld): %s", bytes, cudaGetErrorString(e));
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return set_error("peer_alloc: %s", cudaGetErrorString(e)); }
    memcpy(handle64_out, &h, 64);
    *ptr_out = p;
    return 0;
}
extern "C" int b200rl_peer_open(const unsigned char* handle64, void** ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_error("peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *ptr_out = p;
    return 0;
}
extern "C" int b200rl_peer_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) return set_error("peer_close: %s", cudaGetErrorString(e));
    return 0;
}
extern "C" int b200rl_peer_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) return set_error("peer_free: %s", cudaGetErrorString(e));
    return 0;
}
#else
extern "C" int b200rl_peer_alloc(long long, void**, unsigned char*) { return set_error("peer_alloc: not available in the host emulator"); }
extern "C" int b200rl_peer_open(const unsigned char*, void**) { return set_error("peer_open: not available in the host emulator"); }
extern "C" int b200rl_peer_close(void*) { return 0; }
extern "C" int b200rl_peer_free(void*) { return 0; }
#endif

extern "C" int b200rl_peer_step(int optimizer, float* param, float* s1, float* s2, const float* mask, float* grad_out,
                                const void* const* peer_grad, const void* const* peer_scal, void* const* peer_flags,
                                void* my_flags, void* gsync, double* part, double* scal_out, float* norm_out, int* status,
                                long long n, int rank, int world, unsigned epoch, int grid, unsigned gsync_target,
                                float grad_scale, float max_norm, float lr, float a1, float a2, float eps, int adam_step,
                                void* stream) {
    B200RL_REQUIRE(optimizer == 0 || optimizer == 1, "peer_step: optimizer %d", optimizer);
    B200RL_REQUIRE(n >= 4 && (n & 3) == 0, "peer_step: n=%lld must be a positive multiple of 4", n);
    B200RL_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer_step: rank %d of %d", rank, world);
    B200RL_REQUIRE(epoch >= 1, "peer_step: epoch starts at 1");
#ifdef B200RL_EMUL
    grid = 1;
#endif
    B200RL_REQUIRE(grid >= 1 && grid <= kPeerMaxGrid, "peer_step: grid %d", grid);
    B200RL_REQUIRE((n >> 2) <= (long long)grid * kPeerThreads * kPeerVecPerThread,
                   "peer_step: %lld parameters exceed what %d CTAs keep in registers", n, grid);
    B200RL_REQUIRE(optimizer == 0 || (s2 && adam_step >= 1), "peer_step: Adam needs s2 and step >= 1");
    PeerStepArgs a{};
    for (int r = 0; r < world; ++r) {
        a.peer_grad[r] = (const float*)peer_grad[r];
        a.peer_scal[r] = peer_scal ? (const double*)peer_scal[r] : nullptr;
        a.peer_flag[r] = peer_flags ? (unsigned*)peer_flags[r] + (size_t)rank * kFlagStride : nullptr;
        B200RL_REQUIRE(a.peer_grad[r], "peer_step: null gradient buffer for rank %d", r);
        B200RL_REQUIRE(world == 1 || a.peer_flag[r], "peer_step: null flag array for rank %d", r);
    }
    a.my_flags = (unsigned*)my_flags; a.gsync = (unsigned*)gsync; a.part = part;
    a.param = param; a.s1 = s1; a.s2 = s2; a.mask = mask; a.grad_out = grad_out; a.scal_out = scal_out; a.norm_out = norm_out; a.status = status;
    a.n = n; a.rank = rank; a.world = world; a.epoch = epoch; a.gsync_target = gsync_target;
    a.grad_scale = grad_scale; a.max_norm = max_norm; a.lr = lr; a.a1 = a1; a.a2 = a2; a.eps = eps;
    if (optimizer == 1) {
        const double bc1 = 1.0 - pow((double)a1, (double)adam_step), bc2 = 1.0 - pow((double)a2, (double)adam_step);
        a.adam_step_size = (float)((double)lr / bc1);
        a.adam_rsq_bc2 = (float)(1.0 / sqrt(bc2));
    }
    a.timeout_cycles = g_peer_timeout_ms * 1900000LL;       // ~1.9 GHz SM clock
#ifdef B200RL_EMUL
    if (optimizer == 0) { B200RL_LAUNCH(peer_step_kernel<0>, 1, kPeerThreads, 0, stream, a); }
    else { B200RL_LAUNCH(peer_step_kernel<1>, 1, kPeerThreads, 0, stream, a); }
#else
    // cooperative launch: the grid barrier needs every CTA resident (grid <= 64 CTAs of 256 threads always fits)
    void* kargs[] = {&a};
    const void* fn = optimizer == 0 ? (const void*)peer_step_kernel<0> : (const void*)peer_step_kernel<1>;
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPeerThreads), kargs, 0, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error("peer_step: cudaLaunchCooperativeKernel: %s", cudaGetErrorString(e));
#endif
    return check_launch("peer_step");
}
