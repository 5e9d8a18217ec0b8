// Common device helpers for the b200rl kernels (sm_100a).
//
// B200RL_EMUL: a TEST-ONLY build mode (tests/emul/) that compiles these same kernel
// sources with g++ and runs every CUDA thread as a host thread, so kernel indexing
// can be debugged in the GPU-less build container.  It is never built, loaded or
// shipped by the product (`__graft_entry__.build()` compiles with nvcc only).
#pragma once
#include <stdint.h>
#include <math.h>
#include <stddef.h>
#include "b200rl.h"

#ifdef B200RL_EMUL
#include "../../tests/emul/cuda_emul.h"
#else
#include <cuda_runtime.h>
#define B200RL_DYN_SMEM(T, name) extern __shared__ __align__(128) unsigned char name##_raw_[]; T* name = reinterpret_cast<T*>(name##_raw_)
#define B200RL_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

#ifdef B200RL_EMUL
#define B200RL_MAXNREG(n)
#else
#define B200RL_MAXNREG(n) __maxnreg__(n)
#endif

namespace b200rl {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Block-wide sum of doubles into lane 0 of warp 0 (valid for thread 0 only).
// `scratch` must hold >= 32 doubles of shared memory.  Contains __syncthreads().
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < nw ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// F.smooth_l1_loss(beta=1) value and derivative wrt d = x - y.
__device__ __forceinline__ float smooth_l1(float d) { float a = fabsf(d); return a < 1.0f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float smooth_l1_grad(float d) { return fabsf(d) < 1.0f ? d : (d > 0.0f ? 1.0f : -1.0f); }

}  // namespace b200rl

// ---- packed fp32 pairs: Blackwell issues two independent FMAs with one FFMA2 (fma.rn.f32x2) at the
// same FLOP rate as two FFMAs (tools/ffma2_probe.cu: 126 FMA/clk/SM either way), halving the issue
// slots of FFMA-bound inner loops.  Host twin: a plain struct.
namespace b200rl {
#ifdef B200RL_EMUL
struct f2 { float x, y; };
inline f2 f2_make(float a, float b) { return f2{a, b}; }
inline void f2_fma(f2& d, f2 a, f2 b) { d.x = fmaf(a.x, b.x, d.x); d.y = fmaf(a.y, b.y, d.y); }
inline float f2_lo(f2 v) { return v.x; }
inline float f2_hi(f2 v) { return v.y; }
inline void f2_ld2(const float* p, f2& a, f2& b) { a = f2{p[0], p[1]}; b = f2{p[2], p[3]}; }
#else
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_make(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_fma(f2& d, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ float f2_lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float f2_hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
// 16-byte load of four floats as two packed pairs
__device__ __forceinline__ void f2_ld2(const float* p, f2& a, f2& b) { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p); a = v.x; b = v.y; }
#endif
}  // namespace b200rl

namespace b200rl {
// sigmoid / tanh for the LSTM cell: sigm(x) = rcp(1 + ex2(-x log2 e)) on the two MUFU approximations (ex2: 2 ulp,
// rcp: 1 ulp; absolute error ~1e-7, far inside the rtol 1e-4 parity bar), tanh(x) = 2 sigm(2x) - 1.  Four / five
// instructions: expf + __fdividef cost ~14 with their range handling, and the cell phase of the recurrent
// kernels is issue-bound on exactly these (tools/fwd_prof.cu).  Saturation is exact: ex2 -> inf gives rcp -> 0.
#ifdef B200RL_EMUL
inline float ex2_approx(float x) { return exp2f(x); }
inline float rcp_approx(float x) { return 1.0f / x; }
#else
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
__device__ __forceinline__ float sigm(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_(float x) { return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f); }
// Two activations on THREE MUFU instead of four: with a = e^-x, b = e^-y one reciprocal r = 1/((1+a)(1+b)) serves both,
// sigm(x) = r (1+b), sigm(y) = r (1+a); with b = e^-2y instead, tanh(y) = r (1+a)(1-b).  Inputs are clamped at -30 (sigm(-30) = 9e-14, and the
// product of the two exponentials stays far below FLT_MAX); the cell phase of the tcgen05 forward is MUFU-bound.
__device__ __forceinline__ void sigm_sigm(float x, float y, float& sx, float& sy) {
    const float a = ex2_approx(-1.4426950408889634f * fmaxf(x, -30.0f)), b = ex2_approx(-1.4426950408889634f * fmaxf(y, -30.0f));
    const float pa = 1.0f + a, pb = 1.0f + b, r = rcp_approx(pa * pb);
    sx = r * pb; sy = r * pa;
}
__device__ __forceinline__ void sigm_tanh(float x, float y, float& sx, float& ty) {
    const float a = ex2_approx(-1.4426950408889634f * fmaxf(x, -30.0f)), b = ex2_approx(-2.8853900817779268f * fmaxf(y, -15.0f));
    const float pa = 1.0f + a, pb = 1.0f + b, r = rcp_approx(pa * pb);
    sx = r * pb; ty = (r * pa) * (1.0f - b);
}
#ifdef B200RL_EMUL      // the host emulator tolerates unaligned 16-byte accesses, the GPU traps on them: check here
#define B200RL_ALIGN16(p) do { if (reinterpret_cast<uintptr_t>(p) & 15) { fprintf(stderr, "emul: misaligned 16-byte access %s:%d\n", __FILE__, __LINE__); abort(); } } while (0)
#else
#define B200RL_ALIGN16(p) do { } while (0)
#endif
__device__ __forceinline__ float4 ld4(const float* p) { B200RL_ALIGN16(p); return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { B200RL_ALIGN16(p); *reinterpret_cast<float4*>(p) = v; }
}  // namespace b200rl

namespace b200rl {
// 4-byte asynchronous global -> shared copy (LDGSTS): no destination register, completion via cp_async_wait_all()
#ifdef B200RL_EMUL
inline void cp_async4(float* dst_smem, const float* src) { *dst_smem = *src; }
inline void cp_async_wait_all() {}
#else
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif
}  // namespace b200rl

// Error plumbing shared by the C-ABI entry points.
namespace b200rl {
int set_error(const char* fmt, ...);
int check_launch(const char* what);
}  // namespace b200rl
#define B200RL_REQUIRE(cond, ...) do { if (!(cond)) return ::b200rl::set_error(__VA_ARGS__); } while (0)
