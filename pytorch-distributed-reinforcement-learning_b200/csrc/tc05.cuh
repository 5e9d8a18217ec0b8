// Minimal tcgen05 / TMEM / mbarrier wrappers (sm_100a inline PTX), plus a TEST-ONLY host
// twin (B200RL_EMUL) that interprets the same shared-memory descriptors with plain loops
// so that tile filling, descriptor strides and the epilogue mapping can be debugged
// without a GPU.  Bit layouts follow the CUTLASS headers vendored in this image
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor;
//  cute/atom/mma_traits_sm100.hpp: canonical no-swizzle layouts).
//
// Operands are fp32 containers read as TF32 (top 19 bits).  Canonical SWIZZLE_NONE
// layouts, in bytes, T = 4 elements per 16 B:
//   MN-major: elem(mn, k) at  (mn%4)*4 + (mn/4)*SBO + (k%8)*16 + (k/8)*LBO
//   K-major : elem(mn, k) at  (mn%8)*16 + (mn/8)*SBO + (k%4)*4  + (k/4)*LBO
// i.e. both are built from the same 128-byte "core matrix" of 8 rows x 16 bytes.
#pragma once
#include <string.h>
#include "common.cuh"

namespace b200rl {
namespace tc {

constexpr uint32_t kTf32Mask = 0xffffe000u;

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & kTf32Mask);
    lo = __uint_as_float(__float_as_uint(v - hi) & kTf32Mask);
}

// instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                         // c_format = F32
         | (2u << 7) | (2u << 10)            // a_format = b_format = TF32
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// instruction descriptor for kind::f16 with bf16 inputs, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// fp32 -> bf16 bits by truncation (used for the small third split term only) / back
__host__ __device__ inline unsigned short bf16_bits(float f) { unsigned u; memcpy(&u, &f, 4); return (unsigned short)(u >> 16); }
// fp32 -> bf16 bits, round to nearest even: zero-mean error (truncation is one-sided, and a recurrence accumulates the bias)
__host__ __device__ inline unsigned short bf16_bits_rn(float f) { unsigned u; memcpy(&u, &f, 4); return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16); }

// two fp32 -> one 32-bit word of two bf16 (round to nearest even): `e0` in the low half (the lower k), `e1` in the high half
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float e0, float e1) {
#ifdef B200RL_EMUL
    return (uint32_t)bf16_bits_rn(e0) | ((uint32_t)bf16_bits_rn(e1) << 16);
#else
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
#endif
}

#ifndef B200RL_EMUL
// ------------------------------------------------------------------------------ device
typedef uint64_t SmemDesc;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ SmemDesc make_desc(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t a = smem_u32(smem_ptr);
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3FFF);                 // start address   bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading offset  bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride offset   bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell)
    return d;                                           // base_offset 0, lbo_mode 0, SWIZZLE_NONE
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B200RL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra B200RL_DONE;\n"
        "bra B200RL_WAIT;\n"
        "B200RL_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival per WARP (barrier count = participating warps): 32 per-thread arrivals on one mbarrier serialise
// (measured: 256 arrivals cost ~1.3K cycles per stage in lstm_dx_tc, tools/dx_prof.cu).  Every lane must have
// executed its own proxy fence before the call.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// barrier among a subset of the CTA's warps (id 1..15; `nthreads` a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#ifdef B200RL_PROBE_NOFENCE      // timing experiment only (tools/dx_prof.cu): results are not valid without the fence
__device__ __forceinline__ void fence_async_smem() {}
#else
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; writes the TMEM base address (lane 0, column 0 of the allocation) to *dst
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, one elected thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, SmemDesc a, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T with bf16 inputs, K = 16, both operands K-major (8 bf16 per 16-byte core-matrix row)
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, SmemDesc a, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : the A tile (M = 128 lanes x K = 8 columns, one 32-bit cell per tf32 element)
// is read from tensor memory, so only B crosses the 128 B/clk shared-memory port
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T with bf16 inputs, K = 16: a 32-bit TMEM column of A holds k = 2c (low half) and
// k = 2c + 1 (high half) (measured, tools/tc_probe4.cu); B is K-major, 8 bf16 per 16-byte core-matrix row
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on `bar` once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    uint32_t u[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
          "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
          "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
// several tcgen05.ld in flight, one wait: issue with tmem_ld8_async (results are NOT valid yet), then tmem_ld_wait()
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&u)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 4 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&r)[4]) {
    uint32_t u[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = __uint_as_float(u[i]);
}
// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// store 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        ::"r"(taddr), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])),
          "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])),
          "r"(__float_as_uint(r[8])), "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])), "r"(__float_as_uint(r[11])),
          "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])), "r"(__float_as_uint(r[14])), "r"(__float_as_uint(r[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }

#else
// --------------------------------------------------------------------------- host twin
struct SmemDesc { const unsigned char* base; uint32_t lbo, sbo; };
inline SmemDesc make_desc(const void* p, uint32_t lbo, uint32_t sbo) { return SmemDesc{(const unsigned char*)p, lbo, sbo}; }

struct EmulTmem { float d[128][512]; };
inline EmulTmem& emul_tmem() { static EmulTmem t; return t; }

// host mbarrier: bits [0,16) pending arrivals, [16,32) arrival count, bit 32 phase parity
inline void mbar_init(uint64_t* bar, uint32_t count) {
    if (reinterpret_cast<uintptr_t>(bar) & 7) { fprintf(stderr, "emul: mbarrier at a misaligned address (the GPU traps on this)\n"); abort(); }
    *bar = (uint64_t)count | ((uint64_t)count << 16);
}
inline void fence_mbar_init() {}
inline void mbar_arrive(uint64_t* bar) {
    std::lock_guard<std::mutex> g(emul::atomic_mu());
    uint64_t v = *bar;
    uint32_t pending = (uint32_t)(v & 0xffff) - 1, count = (uint32_t)((v >> 16) & 0xffff), phase = (uint32_t)((v >> 32) & 1);
    if (pending == 0) { pending = count; phase ^= 1; }
    *bar = (uint64_t)pending | ((uint64_t)count << 16) | ((uint64_t)phase << 32);
}
inline void mbar_arrive_warp(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (;;) {
        { std::lock_guard<std::mutex> g(emul::atomic_mu()); if ((uint32_t)((*bar >> 32) & 1) != (parity & 1)) return; }
        std::this_thread::yield();
    }
}
inline void named_bar_sync(int id, int nthreads) { emul::named_barrier(id, nthreads); }
inline void fence_async_smem() {}
inline void fence_before_sync() {}
inline void fence_after_sync() {}
inline void tmem_alloc(uint32_t* dst, uint32_t) { if ((threadIdx.x & 31) == 0) *dst = 0; }
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void mma_commit(uint64_t* bar) { mbar_arrive(bar); }   // emulated MMAs complete synchronously

inline float emul_elem(const SmemDesc& d, bool mn_major, int mn, int k) {
    size_t off = mn_major ? (size_t)(mn % 4) * 4 + (size_t)(mn / 4) * d.sbo + (size_t)(k % 8) * 16 + (size_t)(k / 8) * d.lbo
                          : (size_t)(mn % 8) * 16 + (size_t)(mn / 8) * d.sbo + (size_t)(k % 4) * 4 + (size_t)(k / 4) * d.lbo;
    float v;
    memcpy(&v, d.base + off, 4);
    return __uint_as_float(__float_as_uint(v) & kTf32Mask);
}
inline void mma_tf32(uint32_t d_tmem, SmemDesc a, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    const int M = (int)((idesc >> 24) & 31) << 4, N = (int)((idesc >> 17) & 63) << 3;
    const bool amn = (idesc >> 15) & 1, bmn = (idesc >> 16) & 1;
    EmulTmem& t = emul_tmem();
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = accumulate ? t.d[m][d_tmem + n] : 0.f;
            for (int k = 0; k < 8; ++k) s += emul_elem(a, amn, m, k) * emul_elem(b, bmn, n, k);
            t.d[m][d_tmem + n] = s;
        }
}
inline void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    const int M = (int)((idesc >> 24) & 31) << 4, N = (int)((idesc >> 17) & 63) << 3;
    const bool bmn = (idesc >> 16) & 1;
    EmulTmem& t = emul_tmem();
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = accumulate ? t.d[m][d_tmem + n] : 0.f;
            for (int k = 0; k < 8; ++k) s += __uint_as_float(__float_as_uint(t.d[m][a_tmem + k]) & kTf32Mask) * emul_elem(b, bmn, n, k);
            t.d[m][d_tmem + n] = s;
        }
}
inline void mma_bf16(uint32_t d_tmem, SmemDesc a, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    const int M = (int)((idesc >> 24) & 31) << 4, N = (int)((idesc >> 17) & 63) << 3;
    EmulTmem& t = emul_tmem();
    auto el = [](const SmemDesc& d, int mn, int k) {
        unsigned short h;
        memcpy(&h, d.base + (size_t)(mn % 8) * 16 + (size_t)(mn / 8) * d.sbo + (size_t)(k % 8) * 2 + (size_t)(k / 8) * d.lbo, 2);
        unsigned u = (unsigned)h << 16; float f; memcpy(&f, &u, 4); return f;
    };
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = accumulate ? t.d[m][d_tmem + n] : 0.f;
            for (int k = 0; k < 16; ++k) s += el(a, m, k) * el(b, n, k);
            t.d[m][d_tmem + n] = s;
        }
}
inline void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, SmemDesc b, uint32_t idesc, uint32_t accumulate) {
    const int M = (int)((idesc >> 24) & 31) << 4, N = (int)((idesc >> 17) & 63) << 3;
    EmulTmem& t = emul_tmem();
    auto up = [](unsigned short h) { unsigned u = (unsigned)h << 16; float f; memcpy(&f, &u, 4); return f; };
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = accumulate ? t.d[m][d_tmem + n] : 0.f;
            for (int k = 0; k < 16; ++k) {
                const unsigned cell = __float_as_uint(t.d[m][a_tmem + k / 2]);
                unsigned short bb;
                memcpy(&bb, b.base + (size_t)(n % 8) * 16 + (size_t)(n / 8) * b.sbo + (size_t)(k % 8) * 2 + (size_t)(k / 8) * b.lbo, 2);
                s += up((unsigned short)(k & 1 ? cell >> 16 : cell & 0xffff)) * up(bb);
            }
            t.d[m][d_tmem + n] = s;
        }
}
inline void tmem_st16(uint32_t taddr, const float (&r)[16]) {
    const int lane = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    for (int i = 0; i < 16; ++i) emul_tmem().d[lane][col + i] = r[i];
}
inline void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    const int lane = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    for (int i = 0; i < 32; ++i) r[i] = emul_tmem().d[lane][col + i];
}
inline void tmem_ld8_async(uint32_t taddr, uint32_t (&u)[8]) {
    const int lane = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(emul_tmem().d[lane][col + i]);
}
inline void tmem_ld_wait() {}
inline void tmem_ld4(uint32_t taddr, float (&r)[4]) {
    const int lane = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    for (int i = 0; i < 4; ++i) r[i] = emul_tmem().d[lane][col + i];
}
inline void tmem_ld8(uint32_t taddr, float (&r)[8]) {
    const int lane = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    for (int i = 0; i < 8; ++i) r[i] = emul_tmem().d[lane][col + i];
}
inline void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    const int lane = (int)(taddr >> 16), col = (int)(taddr & 0xffff);
    for (int i = 0; i < 16; ++i) r[i] = emul_tmem().d[lane][col + i];
}
inline uint32_t tmem_addr(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }
#endif

}  // namespace tc
}  // namespace b200rl
