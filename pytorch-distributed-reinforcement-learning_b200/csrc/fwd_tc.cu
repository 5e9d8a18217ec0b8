// Recurrent half of the LSTM forward on tcgen05 (sm_100a).
//
//   pre[b,t,n] = xproj[b,t,n] + sum_k W_hh[n][k] * h[b,t-1,k]      (n = g*64 + j, gates i,f,g,o)
//   c_t = sigm(f) * c_{t-1} + sigm(i) * tanh(g) ;  h_t = sigm(o) * tanh(c_t)
//
// (nn.LSTMCell looped over T, networks/models.py:71-75.)  The input half `xproj` comes from
// lstm_xproj_tc_kernel; this kernel runs the serial part with the per-step GEMM on the tensor
// cores instead of FFMA (the FFMA version is shared-memory-wavefront bound at ~9K cycles per step).
//
// Per CTA: 32 sequences for all T steps.  The GEMM is issued "weights as M": D[256 x 32] =
// W_hh[256 x 64] . h_{t-1}^T[64 x 32], two M = 128 tiles (rows n = 0..127: gates i, f; n = 128..255:
// gates g, o), N = 32 sequences, K = 64.  W_hh is split once into TF32 hi/lo and stays resident IN TENSOR
// MEMORY as the A operand (lane = row n % 128, one column per k: 2 tiles x (hi, lo) x 64 = 256 columns);
// h_{t-1} is the K-major B operand in shared memory (hi/lo), rewritten by the cell phase every step.
// 3xTF32 (hi*hi + hi*lo + lo*hi): 48 tcgen05.mma per step into 64 more TMEM columns.  With A in shared
// memory every MMA pulled 4 KB of weights through the 128 B/clk port (240 KB = 1.9K cycles per step,
// measured with tools/fwd_prof.cu); from TMEM only the 1 KB B slices cross it.
//
// Cell phase, 16 warps: TMEM lane l of tile X holds gate i (l < 64) or f (l >= 64) of unit l % 64, tile Y
// holds g or o, and a warp can only read its own lane quarter -- so warps with id % 4 in {0, 1} see (i, g)
// of 32 units each and id % 4 in {2, 3} see (f, o) of the same units; warps 4g .. 4g+3 take sequences
// 8g .. 8g+7.  Each thread reads its two gate rows for 8 sequences, hands the raw sums of 4 of them to
// its partner thread (same unit, other gate pair) through shared memory and gets the partner's for the
// other 4, then owns 4 complete cells: add xproj, four activations, c, h, saved
// activations, next B operand.  Global accesses are 128-byte (c, h) or 512-byte (xproj in, gates out)
// contiguous per warp and row, each byte touched by exactly one thread.
//
// Four issuer warps, (M tile) x (even / odd K-steps), each with its own accumulator: one thread gets a small
// tcgen05.mma out only every ~50-120 cycles, issues from different threads overlap (tools/tc_probe3.cu), and the
// tensor pipe needs ~16 cycles for M128 x N32 x K8 -- the 48 MMAs of a step are issue-bound.
//
// xg: on entry xproj (B, T, 64, 4), on exit the saved activations, same layout, in place: the thread
// that owns cell (row, unit) reads its 16 bytes one step ahead and overwrites them itself.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

// phase timeline for tools/fwd_prof.cu (never compiled into the product library)
#ifdef B200RL_FWDPROF
__device__ unsigned long long g_fwdprof[16];
#define FWP_INIT unsigned long long fwp_prev = clock64();
#define FWP(i, who) do { if (tid == (who)) { unsigned long long now_ = clock64(); atomicAdd(&g_fwdprof[i], now_ - fwp_prev); fwp_prev = now_; } } while (0)
#else
#define FWP_INIT
#define FWP(i, who) do { } while (0)
#endif

namespace ftk {
constexpr int H = 64, G = 256, TB = 32;
constexpr int NG = 4;                                      // sequence groups: 4 cell warps (one per TMEM lane quarter) each
constexpr int RW = TB / NG, CT = RW / 2;                   // sequences a warp reads / cells a thread owns
constexpr int NI = 4;                                      // MMA issuer warps: (M tile) x (even / odd K-steps), one accumulator each
constexpr int NC = NG * 128, NT = NC + 32 * NI;
constexpr int TM_D = 0, TM_A = NI * TB, TM_COLS = 512;     // TMEM columns: accumulator (mt, khalf) at 32 (2 mt + khalf) ; A at TM_A + (mt*2 + lo)*64 + k
constexpr int B_LBO = 144, B_SBO = 2304;                   // B: 16 core matrices along K, 16-byte skew -> conflict-free 4-byte scatter
constexpr int B_BYTES = (TB / 8) * B_SBO;                  // 9216 per hi / lo
constexpr int EXF = 2 * NG * 2 * RW * 32;                  // swap buffer: [pair][direction][slot][lane] floats
static_assert(RW == 8, "the cell phase reads 8 accumulator columns per tile (tmem_ld8)");
static_assert(15 * B_LBO + 128 <= B_SBO, "core matrices of one 8-row group must fit below SBO");
}  // namespace ftk

static inline size_t fwdt_smem_bytes() { return 2 * (size_t)ftk::B_BYTES + sizeof(float) * ftk::EXF + 64; }

__global__ void __launch_bounds__(ftk::NT, 1)
lstm_fwdt_kernel(float* xg, const float* __restrict__ hx0, const float* __restrict__ cx0,
                 const float* __restrict__ W_hh, float* __restrict__ c_out, float* __restrict__ h_out, int B, int T) {
    using namespace ftk;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* B_hi = smem;
    unsigned char* B_lo = B_hi + B_BYTES;
    float* EX = reinterpret_cast<float*>(B_lo + B_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(EX + EXF);
    uint64_t* b_full = bars, *acc_done = bars + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool compute = warp < NC / 32;

    if (tid == 0) {
        tc::mbar_init(b_full, NC / 32); tc::mbar_init(acc_done, NI);
        tc::fence_mbar_init();
    }
    if (warp == NC / 32) tc::tmem_alloc(tmem_slot, TM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    // one-time: W_hh (PyTorch row order n = g*64 + j) -> TF32 hi/lo in TMEM: warps 0-3 fill tile 0, warps 4-7 tile 1,
    // thread = row n (its own TMEM lane), 16 columns per store
    if (warp < 8) {
        const int mt = warp >> 2, m = 32 * (warp & 3) + lane;
        const float* wrow = W_hh + (size_t)(mt * 128 + m) * H;
#ifdef B200RL_EMUL
        const int st_lane = m;
#else
        const int st_lane = 32 * (warp & 3);
#endif
        for (int k0 = 0; k0 < H; k0 += 16) {
            float hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                const float4 w = ld4(wrow + k0 + e);
                tc::split_tf32(w.x, hi[e], lo[e]); tc::split_tf32(w.y, hi[e + 1], lo[e + 1]);
                tc::split_tf32(w.z, hi[e + 2], lo[e + 2]); tc::split_tf32(w.w, hi[e + 3], lo[e + 3]);
            }
            tc::tmem_st16(tc::tmem_addr(tmem, st_lane, TM_A + (mt * 2 + 0) * H + k0), hi);
            tc::tmem_st16(tc::tmem_addr(tmem, st_lane, TM_A + (mt * 2 + 1) * H + k0), lo);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, TB, false, false);
    const int ntiles = (B + TB - 1) / TB;

    if (!compute) {
        if (lane == 0) {
            const int ii = warp - NC / 32, mt = ii >> 1, kh = ii & 1;   // this issuer's M tile (0 = gates (i, f), 1 = (g, o)) and K-step parity
            const uint32_t Ah = tmem + TM_A + (mt * 2 + 0) * H, Al = tmem + TM_A + (mt * 2 + 1) * H, D = tmem + TM_D + ii * TB;
            uint32_t n = 0;
            FWP_INIT
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int t = 0; t < T; ++t, ++n) {
                    FWP(8, NC);
                    tc::mbar_wait(b_full, n & 1);                 // h_{t-1} operand written, accumulators drained
                    tc::fence_after_sync();
                    FWP(9, NC);
#pragma unroll
                    for (int s = kh; s < H / 8; s += 2) {
                        const tc::SmemDesc bh = tc::make_desc(B_hi + s * 2 * B_LBO, B_LBO, B_SBO);
                        const tc::SmemDesc bl = tc::make_desc(B_lo + s * 2 * B_LBO, B_LBO, B_SBO);
                        tc::mma_tf32_ts(D, Ah + s * 8, bl, idesc, s > 1 ? 1u : 0u);     // (order: one operand changes per MMA)
                        tc::mma_tf32_ts(D, Ah + s * 8, bh, idesc, 1u);
                        tc::mma_tf32_ts(D, Al + s * 8, bh, idesc, 1u);
                    }
                    tc::mma_commit(acc_done);
                    FWP(10, NC);
                }
            }
        }
    } else {
        const int q = warp & 3, ch = warp >> 2;                // TMEM lane quarter, sequence group
        const bool fo = q >= 2;                                // false: gates (i, g) ; true: gates (f, o)
        const int j = (q & 1) * 32 + lane;                     // unit
        const int pair = (q & 1) + 2 * ch;
        float* EXs = EX + pair * (2 * RW * 32) + (fo ? RW * 32 : 0);        // my outgoing slots [slot][lane]
        const float* EXr = EX + pair * (2 * RW * 32) + (fo ? 0 : RW * 32);  // the partner's
        const int rbase = ch * RW;                             // the RW sequences this warp reads from tensor memory
        const int mine = fo ? CT : 0;                          // the CT of them this thread finishes (the others go to the partner)
#ifdef B200RL_EMUL
        const int addr_lane = 32 * q + lane;
#else
        const int addr_lane = 32 * q;
#endif
        auto boff = [&](int r) { return (r & 7) * 16 + (r >> 3) * B_SBO + (j & 3) * 4 + (j >> 2) * B_LBO; };
        uint32_t n = 0;
        FWP_INIT
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b0 = tile * TB;
            // the 8 cells this thread owns: sequences b0 + rbase + mine + i, unit j ; row (b, t) of the (B, T, .) tensors =
            // rowof(i) + t ; sequences past B are clamped for loads and masked for stores
            auto seq = [&](int i) { const int b = b0 + rbase + mine + i; return b < B ? b : B - 1; };
            auto rowof = [&](int i) { return (unsigned)seq(i) * (unsigned)T; };       // B*T < 2^31 (checked by the launcher)
            float cst[CT];
#pragma unroll
            for (int i = 0; i < CT; ++i) {
                const size_t b = (size_t)seq(i);
                cst[i] = cx0[b * H + j];
                float hi, lo;
                tc::split_tf32(hx0[b * H + j], hi, lo);
                const int ob = boff(rbase + mine + i);
                *reinterpret_cast<float*>(B_hi + ob) = hi;
                *reinterpret_cast<float*>(B_lo + ob) = lo;
            }
            float4 xq[CT];                                     // xproj of my cells, one step ahead
#pragma unroll
            for (int i = 0; i < CT; ++i) xq[i] = ld4(xg + ((size_t)rowof(i) * H + j) * 4);
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive_warp(b_full);

            for (int t = 0; t < T; ++t, ++n) {
                FWP(0, 0);
                tc::mbar_wait(acc_done, n & 1);
                tc::fence_after_sync();
                FWP(1, 0);
                float xa[RW], xb[RW];
                {                                                       // gate i (or f) / g (or o) of unit j, RW sequences: even + odd K-step partials
                    uint32_t u0[RW], u1[RW], u2[RW], u3[RW];            // four loads in flight, one wait
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 0 * TB + rbase), u0);
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 1 * TB + rbase), u1);
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 2 * TB + rbase), u2);
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 3 * TB + rbase), u3);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        xa[r] = __uint_as_float(u0[r]) + __uint_as_float(u1[r]);
                        xb[r] = __uint_as_float(u2[r]) + __uint_as_float(u3[r]);
                    }
                }
                FWP(2, 0);
#pragma unroll
                for (int i = 0; i < CT; ++i) { EXs[(2 * i) * 32 + lane] = fo ? xa[i] : xa[CT + i]; EXs[(2 * i + 1) * 32 + lane] = fo ? xb[i] : xb[CT + i]; }
                FWP(3, 0);
                tc::named_bar_sync(1 + pair, 64);
                FWP(4, 0);
                float4 gsv[CT]; float hsv[CT];                          // saved activations: stored to HBM AFTER the step is handed to the issuers
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const float pa = EXr[(2 * i) * 32 + lane], pb = EXr[(2 * i + 1) * 32 + lane];
                    const float4 x = xq[i];
                    if (t + 1 < T) xq[i] = ld4(xg + ((size_t)(rowof(i) + t + 1) * H + j) * 4);
                    float ig, fg, gg, og;                                   // 8 MUFU per cell: (i, g) and (f, o) share a reciprocal each, tanh(c) takes two
                    sigm_tanh((fo ? pa : xa[i]) + x.x, (fo ? pb : xb[i]) + x.z, ig, gg);
                    sigm_sigm((fo ? xa[CT + i] : pa) + x.y, (fo ? xb[CT + i] : pb) + x.w, fg, og);
                    const float cn = fmaf(fg, cst[i], ig * gg);
                    const float hh = og * tanh_(cn);
                    cst[i] = cn;
                    gsv[i] = make_float4(ig, fg, gg, og); hsv[i] = hh;
                    float hi, lo;
                    tc::split_tf32(hh, hi, lo);
                    const int ob = boff(rbase + mine + i);
                    *reinterpret_cast<float*>(B_hi + ob) = hi;
                    *reinterpret_cast<float*>(B_lo + ob) = lo;
                }
                FWP(5, 0);
                if (t + 1 < T) {
                    tc::fence_async_smem();
                    tc::fence_before_sync();
                    tc::mbar_arrive_warp(b_full);
                }
                FWP(6, 0);
#pragma unroll
                for (int i = 0; i < CT; ++i) {                          // (these stores overlap the tensor core's next step)
                    if (b0 + rbase + mine + i < B) {
                        const size_t o = (size_t)(rowof(i) + t) * H + j;
                        st4(xg + o * 4, gsv[i]);
                        c_out[o] = cst[i]; h_out[o] = hsv[i];
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NC / 32) tc::tmem_dealloc(tmem, TM_COLS);
}

int launch_fwd_tc(float* xg, const float* hx0, const float* cx0, const float* W_hh, float* c, float* h,
                  int B, int T, int sms, void* stream) {
    B200RL_REQUIRE((long long)B * T < (1LL << 31), "lstm_fwd(tc): B*T must be below 2^31");
    const int ntiles = (B + ftk::TB - 1) / ftk::TB;
    const size_t smem = fwdt_smem_bytes();
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(lstm_fwdt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lstm_fwd(tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
    B200RL_LAUNCH(lstm_fwdt_kernel, ntiles < sms ? ntiles : sms, ftk::NT, smem, stream, xg, hx0, cx0, W_hh, c, h, B, T);
    return check_launch("lstm_fwd(tc)");
}

}  // namespace b200rl
