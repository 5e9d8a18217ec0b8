// The whole forward of the actor-critic in ONE persistent tcgen05 kernel (sm_100a):
//
//   x_t   = ReLU(Wt . obs_t + bt)                                             (torso, networks/models.py:67-69)
//   pre_t = W_ih . x_t + W_hh . h_{t-1} + b_ih + b_hh ; LSTM cell             (nn.LSTMCell looped over T, models.py:71-75)
//   head  = W_head . h_t + b_head                                             (value / logits heads, models.py:77-78)
//
// Round 1 ran this as three passes (lstm_xproj_tc -> lstm_fwdt -> head_fwd) with a 1 GB HBM round trip between the first two
// (the input projection of ALL (b, t) parked in the `gates` buffer: 478 MB written, 539 MB read back) and a 134 MB re-read of h
// by the third.  Here the input projection never leaves the SM:
//
//   * W_hh AND W_ih stay resident in tensor memory as the A operands of "weights as M" GEMMs D[256 x 32] (two M = 128 tiles:
//     gates (i, f) and (g, o); N = 32 sequences): TF32 high parts 2 x 128 columns, low parts as bf16 pairs 2 x 64 columns,
//     accumulators 4 x 32 columns -- exactly the 512 columns of an SM.
//   * split precision as in the backward kernel: A_hi . B_hi + A_hi . B_lo in kind::tf32, the third term A_lo . B_hi (2^-11 of
//     the result) in kind::f16 (bf16, K = 16): 20 MMAs per (tile, operand) instead of 24.
//   * x_t is produced by dedicated producer warps, TWO steps ahead and off the critical path (4 FMAs per element from the
//     Fin-wide observation), as a K-major B operand (hi / lo / bf16) in a double-buffered shared-memory slot; dedicated issuer
//     warps run the input-projection MMAs of step t+1 into the accumulators WHILE the cell phase of step t runs (they only
//     wait for the accumulators to have been read), so the recurrent MMAs of step t+1 -- the only ones on the serial chain --
//     accumulate on top as soon as h_t is written.  The serial chain carries 40 MMAs per step instead of 48.
//   * the heads: after a thread has handed h_t to the tensor core it multiplies its 4 cells by the head rows of its unit and the
//     warp folds the 16 (sequence, output) sums with 16 shuffles; each unit-half warp stores its partial head (12.6 MB per
//     update instead of re-reading 134 MB of h); after its last step the CTA adds the two halves and the bias of its own sequences.
//   * the tile width (sequences per CTA = N of the MMAs) is a template parameter: 32, or 16 for batches that would leave more
//     than half of the SMs without a tile (`recurrent_tile16`): twice the CTAs, half the cell warps, a shorter cell phase per step.
//
// Register budget (32-sequence tiles): 16 cell + 8 issuer + 4 producer warps = 896 threads = 72 registers each at launch; the
// issuer warp-groups give theirs back (setmaxnreg.dec 40), the producers keep 64 and the cell warp-groups grow to 88
// (setmaxnreg.inc); ptxas allocates per region.  16-sequence tiles: 576 threads with 96 registers each, nothing to move.
//
// Layouts: gates (B, T, 64, 4) = post-activation (i, f, g, o) of a unit as one float4 ; c, h (B, T, 64) ; head_part
// (B, T, 2, NH) ; everything a thread stores is a full 16-byte / 128-byte-per-warp line segment.
#include <cstdlib>
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

#ifdef B200RL_FWD3PROF
__device__ unsigned long long g_fwd3prof[16];
#define F3P_INIT unsigned long long f3p_prev = clock64();
#define F3P(i, who) do { if (tid == (who)) { unsigned long long now_ = clock64(); atomicAdd(&g_fwd3prof[i], now_ - f3p_prev); f3p_prev = now_; } } while (0)
#else
#define F3P_INIT
#define F3P(i, who) do { } while (0)
#endif

namespace ffk {
constexpr int H = 64, MAXF = 8, MAXNH = 4;
constexpr int RW = 8, CT = RW / 2;                         // sequences a warp reads from tensor memory, cells a thread owns
#ifndef B200RL_FWD3_SEP
#define B200RL_FWD3_SEP 1                                  // 1: separate issuer warps for the input projection (+ setmaxnreg) ; 0: the 4 recurrent issuers do both
#endif
constexpr int NI = 4;                                      // issuers per operand: (M tile) x (even / odd K-steps), one accumulator each
constexpr int NIW = B200RL_FWD3_SEP ? 2 * NI : NI;         // issuer warps
constexpr int TM_D = 0, TM_HH = 128, TM_IH = 256, TM_HHL = 384, TM_IHL = 448, TM_COLS = 512;
constexpr int B_LBO = 144, B_SBO = 2304;                   // TF32 B operand: 16 core matrices (4 k) along K, 16-byte skew -> conflict-free 4-byte scatter
constexpr int C_LBO = 144, C_SBO = 1152;                   // bf16 B operand: 8 core matrices (8 k) along K
constexpr int NBAR = 10;                                   // b_full, acc_done, acc_free, x_done, x_full[2], x_free[2], h_full, h_free
#ifndef B200RL_FWD_HEADS_PROD
#define B200RL_FWD_HEADS_PROD 0                            // 0: heads by the cell warps (shuffle fold + partial heads) ; 1: by the producer warps from the h operand in shared
                                                           // memory -- measured SLOWER (0.309 -> 0.360 ms: the producers lose their run-ahead and their 49 KB of operand reads per
                                                           // step land in the window in which the tensor core reads the same operand); kept selectable, parity-tested in the emulator
#endif
constexpr bool HP = B200RL_FWD_HEADS_PROD != 0;
constexpr int HLAG = 3;                                    // the producers run ~3 steps ahead of the cell phase with x: the heads of step i - 3 precede x of step i
static_assert(15 * B_LBO + 128 <= B_SBO && 7 * C_LBO + 128 <= C_SBO, "core matrices of one 8-row group must fit below SBO");
static_assert(RW == 8 && CT == 4, "cell phase layout");
// Everything that depends on the tile width TB (sequences per CTA = the N of every MMA): 32 for full batches, 16 when the
// batch has too few 32-sequence tiles to fill half the SMs (B = 2048: 128 CTAs of 16 sequences instead of 64 of 32 -- the
// kernel is bound by the per-step chain, and half the cells per CTA shorten the cell phase of every step)
template <int TB>
struct Tile {
    static_assert(TB == 32 || TB == 16, "tile width");
    static constexpr int NG = TB / RW;                      // sequence groups = cell warps per TMEM lane quarter
    static constexpr int NPW = TB / 8;                      // producer warps: the torso output x_t as a K-major operand, two steps ahead (thread = unit x 16 sequences)
    static constexpr int NC = NG * 128, NT = NC + 32 * NIW + 32 * NPW;
    static constexpr int B_BYTES = (TB / 8) * B_SBO;        // 9216 per hi / lo at TB = 32
    static constexpr int C_BYTES = (TB / 8) * C_SBO;        // 4608
    static constexpr int OP_BYTES = 2 * B_BYTES + C_BYTES;  // one operand set: hi | lo | bf16(hi)
    static constexpr int EXF = 2 * NG * 2 * RW * 32;        // swap buffer [pair][direction][slot][lane] floats
    static constexpr size_t smem_bytes() { return 3 * (size_t)OP_BYTES + sizeof(float) * (EXF + H * 4 + H * MAXF + H + H * 4) + 8 * NBAR + 64; }
};
}  // namespace ffk

template <bool F4, int TB>
__global__ void __launch_bounds__(ffk::Tile<TB>::NT, 1)
lstm_fwd_fused_kernel(const float* __restrict__ u, const float* __restrict__ hx0, const float* __restrict__ cx0,
                      const float* __restrict__ Wt, const float* __restrict__ bt, const float* __restrict__ W_ih,
                      const float* __restrict__ W_hh, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                      const float* __restrict__ W_head, float* __restrict__ g_out, float* __restrict__ c_out,
                      float* __restrict__ h_out, float* __restrict__ head_part, const float* __restrict__ b_head,
                      float* __restrict__ head_out, int B, int T, int Fin, int NH) {
    using namespace ffk;
    using TL = Tile<TB>;
    constexpr int NPW = TL::NPW, NC = TL::NC, NT = TL::NT, B_BYTES = TL::B_BYTES, OP_BYTES = TL::OP_BYTES, EXF = TL::EXF;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* OPH = smem;                              // h_{t-1}: hi | lo | bf16
    unsigned char* OPX = smem + OP_BYTES;                   // x_t: two sets
    float* EX = reinterpret_cast<float*>(smem + 3 * OP_BYTES);
    float* BI = EX + EXF;                                   // BI[j*4 + g] = b_ih[g*64+j] + b_hh[g*64+j]
    float* WT = BI + H * 4;                                 // WT[j*8 + f]
    float* BT = WT + H * MAXF;                              // bt[j]
    float* WH = BT + H;                                     // WH[j*4 + o]
    uint64_t* bars = reinterpret_cast<uint64_t*>(WH + H * 4);
    uint64_t* b_full = bars, *acc_done = bars + 1, *acc_free = bars + 2, *x_done = bars + 3, *x_full = bars + 4, *x_free = bars + 6;   // x_full[2], x_free[2]
    uint64_t* h_full = bars + 8, *h_free = bars + 9;        // (HP) h_t sits in OPH ; the producers have read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool compute = warp < NC / 32;
    const bool producer = warp >= NC / 32 + NIW;

    for (int i = tid; i < H * 4; i += NT) { const int j = i >> 2, g = i & 3; BI[i] = b_ih[g * H + j] + b_hh[g * H + j]; }
    for (int i = tid; i < H * MAXF; i += NT) { const int j = i >> 3, f = i & 7; WT[i] = f < Fin ? Wt[j * Fin + f] : 0.f; }
    for (int i = tid; i < H; i += NT) BT[i] = bt[i];
    const bool heads = head_part != nullptr;
    if (HP) { for (int i = tid; i < H * 4; i += NT) { const int o = i >> 6, j = i & 63; WH[i] = (heads && o < NH) ? W_head[o * H + j] : 0.f; } }   // WH[o*64 + j]
    else { for (int i = tid; i < H * 4; i += NT) { const int j = i >> 2, o = i & 3; WH[i] = (heads && o < NH) ? W_head[o * H + j] : 0.f; } }       // WH[j*4 + o]
    if (tid == 0) {
        tc::mbar_init(b_full, NC / 32); tc::mbar_init(acc_done, NI); tc::mbar_init(acc_free, NC / 32); tc::mbar_init(x_done, NI);
        tc::mbar_init(x_full, NPW); tc::mbar_init(x_full + 1, NPW);
        tc::mbar_init(x_free, NI); tc::mbar_init(x_free + 1, NI);
        tc::mbar_init(h_full, NC / 32); tc::mbar_init(h_free, NPW);
        tc::fence_mbar_init();
    }
    if (warp == NC / 32) tc::tmem_alloc(tmem_slot, TM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    // one-time: W_hh and W_ih (PyTorch row order n = g*64 + j) -> tensor memory: thread = row n % 128 (its own TMEM lane),
    // TF32 high parts one column per k, low parts as bf16 pairs (column c = k 2c | k 2c+1 << 16)
    if (warp < 16) {
        const int which = warp >> 3, mt = (warp >> 2) & 1, m = 32 * (warp & 3) + lane;
        const float* wrow = (which ? W_ih : W_hh) + (size_t)(mt * 128 + m) * H;
#ifdef B200RL_EMUL
        const int st_lane = m;
#else
        const int st_lane = 32 * (warp & 3);
#endif
        for (int k0 = 0; k0 < H; k0 += 32) {
            float lo16[16];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                    const float4 w = ld4(wrow + k0 + hh * 16 + e);
                    tc::split_tf32(w.x, hi[e], lo[e]); tc::split_tf32(w.y, hi[e + 1], lo[e + 1]);
                    tc::split_tf32(w.z, hi[e + 2], lo[e + 2]); tc::split_tf32(w.w, hi[e + 3], lo[e + 3]);
                }
                tc::tmem_st16(tc::tmem_addr(tmem, st_lane, (which ? TM_IH : TM_HH) + mt * H + k0 + hh * 16), hi);
#pragma unroll
                for (int e = 0; e < 16; e += 2)
                    lo16[hh * 8 + e / 2] = __uint_as_float((unsigned)tc::bf16_bits_rn(lo[e]) | ((unsigned)tc::bf16_bits_rn(lo[e + 1]) << 16));
            }
            tc::tmem_st16(tc::tmem_addr(tmem, st_lane, (which ? TM_IHL : TM_HHL) + mt * (H / 2) + k0 / 2), lo16);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, TB, false, false), idesc16 = tc::make_idesc_bf16(128, TB);
    const int ntiles = (B + TB - 1) / TB;
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long nsteps = (long long)my_tiles * T;

    // 10 MMAs of one (M tile, K parity) share of D (+)= A . OP^T: A_hi . OP_lo, A_hi . OP_hi over 4 K-steps of 8, A_lo . OP_hi(bf16) over 2 of 16
    auto issue = [&](uint32_t D, uint32_t Ahi, uint32_t Alo, const unsigned char* op, int kh, bool fresh) {
#pragma unroll
        for (int s = kh; s < H / 8; s += 2) {
            const tc::SmemDesc bh = tc::make_desc(op + s * 2 * B_LBO, B_LBO, B_SBO);
            const tc::SmemDesc bl = tc::make_desc(op + B_BYTES + s * 2 * B_LBO, B_LBO, B_SBO);
            tc::mma_tf32_ts(D, Ahi + s * 8, bl, idesc, (fresh && s == kh) ? 0u : 1u);
            tc::mma_tf32_ts(D, Ahi + s * 8, bh, idesc, 1u);
        }
#pragma unroll
        for (int s = kh; s < H / 16; s += 2)
            tc::mma_bf16_ts(D, Alo + s * 8, tc::make_desc(op + 2 * B_BYTES + s * 2 * C_LBO, C_LBO, C_SBO), idesc16, 1u);
    };

    if (producer) {
#ifndef B200RL_EMUL
        if constexpr (F4 && TB == 32) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");   // (generic torso widths keep the 72 registers of the launch: the 8192 the cell warps gain are exactly what
                                                                                            // the issuers give back ; the 576 threads of a 16-sequence tile start with 96 each: nothing to move)
#endif
        // x[r][j] = ReLU(bt[j] + Wt[j] . obs[r, t]) for the tile's 32 sequences: thread = (unit j, 16 of the sequences), element
        // (r, j) of operand set t & 1 as TF32 hi / lo and bf16(hi)
        const int pt = tid - (NC + 32 * NIW), j = pt & 63, r0 = (pt >> 6) * 16;
        auto tf = [&](int r) { return (r & 7) * 16 + (r >> 3) * B_SBO + (j & 3) * 4 + (j >> 2) * B_LBO; };
        auto bf = [&](int r) { return 2 * B_BYTES + (r & 7) * 16 + (r >> 3) * C_SBO + (j & 7) * 2 + (j >> 3) * C_LBO; };
        const float btj = BT[j];
        float wt[F4 ? 4 : 1];                                   // (wider / narrower torsos read their weights from shared memory in the loop: registers)
#pragma unroll
        for (int f = 0; f < (F4 ? 4 : 1); ++f) wt[f] = WT[j * MAXF + f];
        uint32_t written[2] = {0, 0};
        // (HP) heads of one step from the h operand the cell warps just wrote: thread = (sequence r = 8 * warp + (lane & 7), unit
        // quarter lane >> 3): h = hi + lo of 16 units as four 16-byte reads each (a quarter-warp reads 128 contiguous bytes),
        // 4 x 16 FMAs against W_head^T, two shuffle-adds over the quarters, + bias -> head_out (no partial heads, no fold)
        const int hw = pt >> 5, hrr = lane & 7, hq = lane >> 3, hr = 8 * hw + hrr;
        long long hn = 0;                                       // head steps done
        int htile = blockIdx.x, ht = 0;
        auto do_head = [&]() {
            tc::mbar_wait(h_full, (uint32_t)hn & 1);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const unsigned char* base = OPH + hrr * 16 + hw * B_SBO + (hq * 4) * B_LBO;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 vh = *reinterpret_cast<const float4*>(base + q * B_LBO);
                const float4 vl = *reinterpret_cast<const float4*>(base + B_BYTES + q * B_LBO);
                const float4 hv = make_float4(vh.x + vl.x, vh.y + vl.y, vh.z + vl.z, vh.w + vl.w);
                const int j0 = hq * 16 + q * 4;
                const float4 w0 = ld4(WH + j0), w1 = ld4(WH + H + j0), w2 = ld4(WH + 2 * H + j0), w3 = ld4(WH + 3 * H + j0);
                a0 = fmaf(hv.x, w0.x, fmaf(hv.y, w0.y, fmaf(hv.z, w0.z, fmaf(hv.w, w0.w, a0))));
                a1 = fmaf(hv.x, w1.x, fmaf(hv.y, w1.y, fmaf(hv.z, w1.z, fmaf(hv.w, w1.w, a1))));
                a2 = fmaf(hv.x, w2.x, fmaf(hv.y, w2.y, fmaf(hv.z, w2.z, fmaf(hv.w, w2.w, a2))));
                a3 = fmaf(hv.x, w3.x, fmaf(hv.y, w3.y, fmaf(hv.z, w3.z, fmaf(hv.w, w3.w, a3))));
            }
            tc::mbar_arrive_warp(h_free);                       // (every lane has read its part of the operand: the cell warps may overwrite it)
#pragma unroll
            for (int m = 8; m <= 16; m <<= 1) {
                a0 += __shfl_xor_sync(kFull, a0, m); a1 += __shfl_xor_sync(kFull, a1, m);
                a2 += __shfl_xor_sync(kFull, a2, m); a3 += __shfl_xor_sync(kFull, a3, m);
            }
            const int b = htile * TB + hr;
            if (hq == 0 && b < B) {
                float* o = head_out + ((size_t)b * T + ht) * NH;
                o[0] = a0 + __ldg(b_head);                              // (8 lanes per warp and step; the bias stays in L1)
                if (NH > 1) o[1] = a1 + __ldg(b_head + 1);
                if (NH > 2) o[2] = a2 + __ldg(b_head + 2);
                if (NH > 3) o[3] = a3 + __ldg(b_head + 3);
            }
            ++hn;
            if (++ht == T) { ht = 0; htile += gridDim.x; }
        };
        long long xn = 0;                                       // x steps produced
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b0 = tile * TB;
            for (int t = 0; t < T; ++t) {
                const int p = t & 1;
                unsigned char* op = OPX + p * OP_BYTES;
                // heads FIRST: the cell phase of step i - 2 waits for them before it frees the accumulators that x_i's operand set
                // may be waiting on (with an odd T two consecutive steps share a set across a tile boundary)
                if (HP && heads && xn >= HLAG) do_head();
                ++xn;
                float xv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {                          // (the loads of a step do not wait for the operand set)
                    const int b = b0 + r0 + i;
                    const size_t row = (size_t)(b < B ? b : B - 1) * T + t;
                    float x = btj;
                    if (F4) {
                        const float4 o = ld4(u + row * 4);
                        x = fmaf(wt[0], o.x, fmaf(wt[1], o.y, fmaf(wt[2], o.z, fmaf(wt[3], o.w, x))));
                    } else {
#pragma unroll
                        for (int f = 0; f < MAXF; ++f) if (f < Fin) x = fmaf(WT[j * MAXF + f], __ldg(u + row * Fin + f), x);
                    }
                    xv[i] = fmaxf(x, 0.f);
                }
                if (written[p] > 0) tc::mbar_wait(x_free + p, (written[p] - 1) & 1);     // the MMAs that last read this set are done
                ++written[p];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float hi, lo;
                    tc::split_tf32(xv[i], hi, lo);
                    *reinterpret_cast<float*>(op + tf(r0 + i)) = hi;
                    *reinterpret_cast<float*>(op + B_BYTES + tf(r0 + i)) = lo;
                    *reinterpret_cast<unsigned short*>(op + bf(r0 + i)) = tc::bf16_bits_rn(hi);
                }
                tc::fence_async_smem();
                tc::mbar_arrive_warp(x_full + p);
            }
        }
        if (HP && heads) while (hn < nsteps) do_head();         // the last HLAG steps
    } else if (!compute) {
#if !defined(B200RL_EMUL)
        if constexpr (TB == 32) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
#endif
        if (lane == 0) {
            const int iw = warp - NC / 32;
            const int ii = iw & (NI - 1), mt = ii >> 1, kh = ii & 1;
            const uint32_t D = tmem + TM_D + ii * TB;
            const uint32_t Ahh = tmem + TM_HH + mt * H, Ahhl = tmem + TM_HHL + mt * (H / 2);
            const uint32_t Aih = tmem + TM_IH + mt * H, Aihl = tmem + TM_IHL + mt * (H / 2);
#if B200RL_FWD3_SEP
            F3P_INIT
            if (iw >= NI) {                                  // input-projection issuer: runs one step ahead of the cell phase
                uint32_t xc[2] = {0, 0};
                long long N = 0;
                for (int tl = 0; tl < my_tiles; ++tl)
                    for (int t = 0; t < T; ++t, ++N) {
                        const int p = t & 1;
                        F3P(11, NC + 32 * NI);
                        tc::mbar_wait(x_full + p, xc[p] & 1); ++xc[p];
                        F3P(12, NC + 32 * NI);
                        if (N > 0) tc::mbar_wait(acc_free, (uint32_t)(N - 1) & 1);
                        tc::fence_after_sync();
                        F3P(13, NC + 32 * NI);
                        issue(D, Aih, Aihl, OPX + p * OP_BYTES, kh, true);
                        tc::mma_commit(x_done);
                        tc::mma_commit(x_free + p);          // (the producers may refill this operand set once these MMAs have read it)
                        F3P(14, NC + 32 * NI);
                    }
            } else {
                for (long long N = 0; N < nsteps; ++N) {
                    F3P(7, NC);
                    tc::mbar_wait(b_full, (uint32_t)N & 1);  // h_{t-1} operand written
                    F3P(8, NC);
                    tc::mbar_wait(x_done, (uint32_t)N & 1);  // the input projection of this step sits in the accumulator
                    tc::fence_after_sync();
                    F3P(9, NC);
                    issue(D, Ahh, Ahhl, OPH, kh, false);
                    tc::mma_commit(acc_done);
                    F3P(10, NC);
                }
            }
#else
            uint32_t xc[2] = {0, 0};
            long long N = 0;
            for (int tl = 0; tl < my_tiles; ++tl)
                for (int t = 0; t < T; ++t, ++N) {
                    const int p = t & 1;
                    tc::mbar_wait(x_full + p, xc[p] & 1); ++xc[p];
                    if (N > 0) tc::mbar_wait(acc_free, (uint32_t)(N - 1) & 1);
                    tc::fence_after_sync();
                    issue(D, Aih, Aihl, OPX + p * OP_BYTES, kh, true);
                    tc::mma_commit(x_free + p);
                    tc::mbar_wait(b_full, (uint32_t)N & 1);
                    tc::fence_after_sync();
                    issue(D, Ahh, Ahhl, OPH, kh, false);
                    tc::mma_commit(acc_done);
                }
#endif
        }
    } else {
#if !defined(B200RL_EMUL)
        if constexpr (TB == 32) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
#endif
        const int q = warp & 3, ch = warp >> 2;                // TMEM lane quarter, sequence group
        const bool fo = q >= 2;                                // false: gates (i, g) ; true: gates (f, o)
        const int half = q & 1;                                // which 32 of the 64 units
        const int j = half * 32 + lane;                        // unit
        const int pair = half + 2 * ch;
        float* EXs = EX + pair * (2 * RW * 32) + (fo ? RW * 32 : 0);        // my outgoing slots [slot][lane]
        const float* EXr = EX + pair * (2 * RW * 32) + (fo ? 0 : RW * 32);  // the partner's
        const int rbase = ch * RW;                             // the RW sequences this warp reads from tensor memory
        const int mine = fo ? CT : 0;                          // the CT of them this thread finishes
#ifdef B200RL_EMUL
        const int addr_lane = 32 * q + lane;
#else
        const int addr_lane = 32 * q;
#endif
        // element (sequence r, unit j) of a K-major operand set: TF32 hi/lo at tf(r), bf16 at bf(r)
        auto tf = [&](int r) { return (r & 7) * 16 + (r >> 3) * B_SBO + (j & 3) * 4 + (j >> 2) * B_LBO; };
        auto bf = [&](int r) { return 2 * B_BYTES + (r & 7) * 16 + (r >> 3) * C_SBO + (j & 7) * 2 + (j >> 3) * C_LBO; };
        auto put = [&](unsigned char* op, int r, float v) {
            float hi, lo;
            tc::split_tf32(v, hi, lo);
            *reinterpret_cast<float*>(op + tf(r)) = hi;
            *reinterpret_cast<float*>(op + B_BYTES + tf(r)) = lo;
            *reinterpret_cast<unsigned short*>(op + bf(r)) = tc::bf16_bits_rn(hi);
        };
        const float4 bias = ld4(BI + j * 4);
        uint32_t n = 0;
        F3P_INIT
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b0 = tile * TB;
            auto seq = [&](int i) { const int b = b0 + rbase + mine + i; return b < B ? b : B - 1; };
            auto rowof = [&](int i) { return (unsigned)seq(i) * (unsigned)T; };       // B*T < 2^31 (checked by the launcher)
            float cst[CT];
            if (HP && heads && n > 0) tc::mbar_wait(h_free, (n - 1) & 1);     // the heads of the previous tile's last step have been read out of OPH
#pragma unroll
            for (int i = 0; i < CT; ++i) {
                const size_t b = (size_t)seq(i);
                cst[i] = cx0[b * H + j];
                put(OPH, rbase + mine + i, hx0[b * H + j]);
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive_warp(b_full);

            for (int t = 0; t < T; ++t, ++n) {
                F3P(0, 0);
                if (HP && heads && t > 0) tc::mbar_wait(h_free, (n - 1) & 1);  // (long done: the producers read h_{t-1} while the tensor core ran this step)
                tc::mbar_wait(acc_done, n & 1);
                tc::fence_after_sync();
                F3P(1, 0);
                float xa[RW], xb[RW];
                {                                                       // gate i (or f) / g (or o) of unit j, RW sequences: even + odd K-step partials
                    uint32_t u0[RW], u1[RW], u2[RW], u3[RW];            // four loads in flight, one wait
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 0 * TB + rbase), u0);
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 1 * TB + rbase), u1);
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 2 * TB + rbase), u2);
                    tc::tmem_ld8_async(tc::tmem_addr(tmem, addr_lane, TM_D + 3 * TB + rbase), u3);
                    tc::tmem_ld_wait();
                    tc::fence_before_sync();
                    tc::mbar_arrive_warp(acc_free);                     // the accumulators may take the next step's input projection
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        xa[r] = __uint_as_float(u0[r]) + __uint_as_float(u1[r]);
                        xb[r] = __uint_as_float(u2[r]) + __uint_as_float(u3[r]);
                    }
                }
                F3P(2, 0);
#pragma unroll
                for (int i = 0; i < CT; ++i) { EXs[(2 * i) * 32 + lane] = fo ? xa[i] : xa[CT + i]; EXs[(2 * i + 1) * 32 + lane] = fo ? xb[i] : xb[CT + i]; }
                tc::named_bar_sync(1 + pair, 64);
                F3P(3, 0);
                float4 gsv[CT]; float hsv[CT];                          // saved activations: stored to HBM AFTER the step is handed to the issuers
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const float pa = EXr[(2 * i) * 32 + lane], pb = EXr[(2 * i + 1) * 32 + lane];
                    float ig, fg, gg, og;                                   // 8 MUFU per cell: (i, g) and (f, o) share a reciprocal each, tanh(c) takes two
                    sigm_tanh((fo ? pa : xa[i]) + bias.x, (fo ? pb : xb[i]) + bias.z, ig, gg);
                    sigm_sigm((fo ? xa[CT + i] : pa) + bias.y, (fo ? xb[CT + i] : pb) + bias.w, fg, og);
                    const float cn = fmaf(fg, cst[i], ig * gg);
                    const float hh = og * tanh_(cn);
                    cst[i] = cn;
                    gsv[i] = make_float4(ig, fg, gg, og); hsv[i] = hh;
                    put(OPH, rbase + mine + i, hh);
                }
                F3P(4, 0);
                if (t + 1 < T) {
                    tc::fence_async_smem();
                    tc::fence_before_sync();
                    tc::mbar_arrive_warp(b_full);
                }
                if (HP && heads) tc::mbar_arrive_warp(h_full);          // h_t is in shared memory (every step, also the last of a tile)
                F3P(5, 0);
                // ---- off the serial chain from here: the tensor core is running step t+1
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    if (b0 + rbase + mine + i < B) {
                        const size_t o = (size_t)(rowof(i) + t) * H + j;
                        st4(g_out + o * 4, gsv[i]);
                        c_out[o] = cst[i]; h_out[o] = hsv[i];
                    }
                }
                if (!HP && head_part != nullptr) {
                    // heads: v[i*4 + o] = h[seq i][j] * W_head[o][j], folded over the warp's 32 units: lane l ends with slot (l >> 1) & 15
                    const float4 wh = ld4(WH + j * 4);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < CT; ++i) { v[i * 4] = hsv[i] * wh.x; v[i * 4 + 1] = hsv[i] * wh.y; v[i * 4 + 2] = hsv[i] * wh.z; v[i * 4 + 3] = hsv[i] * wh.w; }
#pragma unroll
                    for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0;
#pragma unroll
                        for (int k = 0; k < w; ++k) {
                            const float send = up ? v[k] : v[k + w], keep = up ? v[k + w] : v[k];
                            v[k] = keep + __shfl_xor_sync(kFull, send, bit);
                        }
                    }
                    v[0] += __shfl_xor_sync(kFull, v[0], 1);
                    const int slot = (lane >> 1) & 15, si = slot >> 2, so = slot & 3;
                    if ((lane & 1) == 0 && so < NH && b0 + rbase + mine + si < B)
                        head_part[((size_t)(rowof(si) + t) * 2 + half) * NH + so] = v[0];
                }
                F3P(6, 0);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NC / 32) tc::tmem_dealloc(tmem, TM_COLS);
    // ---- heads of the CTA's own sequences: head_out[r][o] = (part[r][0][o] + part[r][1][o]) + b_head[o].  The partial sums were
    // written by this CTA (visible after the barrier, served by L2): all 28 warps fold them in ~2 us instead of a separate kernel.
    if (!HP && head_part != nullptr && head_out != nullptr) {
        for (int tile = blockIdx.x; tile < (B + TB - 1) / TB; tile += gridDim.x) {
            const int b0 = tile * TB, nb = (B - b0 < TB) ? B - b0 : TB;
            const long long base = (long long)b0 * T * NH, n = (long long)nb * T * NH;
            for (long long i = threadIdx.x; i < n; i += blockDim.x) {
                const long long r = (base + i) / NH;
                const int o = (int)((base + i) - r * NH);
                head_out[base + i] = head_part[(r * 2) * NH + o] + head_part[(r * 2 + 1) * NH + o] + b_head[o];
            }
        }
    }
}

template <bool F4, int TB>
static int launch_fwd_fused_t(const float* u, const float* hx0, const float* cx0, const float* Wt, const float* bt, const float* W_ih,
                              const float* W_hh, const float* b_ih, const float* b_hh, const float* W_head, const float* b_head,
                              float* gates, float* c, float* h, float* head_part, float* head_out,
                              int B, int T, int Fin, int NH, int sms, void* stream) {
    const int ntiles = (B + TB - 1) / TB;
    const size_t smem = ffk::Tile<TB>::smem_bytes();
    const int grid = ntiles < sms ? ntiles : sms;
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(lstm_fwd_fused_kernel<F4, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lstm_fwd_fused: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
    B200RL_LAUNCH((lstm_fwd_fused_kernel<F4, TB>), grid, ffk::Tile<TB>::NT, smem, stream, u, hx0, cx0, Wt, bt, W_ih, W_hh, b_ih, b_hh, W_head,
                  gates, c, h, head_part, b_head, head_out, B, T, Fin, NH);
    return check_launch("lstm_fwd_fused");
}

// 16-sequence tiles when the batch would leave more than half of the SMs without a 32-sequence tile
bool recurrent_tile16(int B, int sms) {
    static const int mode = [] { const char* e = getenv("B200RL_TILE16"); return e ? atoi(e) : -1; }();    // 0: never, 1: always (A/B), unset: by batch size
    if (mode == 0) return false;
    if (mode == 1) return true;
    return (B + 31) / 32 <= sms / 2;
}

int launch_fwd_fused(const float* u, const float* hx0, const float* cx0, const float* Wt, const float* bt, const float* W_ih,
                     const float* W_hh, const float* b_ih, const float* b_hh, const float* W_head, const float* b_head,
                     float* gates, float* c, float* h, float* head_part, float* head_out,
                     int B, int T, int Fin, int NH, int sms, void* stream) {
    B200RL_REQUIRE((long long)B * T < (1LL << 31), "lstm_fwd_fused: B*T must be below 2^31");
    B200RL_REQUIRE(Fin >= 1 && Fin <= ffk::MAXF, "lstm_fwd_fused: Fin=%d out of range [1,%d]", Fin, ffk::MAXF);
    B200RL_REQUIRE(head_part == nullptr || (NH >= 1 && NH <= ffk::MAXNH && head_out != nullptr), "lstm_fwd_fused: fused heads need 1 <= NH <= %d", ffk::MAXNH);
#define B200RL_FF_ARGS u, hx0, cx0, Wt, bt, W_ih, W_hh, b_ih, b_hh, W_head, b_head, gates, c, h, head_part, head_out, B, T, Fin, NH, sms, stream
    if (recurrent_tile16(B, sms)) return Fin == 4 ? launch_fwd_fused_t<true, 16>(B200RL_FF_ARGS) : launch_fwd_fused_t<false, 16>(B200RL_FF_ARGS);
    return Fin == 4 ? launch_fwd_fused_t<true, 32>(B200RL_FF_ARGS) : launch_fwd_fused_t<false, 32>(B200RL_FF_ARGS);
#undef B200RL_FF_ARGS
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_lstm_fwd_fused(const float* u, const float* hx0, const float* cx0,
                                     const float* Wt, const float* bt, const float* W_ih, const float* W_hh,
                                     const float* b_ih, const float* b_hh, const float* W_head, const float* b_head,
                                     float* gates, float* c, float* h, float* head_part, float* head_out,
                                     int B, int T, int Fin, int NH, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1, "lstm_fwd_fused: bad shape B=%d T=%d", B, T);
    int sms = 148;
#ifndef B200RL_EMUL
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#else
    sms = 4;
#endif
    return launch_fwd_fused(u, hx0, cx0, Wt, bt, W_ih, W_hh, b_ih, b_hh, W_head, b_head, gates, c, h, head_part, head_out,
                            B, T, Fin, NH, sms, stream);
}
