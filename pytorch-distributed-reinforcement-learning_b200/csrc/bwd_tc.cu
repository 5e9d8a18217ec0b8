// Serial half of the LSTM backward (BPTT) on tcgen05 (sm_100a).
//
// Per step t = T-1 .. 0 (autograd of nn.LSTMCell looped over T, networks/models.py:71-75, replayed by
// loss.backward(), ppo/learning.py:92):
//   dh   = dh_rec + dhead_t . W_head                         tc = tanh(c_t)
//   d_o  = dh tc ;  dct = dh o (1 - tc^2) + dc ;  d_i = dct g ;  d_g = dct i ;  d_f = dct c_{t-1} ;  dc = dct f
//   dpre = (d_i i(1-i), d_f f(1-f), d_g (1-g^2), d_o o(1-o))                    -> `gates`, in place
//   dh_rec[r][k] = sum_p dpre[r][p] W_hh[n(p)][k]           dx[r][k] = sum_p dpre[r][p] W_ih[n(p)][k]
//
// The two GEMMs of a step are ONE tcgen05 accumulation with the weights as M:
//   D[128 x 32] = [W_hh^T ; W_ih^T][128 x 256] . dpre_t^T[256 x 32]     (lanes 0-63: dh_rec, lanes 64-127: dx)
// so the torso gradient's GEMM (a separate 537 MB pass over dpre before) rides along for free.
// Split precision: A_hi (TF32) lives in tensor memory (256 columns) and meets dpre_hi and dpre_lo (TF32, K-major B
// operands in shared memory, rewritten every step by the cell phase) in two kind::tf32 MMAs per K-step; the third term
// A_lo . dpre_hi is ~2^-11 of the result, so it runs in bf16 (kind::f16, K = 16: half the MMAs): A_lo as bf16 pairs
// is 128 more TMEM columns and dpre_hi gets a bf16 copy.  80 MMAs per step; one thread gets a small tcgen05.mma out only
// every ~60-120 cycles while issues from different threads overlap (tools/tc_probe3.cu), so TWO issuer warps split K
// into halves with an accumulator each and the readers sum the partials (1 / 2 / 3 / 4 issuers: 0.68 / 0.43 / 0.50 / 0.59 ms --
// every extra accumulator pushes 32 more K columns of A_lo out of tensor memory into SS-form MMAs).
//
// Cell phase -- SYMMETRIC since round 2.  TMEM lanes are bound to the warp's quarter: dh_rec (lanes 0-63) is only visible
// to warps with id % 4 in {0, 1} ("A"), dx (lanes 64-127) only to {2, 3} ("B"), and the two kinds sit on different SM
// sub-partitions.  Round 1 gave the B warps ("helpers") ALL the dh-independent work of a step for the 8 sequences of a
// group (loads, tanh, seven coefficients per cell through shared memory) and measured them as the bottleneck: 6.7K
// cycles per step of helper instructions while the A warps idled ~4.7K waiting (tools/bwd_prof.cu).  Now thread A(unit j)
// and thread B(unit j) of a group each OWN 4 of its 8 sequences end to end: they prefetch their cells' saved activations
// and dhead rows, turn them into seven per-cell coefficients IN REGISTERS while the tensor core runs the previous step,
// and after the accumulator lands do 7 flops per cell.  The only exchange is dh_rec of B's four cells (A reads it from
// tensor memory, hands it over through 512 bytes of shared memory); B also drains dx.  Half the instructions per
// thread, no coefficient traffic through shared memory, both sub-partition pairs equally loaded.
//
// Fused small gradients (template NHT > 0, i.e. NH <= 4 known at compile time): every cell thread also accumulates
//   dW_head[o][j] += dhead[b,t,o] h[b,t,j] (h = o tanh(c) is recomputed), db_head[o] += dhead[b,t,o]
// and, with `fz.torso`, the B threads
//   dWt[j][f] += dxm obs[b,t,f], dbt[j] += dxm, dxm = dx[b,t,j] [x[b,t,j] > 0]   (dx straight from the accumulator)
// in registers across all T steps; ONE slab row per CTA is written at the end.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

// phase timeline for tools/bwd_prof.cu (never compiled into the product library)
#ifdef B200RL_BWDPROF
__device__ unsigned long long g_bwdprof[24];
#define BWP_INIT unsigned long long bwp_prev = clock64();
#define BWP(i, who) do { if (tid == (who)) { unsigned long long now_ = clock64(); atomicAdd(&g_bwdprof[i], now_ - bwp_prev); bwp_prev = now_; } } while (0)
#else
#define BWP_INIT
#define BWP(i, who) do { } while (0)
#endif

namespace btk {
#ifndef B200RL_BWD_NI
#define B200RL_BWD_NI 2                                   // issuer warps of the instantiations without fused torso gradients
#endif
#ifndef B200RL_BWD_NI_TORSO
#define B200RL_BWD_NI_TORSO 4                             // ... of the fully fused ones (measured, B = 4096, T = 128, NH = 3: 0.479 -> 0.459 ms; without the fusion 2 is faster)
#endif
#ifndef B200RL_BWD_CELL_REGS
#define B200RL_BWD_CELL_REGS 104                          // (4 issuers) the CTA owns NT x 96 registers from launch: 512 x 104 + 128 x 40 <= 640 x 96
#endif
constexpr int H = 64, G = 256;
// tile width TB = sequences per CTA = N of every MMA: 32, or 16 when 32-sequence tiles would leave more than half of the SMs idle
// (recurrent_tile16, launchers.cuh); NC = 4 lane quarters x TB / 8 sequence groups cell warps (A + B), + NI MMA issuer warps
constexpr int ncell(int tb) { return (tb / 8) * 128; }
constexpr int nthreads(int ni, int tb) { return ncell(tb) + 32 * ni; }
constexpr int B_LBO = 144, B_SBO = 9216;                            // B: 64 core matrices along K, 16-byte skew -> conflict-free 16-byte stores ; (TB / 8) * B_SBO bytes per hi / lo
static_assert(63 * B_LBO + 128 <= B_SBO, "core matrices of one 8-row group must fit below SBO");
constexpr int C_LBO = 144, C_SBO = 4608;                            // bf16 copy of dpre_hi (third split term): 32 core matrices (8 k) along K, same skew ; (TB / 8) * C_SBO bytes
static_assert(31 * C_LBO + 128 <= C_SBO, "core matrices of one 8-row group must fit below SBO");
constexpr int TM_D = 0, TM_COLS = 512;                              // accumulator i at columns TB i ; then A_hi (G columns) ; then A_lo: bf16 pairs, G/2 columns
static_assert(4 * 32 + G + G / 2 <= TM_COLS, "tensor memory budget (up to 4 accumulators)");
static_assert(ncell(32) * B200RL_BWD_CELL_REGS + 128 * 40 <= nthreads(4, 32) * 96, "setmaxnreg.inc would wait forever: the CTA's register pool is what it was launched with");
constexpr int CT = 4;                                               // cells (sequences) a thread owns
constexpr int RED = 4 * 10 * H;                                     // end-of-kernel reduction scratch: [sequence group][10][unit]
}  // namespace btk

static inline size_t bwdt_smem_bytes(int NH, int TB) {
    return 2 * (size_t)(TB / 8) * btk::B_SBO + (TB / 8) * btk::C_SBO
         + sizeof(float) * (size_t)(NH * btk::H + 16 * btk::CT * NH + 8 * btk::CT * 32 + 8 * 32 + btk::H * 4 + btk::H + btk::RED) + 64;
}

struct BwdFuse {                 // fused head / torso gradients (null partials = off)
    const float* u; const float* Wt; const float* bt; float* partials;
    long long P, off_Wt, off_bt, off_Whead, off_bhead;
    int torso;                   // 1: dWt / dbt accumulated here (dx never stored) ; 0: only the head gradients, dx is stored
};

template <int NHT, int NI, int TB> // NHT 0: runtime NH, small gradients left to their own kernels ; 1..4: NH == NHT at compile time, `fz` gradients fused ; NI: issuer warps ; TB: tile width
__global__ void __launch_bounds__(btk::nthreads(NI, TB), 1)
lstm_bwdt_kernel(const float* __restrict__ dhead, const float* __restrict__ W_head, const float* __restrict__ W_hh,
                 const float* __restrict__ W_ih, float* gates, const float* __restrict__ c_sav,
                 const float* __restrict__ cx0, float* __restrict__ dx, int B, int T, int NH, BwdFuse fz) {
    using namespace btk;
    static_assert(TB == 32 || TB == 16, "tile width");
    constexpr int NC = (TB / 8) * 128, NG = TB / 8, NT = NC + 32 * NI, TM_AH = NI * TB, TM_AL = TM_AH + G;
    constexpr int B_BYTES = (TB / 8) * B_SBO, C_BYTES = (TB / 8) * C_SBO;
    constexpr bool fuse = NHT > 0;
    if (fuse) NH = NHT;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* B_hi = smem;
    unsigned char* B_lo = smem + B_BYTES;
    unsigned char* B_16 = smem + 2 * B_BYTES;                        // dpre_hi again, as bf16 (K-major, 8 per core-matrix row)
    float* WhS = reinterpret_cast<float*>(B_16 + C_BYTES);           // WhS[c*64 + j]
    float* DHS = WhS + NH * H;                                       // DHS[cell warp][4 sequences * NH]: dhead rows of the step being prepared
    float* DHX = DHS + 16 * CT * NH;                                 // DHX[pair][4][lane]: dh_rec of the cells the B thread finishes
    float* OBS = DHX + 8 * CT * 32;                                  // OBS[pair][8 sequences * 4]: observations of the step whose dx is drained next
    float* WTS = OBS + 8 * 32;                                       // Wt[j][0..3]
    float* BTS = WTS + H * 4;                                        // bt[j]
    float* REDS = BTS + H;                                           // end-of-kernel reduction scratch
    uint64_t* bars = reinterpret_cast<uint64_t*>(REDS + RED);        // (8-byte aligned: an even number of floats precedes for every NH)
    uint64_t* b_full = bars, *acc_done = bars + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cellw = warp < NC / 32;
    const int q = warp & 3, cg = (warp >> 2) & 3;                   // TMEM lane quarter, group of 8 sequences

    for (int i = tid; i < NH * H; i += NT) WhS[i] = W_head[i];
    const bool ftorso = fuse && fz.torso != 0;
    if (ftorso) {
        for (int i = tid; i < H * 4; i += NT) WTS[i] = fz.Wt[i];
        for (int i = tid; i < H; i += NT) BTS[i] = fz.bt[i];
    }
    if (tid == 0) {
        tc::mbar_init(b_full, NC / 32); tc::mbar_init(acc_done, NI);
        tc::fence_mbar_init();
    }
    if (warp == NC / 32) tc::tmem_alloc(tmem_slot, TM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
#ifdef B200RL_EMUL
    const int my_lane = 32 * q + lane;
#else
    const int my_lane = 32 * q;
#endif
    // one-time: A[m][p] = W_hh[n(p)][m] (m < 64) | W_ih[n(p)][m-64], n(p) = (p&3)*64 + (p>>2); thread = row m (its TMEM
    // lane), warp (q, cg) fills columns p in [64 cg, 64 cg + 64)
    if (cellw) {
        const int m = 32 * q + lane;
        const float* wsrc = (m < H ? W_hh : W_ih) + (m & (H - 1));
        for (int cgw = cg; cgw < 4; cgw += NG)                   // (16-sequence tiles have two sequence groups: each fills two column blocks)
        for (int k0 = 0; k0 < 64; k0 += 32) {
            float lo16[16];                                          // 32 bf16 of A_lo = 16 packed columns
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int p = cgw * 64 + k0 + hh * 16 + e, n = (p & 3) * H + (p >> 2);
                    tc::split_tf32(wsrc[(size_t)n * H], hi[e], lo[e]);
                }
                tc::tmem_st16(tc::tmem_addr(tmem, my_lane, TM_AH + cgw * 64 + k0 + hh * 16), hi);
#pragma unroll
                for (int e = 0; e < 16; e += 2)
                    lo16[hh * 8 + e / 2] = __uint_as_float((unsigned)tc::bf16_bits(lo[e]) | ((unsigned)tc::bf16_bits(lo[e + 1]) << 16));
            }
            tc::tmem_st16(tc::tmem_addr(tmem, my_lane, TM_AL + (cgw * 64 + k0) / 2), lo16);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, TB, false, false), idesc16 = tc::make_idesc_bf16(128, TB);
    const int ntiles = (B + TB - 1) / TB;

    if (!cellw) {
#if !defined(B200RL_EMUL)
        if constexpr (NI == 4 && TB == 32) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");       // the four issuer warps are one warp-group: their registers go to the cell warps
#endif
        if (lane == 0) {
            const int ii = warp - NC / 32;                           // K-steps [ii 32/NI, (ii+1) 32/NI) into accumulator ii
            const uint32_t D = tmem + TM_D + ii * TB;
            uint32_t n = 0;
            BWP_INIT
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int t = 0; t < T; ++t, ++n) {
                    BWP(16, NC);
                    tc::mbar_wait(b_full, n & 1);                    // dpre_t operand written, accumulators drained
                    tc::fence_after_sync();
                    BWP(17, NC);
#pragma unroll
                    for (int s = ii * (G / 8 / NI); s < (ii + 1) * (G / 8 / NI); ++s) {     // a contiguous K range per issuer
                        const int p0 = s * 8;
                        const tc::SmemDesc bh = tc::make_desc(B_hi + s * 2 * B_LBO, B_LBO, B_SBO);
                        const tc::SmemDesc bl = tc::make_desc(B_lo + s * 2 * B_LBO, B_LBO, B_SBO);
                        tc::mma_tf32_ts(D, tmem + TM_AH + p0, bl, idesc, s > ii * (G / 8 / NI) ? 1u : 0u);   // (order: one operand changes per MMA)
                        tc::mma_tf32_ts(D, tmem + TM_AH + p0, bh, idesc, 1u);
                        if (s & 1)                                   // third split term A_lo . dpre_hi in bf16, K = 16: one MMA per two K-steps
                            tc::mma_bf16_ts(D, tmem + TM_AL + (p0 - 8) / 2, tc::make_desc(B_16 + (s >> 1) * 2 * C_LBO, C_LBO, C_SBO), idesc16, 1u);
                    }
                    tc::mma_commit(acc_done);
                    BWP(18, NC);
                }
            }
        }
    } else {
#if !defined(B200RL_EMUL)
        if constexpr (NI == 4 && TB == 32) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(B200RL_BWD_CELL_REGS));   // (16-sequence tiles: 384 threads, no register shortage)
#endif
        const bool isB = q >= 2;                                     // B threads see the dx lanes, A threads the dh_rec lanes
        const int half = q & 1;
        const int j = half * 32 + lane;                              // LSTM unit (cell math) = dx column (B, TMEM side)
        const int pair = half * 4 + cg;                              // the A and the B warp of (unit half, sequence group)
        const int mine = isB ? CT : 0;                               // my 4 of the group's 8 sequences
        float* DX = DHX + pair * (CT * 32) + lane;                   // + i*32
        float* dhs = DHS + warp * (CT * NH);                         // my warp's dhead rows [i*NH + c]
        float* obs_s = OBS + pair * 32;
        auto read_acc = [&](float (&d)[8]) {                         // my lane's 8 sequences: the NI partial accumulators summed
            uint32_t u[NI][8];                                       // all loads in flight, one wait
#pragma unroll
            for (int a = 0; a < NI; ++a) tc::tmem_ld8_async(tc::tmem_addr(tmem, my_lane, TM_D + a * TB + cg * 8), u[a]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                d[i] = __uint_as_float(u[0][i]);
#pragma unroll
                for (int a = 1; a < NI; ++a) d[i] += __uint_as_float(u[a][i]);
            }
        };
        uint32_t nb = 0;                                             // accumulator batches waited for so far
        // fused small gradients, summed over every step of every tile of this CTA: dW_head[o][j], db_head[lane] (lanes < NH of the
        // half-0 warps), and on B threads dWt[j][f], dbt[j]
        float aWh[NHT > 0 ? NHT : 1], aWt[4], abt = 0.f, abh = 0.f;
#pragma unroll
        for (int o = 0; o < (NHT > 0 ? NHT : 1); ++o) aWh[o] = 0.f;
#pragma unroll
        for (int f = 0; f < 4; ++f) aWt[f] = 0.f;
        BWP_INIT
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b0 = tile * TB + cg * 8;                       // first sequence of the group ; mine: b0 + mine + i
            auto seq = [&](int i) { const int b = b0 + mine + i; return b < B ? b : B - 1; };
            auto rowof = [&](int i) { return (unsigned)seq(i) * (unsigned)T; };   // B*T < 2^31 (checked by the launcher)
            auto ok = [&](int i) { return b0 + mine + i < B; };
            float4 gcur[CT]; float ccur[CT], cprev[CT];
            float cf_w[CT], cf_po[CT], cf_q[CT], cf_pi[CT], cf_pf[CT], cf_pg[CT], cf_f[CT];    // the step's coefficients, in registers
            auto load_step = [&](int t) {                            // saved activations of step t and c_{t-1}, my 4 cells
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const size_t o = (size_t)(rowof(i) + t) * H + j;
                    gcur[i] = ld4(gates + o * 4);
                    cprev[i] = t > 0 ? c_sav[o - H] : cx0[(size_t)seq(i) * H + j];
                }
            };
            auto fetch_dhead = [&](int t) {                          // asynchronous: lands in shared memory, no registers held
                __syncwarp();                                        // the previous step's values have been consumed by every lane
                for (int e = lane; e < CT * NH; e += 32) {
                    const int i = e / NH, c = e - i * NH;
                    const int bb = b0 + mine + i;
                    cp_async4(dhs + e, dhead + ((size_t)(bb < B ? bb : B - 1) * T + t) * NH + c);
                }
            };
            auto fetch_obs = [&](int t) {                            // (B, fused torso) observations of step t for the group's 8 sequences
                __syncwarp();
                const int bb = b0 + (lane >> 2);
                cp_async4(obs_s + lane, fz.u + ((size_t)(bb < B ? bb : B - 1) * T + t) * 4 + (lane & 3));
            };
            auto torso = [&](const float (&dxv)[8]) {                // dWt, dbt from dx of the step whose observations sit in obs_s
                const float4 w = ld4(WTS + j * 4);
                const float btj = BTS[j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 o4 = ld4(obs_s + i * 4);
                    const float x = fmaf(w.x, o4.x, fmaf(w.y, o4.y, fmaf(w.z, o4.z, fmaf(w.w, o4.w, btj))));
                    const float d = (x > 0.f && b0 + i < B) ? dxv[i] : 0.f;
                    aWt[0] = fmaf(d, o4.x, aWt[0]); aWt[1] = fmaf(d, o4.y, aWt[1]);
                    aWt[2] = fmaf(d, o4.z, aWt[2]); aWt[3] = fmaf(d, o4.w, aWt[3]);
                    abt += d;
                }
            };
            auto coefs = [&]() {                                     // everything of the loaded step that does not need dh
                cp_async_wait_all();
                __syncwarp();
                float tch[CT];
#pragma unroll
                for (int i = 0; i < CT; ++i) { cf_w[i] = 0.f; tch[i] = tanh_(ccur[i]); }
                if (fuse) {
                    // compile-time head count: fully unrolled ; dW_head[c][j] += dhead[b,t,c] h[b,t,j] rides on the same dhead values
#pragma unroll
                    for (int c = 0; c < (NHT > 0 ? NHT : 1); ++c) {
                        const float w = WhS[c * H + j];
                        float a = 0.f;
#pragma unroll
                        for (int i = 0; i < CT; ++i) {
                            const float d = dhs[i * NHT + c];
                            cf_w[i] = fmaf(d, w, cf_w[i]);
                            a = fmaf(ok(i) ? d : 0.f, gcur[i].w * tch[i], a);
                        }
                        aWh[c] += a;
                    }
                    if (half == 0 && lane < NHT) {
#pragma unroll
                        for (int i = 0; i < CT; ++i) abh += ok(i) ? dhs[i * NHT + lane] : 0.f;
                    }
                } else {
                    // head index outermost: ONE runtime loop with independent FMA chains inside
                    for (int c = 0; c < NH; ++c) {
                        const float w = WhS[c * H + j];
#pragma unroll
                        for (int i = 0; i < CT; ++i) cf_w[i] = fmaf(dhs[i * NH + c], w, cf_w[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const float4 g4 = gcur[i];                       // i, f, g, o (post-activation)
                    const float th = tch[i];
                    cf_po[i] = th * g4.w * (1.f - g4.w);
                    cf_q[i] = g4.w * (1.f - th * th);
                    cf_pi[i] = g4.z * g4.x * (1.f - g4.x);
                    cf_pf[i] = cprev[i] * g4.y * (1.f - g4.y);
                    cf_pg[i] = g4.x * (1.f - g4.z * g4.z);
                    cf_f[i] = g4.y;
                }
            };
#pragma unroll
            for (int i = 0; i < CT; ++i) ccur[i] = c_sav[(size_t)(rowof(i) + T - 1) * H + j];
            load_step(T - 1);
            fetch_dhead(T - 1);
            coefs();
            float dc[CT];
#pragma unroll
            for (int i = 0; i < CT; ++i) dc[i] = 0.f;
            for (int t = T - 1; t >= 0; --t) {
                BWP(0, 0); BWP(8, 64);
                // next step's inputs: c_{t-1} becomes c_t ; the loads fly during this step's hand-over and are consumed by coefs() below
#pragma unroll
                for (int i = 0; i < CT; ++i) ccur[i] = cprev[i];
                if (t > 0) { load_step(t - 1); fetch_dhead(t - 1); }
                float acc[8];                                        // A: dh_rec of the group's 8 sequences ; B: dx_{t+1} of them
                if (t < T - 1) {
                    tc::mbar_wait(acc_done, nb & 1); ++nb;
                    tc::fence_after_sync();
                    BWP(1, 0); BWP(9, 64);
                    read_acc(acc);
                    tc::fence_before_sync();
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                }
                BWP(2, 0); BWP(10, 64);
                float dh4[CT];
                if (!isB) {
#pragma unroll
                    for (int i = 0; i < CT; ++i) { DX[i * 32] = acc[CT + i]; dh4[i] = acc[i]; }
                    tc::named_bar_sync(1 + pair, 64);
                } else {
                    tc::named_bar_sync(1 + pair, 64);                // the A thread has published dh_rec of my cells
#pragma unroll
                    for (int i = 0; i < CT; ++i) dh4[i] = DX[i * 32];
                }
                BWP(3, 0); BWP(11, 64);
                // the dh-dependent tail: 7 flops per cell from the register coefficients ; dpre to the B operand (hi / lo / bf16)
                float4 dps[CT];
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const float dh = dh4[i] + cf_w[i];
                    const float dct = fmaf(dh, cf_q[i], dc[i]);
                    const float4 dp = make_float4(dct * cf_pi[i], dct * cf_pf[i], dct * cf_pg[i], dh * cf_po[i]);
                    dc[i] = dct * cf_f[i];
                    dps[i] = dp;
                    float4 hi, lo;
                    tc::split_tf32(dp.x, hi.x, lo.x); tc::split_tf32(dp.y, hi.y, lo.y);
                    tc::split_tf32(dp.z, hi.z, lo.z); tc::split_tf32(dp.w, hi.w, lo.w);
                    const int r = cg * 8 + mine + i;
                    const int ob = (r & 7) * 16 + (r >> 3) * B_SBO + j * B_LBO;
                    *reinterpret_cast<float4*>(B_hi + ob) = hi;
                    *reinterpret_cast<float4*>(B_lo + ob) = lo;
                    const unsigned long long h16 = (unsigned long long)((unsigned)tc::bf16_bits(hi.x) | ((unsigned)tc::bf16_bits(hi.y) << 16)) |
                                                   ((unsigned long long)((unsigned)tc::bf16_bits(hi.z) | ((unsigned)tc::bf16_bits(hi.w) << 16)) << 32);
                    *reinterpret_cast<unsigned long long*>(B_16 + (r & 7) * 16 + (r >> 3) * C_SBO + (j >> 1) * C_LBO + (j & 1) * 8) = h16;
                }
                BWP(4, 0); BWP(12, 64);
                tc::fence_async_smem();
                tc::mbar_arrive_warp(b_full);
                BWP(5, 0); BWP(13, 64);
                // ---- off the serial chain: the tensor core is running this step's GEMM
#pragma unroll
                for (int i = 0; i < CT; ++i)
                    if (ok(i)) st4(gates + ((size_t)(rowof(i) + t) * H + j) * 4, dps[i]);
                if (isB && t < T - 1) {
                    if (ftorso) {
                        torso(acc);                                  // (obs_{t+1} landed before the previous coefs())
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int bb = b0 + i;
                            if (bb < B) dx[((size_t)bb * T + t + 1) * H + j] = acc[i];
                        }
                    }
                }
                if (isB && ftorso) fetch_obs(t);                     // for dx_t, drained in the next iteration
                if (t > 0) coefs();                                  // step t-1, from the loads issued at the top of this iteration
                BWP(6, 0); BWP(14, 64);
            }
            tc::mbar_wait(acc_done, nb & 1); ++nb;                   // batch of step 0 (dx_0)
            tc::fence_after_sync();
            if (isB) {
                float d[8];
                read_acc(d);
                if (ftorso) {
                    cp_async_wait_all();
                    __syncwarp();
                    torso(d);
                    __syncwarp();                                    // (the next tile's first fetch_obs overwrites the slot)
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int bb = b0 + i;
                        if (bb < B) dx[(size_t)bb * T * H + j] = d[i];
                    }
                }
            }
            tc::fence_before_sync();
        }
        if (fuse) {
            // hand the register accumulators to the CTA-wide reduction below: A and B threads of a (group, unit) own different
            // sequences, so both contribute
            float* red = REDS + cg * (10 * H) + j;
            if (!isB) {
#pragma unroll
                for (int o = 0; o < 4; ++o) red[o * H] = o < NHT ? aWh[o < NHT ? o : 0] : 0.f;
                red[9 * H] = abh;
            }
            tc::named_bar_sync(1 + pair, 64);
            if (isB) {
#pragma unroll
                for (int o = 0; o < 4; ++o) red[o * H] += o < NHT ? aWh[o < NHT ? o : 0] : 0.f;
#pragma unroll
                for (int f = 0; f < 4; ++f) red[(4 + f) * H] = aWt[f];
                red[8 * H] = abt;
                red[9 * H] += abh;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (fuse) {
        float* slab = fz.partials + (size_t)blockIdx.x * fz.P;
        for (int e = tid; e < 10 * H; e += NT) {
            const int k = e / H, jj = e - k * H;
            float v = REDS[e] + REDS[10 * H + e];
            if (NG == 4) v += REDS[2 * 10 * H + e] + REDS[3 * 10 * H + e];
            if (k < 4) { if (k < NHT) slab[fz.off_Whead + (size_t)k * H + jj] = v; }
            else if (k < 8) { if (ftorso) slab[fz.off_Wt + (size_t)jj * 4 + (k - 4)] = v; }
            else if (k == 8) { if (ftorso) slab[fz.off_bt + jj] = v; }
            else if (jj < NHT) slab[fz.off_bhead + jj] = v;
        }
    }
    if (warp == NC / 32) tc::tmem_dealloc(tmem, TM_COLS);
}

int launch_bwd_tc(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih, float* gates,
                  const float* c, const float* cx0, float* dx, int B, int T, int NH, int sms, void* stream) {
    return launch_bwd_tc_fused(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0,
                               B, T, 0, NH, sms, stream);
}

template <int NHT, int NI, int TB>
static int launch_bwdt_ni(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih, float* gates,
                       const float* c, const float* cx0, float* dx, int B, int T, int NH, BwdFuse fz, int sms, void* stream) {
    const int ntiles = (B + TB - 1) / TB;
    const size_t smem = bwdt_smem_bytes(NH, TB);
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(lstm_bwdt_kernel<NHT, NI, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lstm_bwd(tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
    B200RL_LAUNCH((lstm_bwdt_kernel<NHT, NI, TB>), ntiles < sms ? ntiles : sms, btk::nthreads(NI, TB), smem, stream, dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz);
    return check_launch("lstm_bwd(tc)");
}

template <int NHT>
static int launch_bwdt(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih, float* gates,
                       const float* c, const float* cx0, float* dx, int B, int T, int NH, BwdFuse fz, int sms, void* stream) {
    constexpr int NIT = NHT > 0 ? B200RL_BWD_NI_TORSO : B200RL_BWD_NI;
    // (T >= 2: a CTA writes one slab row of the fused gradients, and there are min(2 SMs, B T / 32) rows)
    if (recurrent_tile16(B, sms) && T >= 2) {
        if (NHT > 0 && fz.torso) return launch_bwdt_ni<NHT, NIT, 16>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
        return launch_bwdt_ni<NHT, B200RL_BWD_NI, 16>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
    }
    if (NHT > 0 && fz.torso) return launch_bwdt_ni<NHT, NIT, 32>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
    return launch_bwdt_ni<NHT, B200RL_BWD_NI, 32>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
}

// partials != nullptr: head gradients fused (NH <= 4); with dx == nullptr also the torso gradients (Fin == 4) and dx is never stored
int launch_bwd_tc_fused(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih, float* gates,
                        const float* c, const float* cx0, float* dx, const float* u, const float* Wt, const float* bt,
                        float* partials, long long P, long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                        int B, int T, int Fin, int NH, int sms, void* stream) {
    B200RL_REQUIRE((long long)B * T < (1LL << 31), "lstm_bwd(tc): B*T must be below 2^31");
    B200RL_REQUIRE(W_ih != nullptr, "lstm_bwd(tc): needs W_ih");
    BwdFuse fz{u, Wt, bt, partials, P, off_Wt, off_bt, off_Whead, off_bhead, dx == nullptr ? 1 : 0};
    if (partials == nullptr) {
        B200RL_REQUIRE(dx != nullptr, "lstm_bwd(tc): needs the dx workspace");
        return launch_bwdt<0>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
    }
    B200RL_REQUIRE(NH >= 1 && NH <= 4, "lstm_bwd(tc, fused gradients): needs NH <= 4 (NH=%d)", NH);
    B200RL_REQUIRE(dx != nullptr || (Fin == 4 && u && Wt && bt), "lstm_bwd(tc, fused torso gradients): needs Fin == 4 (Fin=%d)", Fin);
    switch (NH) {
        case 1: return launch_bwdt<1>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
        case 2: return launch_bwdt<2>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
        case 3: return launch_bwdt<3>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
        default: return launch_bwdt<4>(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, fz, sms, stream);
    }
}

}  // namespace b200rl
