// Recurrent half of the LSTM forward on tcgen05, PING-PONG version (sm_100a).
//
// Same math and data layout as fwd_tc.cu (read that header first).  There the per-step GEMM (tensor core) and the cell
// phase (MUFU / FP32 / LSU) of a CTA's 32 sequences alternate, each waiting for the other (profiles/r1n: 0.9K + 0.5K
// cycles of MMAs and hand-over, 2.3K of cell phase, per step).  Here the 32 sequences are two groups of 16 with their own
// accumulators, h operand and issuer threads; every cell warp owns cells in both groups and works on group B while the
// tensor core runs group A's next step, and vice versa.  Small MMAs are issue-bound per thread (tools/tc_probe3.cu), so
// doubling the MMA count (N = 16 instead of 32) costs nothing as long as each group has its own issuer threads: 8 issuer
// warps = (group) x (M tile) x (even / odd K-steps), 12 MMAs each per step.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

namespace ft2 {
constexpr int H = 64, G = 256, TB = 32, GR = 16;           // sequences per CTA, per group
constexpr int NGW = 4;                                     // row sub-groups of a group: 4 cell warps (one per TMEM lane quarter) each
constexpr int RW = GR / NGW, CT = RW / 2;                  // sequences a warp reads per group (4) / cells a thread owns per group (2)
constexpr int NI = 8;                                      // MMA issuer warps: (group) x (M tile) x (even / odd K-steps)
constexpr int NC = NGW * 128, NT = NC + 32 * NI;
constexpr int TM_D = 0, TM_A = NI * GR, TM_COLS = 512;     // accumulator (g, mt, kh) at 16 ((2 g + mt) 2 + kh) ; A at TM_A + (mt*2 + lo)*64 + k
constexpr int B_LBO = 144, B_SBO = 2304;                   // B: 16 core matrices along K, 16-byte skew -> conflict-free 4-byte scatter
constexpr int B_BYTES = (GR / 8) * B_SBO;                  // 4608 per group, per hi / lo
constexpr int EXG = 2 * NGW * 2 * RW * 32;                 // swap buffer of one group: [pair][direction][slot][lane] floats
static_assert(RW == 4, "the cell phase reads 4 accumulator columns per tile (tmem_ld4)");
static_assert(15 * B_LBO + 128 <= B_SBO, "core matrices of one 8-row group must fit below SBO");
static_assert(TM_A + 4 * H <= TM_COLS, "tensor memory budget");
}  // namespace ft2

static inline size_t fwdt2_smem_bytes() {
    const size_t need = 2 * 2 * (size_t)ft2::B_BYTES + sizeof(float) * 2 * ft2::EXG + 64;
    return need < 120 * 1024 ? 120 * 1024 : need;          // (owns all 512 TMEM columns: keep a second CTA off the SM)
}

__global__ void __launch_bounds__(ft2::NT, 1)
lstm_fwdt2_kernel(float* xg, const float* __restrict__ hx0, const float* __restrict__ cx0,
                  const float* __restrict__ W_hh, float* __restrict__ c_out, float* __restrict__ h_out, int B, int T) {
    using namespace ft2;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* Bop = smem;                              // group g: hi at g*2*B_BYTES, lo right after
    float* EX = reinterpret_cast<float*>(smem + 2 * 2 * B_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(EX + 2 * EXG);
    uint64_t* b_full = bars, *acc_done = bars + 2;          // [group]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool compute = warp < NC / 32;

    if (tid == 0) {
        for (int g = 0; g < 2; ++g) { tc::mbar_init(&b_full[g], NC / 32); tc::mbar_init(&acc_done[g], NI / 2); }
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
    const uint32_t idesc = tc::make_idesc_tf32(128, GR, false, false);
    const int ntiles = (B + TB - 1) / TB;

    if (!compute) {
        if (lane == 0) {
            const int ii = warp - NC / 32, g = ii >> 2, mt = (ii >> 1) & 1, kh = ii & 1;
            const uint32_t Ah = tmem + TM_A + (mt * 2 + 0) * H, Al = tmem + TM_A + (mt * 2 + 1) * H, D = tmem + TM_D + ii * GR;
            unsigned char* Bh = Bop + g * 2 * B_BYTES, *Bl = Bh + B_BYTES;
            uint32_t n = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int t = 0; t < T; ++t, ++n) {
                    tc::mbar_wait(&b_full[g], n & 1);             // this group's h_{t-1} operand written, its accumulators drained
                    tc::fence_after_sync();
#pragma unroll
                    for (int s = kh; s < H / 8; s += 2) {
                        const tc::SmemDesc bh = tc::make_desc(Bh + s * 2 * B_LBO, B_LBO, B_SBO);
                        const tc::SmemDesc bl = tc::make_desc(Bl + s * 2 * B_LBO, B_LBO, B_SBO);
                        tc::mma_tf32_ts(D, Ah + s * 8, bh, idesc, s > 1 ? 1u : 0u);
                        tc::mma_tf32_ts(D, Ah + s * 8, bl, idesc, 1u);
                        tc::mma_tf32_ts(D, Al + s * 8, bh, idesc, 1u);
                    }
                    tc::mma_commit(&acc_done[g]);
                }
            }
        }
    } else {
        const int q = warp & 3, cq = warp >> 2;                // TMEM lane quarter, row sub-group (4 sequences of each group)
        const bool fo = q >= 2;                                // false: gates (i, g) ; true: gates (f, o)
        const int j = (q & 1) * 32 + lane;                     // unit
        const int pair = (q & 1) + 2 * cq;
        const int mine = fo ? CT : 0;                          // of the warp's RW sequences, this thread finishes [mine, mine + CT)
#ifdef B200RL_EMUL
        const int addr_lane = 32 * q + lane;
#else
        const int addr_lane = 32 * q;
#endif
        auto boff = [&](int r) { return (r & 7) * 16 + (r >> 3) * B_SBO + (j & 3) * 4 + (j >> 2) * B_LBO; };   // r = row inside the group
        uint32_t n = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int b0 = tile * TB;
            // my cells: group g, sequences b0 + 16 g + 4 cq + mine + i (i < CT), unit j
            auto seqi = [&](int g, int i) { const int b = b0 + g * GR + cq * RW + mine + i; return b < B ? b : B - 1; };
            auto okay = [&](int g, int i) { return b0 + g * GR + cq * RW + mine + i < B; };
            auto rowof = [&](int g, int i) { return (unsigned)seqi(g, i) * (unsigned)T; };      // B*T < 2^31 (checked by the launcher)
            float cst[2][CT];
            float4 xq[2][CT];                                  // xproj of my cells, one step ahead
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                unsigned char* Bh = Bop + g * 2 * B_BYTES, *Bl = Bh + B_BYTES;
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    const size_t b = (size_t)seqi(g, i);
                    cst[g][i] = cx0[b * H + j];
                    float hi, lo;
                    tc::split_tf32(hx0[b * H + j], hi, lo);
                    const int ob = boff(cq * RW + mine + i);
                    *reinterpret_cast<float*>(Bh + ob) = hi;
                    *reinterpret_cast<float*>(Bl + ob) = lo;
                    xq[g][i] = ld4(xg + ((size_t)rowof(g, i) * H + j) * 4);
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive_warp(&b_full[0]);
            tc::mbar_arrive_warp(&b_full[1]);

            for (int t = 0; t < T; ++t, ++n) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    unsigned char* Bh = Bop + g * 2 * B_BYTES, *Bl = Bh + B_BYTES;
                    float* EXs = EX + g * EXG + pair * (2 * RW * 32) + (fo ? RW * 32 : 0);          // my outgoing slots [slot][lane]
                    const float* EXr = EX + g * EXG + pair * (2 * RW * 32) + (fo ? 0 : RW * 32);    // the partner's
                    tc::mbar_wait(&acc_done[g], n & 1);
                    tc::fence_after_sync();
                    float xa[RW], xb[RW];
                    {                                           // gate i (or f) / g (or o) of unit j, RW sequences: even + odd K-step partials
                        float e[RW];
                        const int c0 = TM_D + g * 4 * GR + cq * RW;
                        tc::tmem_ld4(tc::tmem_addr(tmem, addr_lane, c0 + 0 * GR), xa);
                        tc::tmem_ld4(tc::tmem_addr(tmem, addr_lane, c0 + 1 * GR), e);
#pragma unroll
                        for (int r = 0; r < RW; ++r) xa[r] += e[r];
                        tc::tmem_ld4(tc::tmem_addr(tmem, addr_lane, c0 + 2 * GR), xb);
                        tc::tmem_ld4(tc::tmem_addr(tmem, addr_lane, c0 + 3 * GR), e);
#pragma unroll
                        for (int r = 0; r < RW; ++r) xb[r] += e[r];
                    }
#pragma unroll
                    for (int i = 0; i < CT; ++i) { EXs[(2 * i) * 32 + lane] = fo ? xa[i] : xa[CT + i]; EXs[(2 * i + 1) * 32 + lane] = fo ? xb[i] : xb[CT + i]; }
                    tc::named_bar_sync(1 + pair, 64);
                    float4 gsv[CT]; float hsv[CT];             // saved activations: stored to HBM AFTER the step is handed to the issuers
#pragma unroll
                    for (int i = 0; i < CT; ++i) {
                        const float pa = EXr[(2 * i) * 32 + lane], pb = EXr[(2 * i + 1) * 32 + lane];
                        const float4 x = xq[g][i];
                        if (t + 1 < T) xq[g][i] = ld4(xg + ((size_t)(rowof(g, i) + t + 1) * H + j) * 4);
                        float ig, fg, gg, og;
                        sigm_tanh((fo ? pa : xa[i]) + x.x, (fo ? pb : xb[i]) + x.z, ig, gg);
                        sigm_sigm((fo ? xa[CT + i] : pa) + x.y, (fo ? xb[CT + i] : pb) + x.w, fg, og);
                        const float cn = fmaf(fg, cst[g][i], ig * gg);
                        const float hh = og * tanh_(cn);
                        cst[g][i] = cn;
                        gsv[i] = make_float4(ig, fg, gg, og); hsv[i] = hh;
                        float hi, lo;
                        tc::split_tf32(hh, hi, lo);
                        const int ob = boff(cq * RW + mine + i);
                        *reinterpret_cast<float*>(Bh + ob) = hi;
                        *reinterpret_cast<float*>(Bl + ob) = lo;
                    }
                    if (t + 1 < T) {
                        tc::fence_async_smem();
                        tc::fence_before_sync();
                        tc::mbar_arrive_warp(&b_full[g]);
                    }
#pragma unroll
                    for (int i = 0; i < CT; ++i) {              // (these stores overlap the tensor core's next step)
                        if (okay(g, i)) {
                            const size_t o = (size_t)(rowof(g, i) + t) * H + j;
                            st4(xg + o * 4, gsv[i]);
                            c_out[o] = cst[g][i]; h_out[o] = hsv[i];
                        }
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NC / 32) tc::tmem_dealloc(tmem, TM_COLS);
}

int launch_fwd_tc2(float* xg, const float* hx0, const float* cx0, const float* W_hh, float* c, float* h,
                   int B, int T, int sms, void* stream) {
    B200RL_REQUIRE((long long)B * T < (1LL << 31), "lstm_fwd(tc2): B*T must be below 2^31");
    const int ntiles = (B + ft2::TB - 1) / ft2::TB;
    const size_t smem = fwdt2_smem_bytes();
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(lstm_fwdt2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lstm_fwd(tc2): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
    B200RL_LAUNCH(lstm_fwdt2_kernel, ntiles < sms ? ntiles : sms, ft2::NT, smem, stream, xg, hx0, cx0, W_hh, c, h, B, T);
    return check_launch("lstm_fwd(tc2)");
}

}  // namespace b200rl
