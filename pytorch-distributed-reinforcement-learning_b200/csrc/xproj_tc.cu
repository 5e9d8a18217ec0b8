// LSTM input projection for ALL timesteps at once on tcgen05 (sm_100a):
//
//   xproj[b,t,p] = b_ih[n(p)] + b_hh[n(p)] + sum_k ReLU(u[b,t,:] . Wt[k,:] + bt[k]) * W_ih[n(p)][k]
//
// This half of nn.LSTMCell's gate pre-activation (networks/models.py:67-75) does not depend on the
// recurrence, so it leaves the serial kernel: one [256 x 64] . [64 x B*T] GEMM with full parallelism.
// Output columns keep the gate-interleaved order p = j*4 + g of the saved activations (n(p) = g*64 + j is
// the PyTorch row); the recurrent kernel adds its accumulators to them.
//
// Weights on M: W_ih, rows permuted into the p order, is split once into TF32 hi/lo and lives in TENSOR
// MEMORY as the A operand (2 M tiles x (hi, lo) x 64 k = 256 columns); a tile of 64 (b,t) rows is the
// K-major B operand (N = 64): a producer thread recomputes x for its row (4-wide obs), splits hi/lo and
// writes 16-byte core-matrix rows.  3xTF32 into one of two accumulator sets (2 x 2 x 64 columns), so the
// epilogue of tile i overlaps the MMAs of tile i+1.  With the gate index on the TMEM lanes the epilogue
// needs no transposition: a warp's 32 lanes hold 32 consecutive p of one (b,t) row = one 128-byte line
// per store.  (The first version had the rows on M and W_ih as a 128 KB shared-memory B operand; its
// epilogue threads each owned a ROW and went through a shared-memory patch to get 64-byte segments:
// 0.156 ms for the 537 MB write, profiles/r1m_ncu_summary.md.)
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

namespace xpk {
constexpr int H = 64, G = 256, NP = 512, NI = 2, NT = NP + 32 * NI;   // 16 producer / epilogue warps + one MMA issuer warp per M tile
constexpr int TR = 64;                                       // (b,t) rows per tile = MMA N
constexpr int B_LBO = 128, B_SBO = (H / 4) * 128;            // B: 16 core matrices along K per 8-row group (2048)
constexpr int B_BYTES = (TR / 8) * B_SBO;                    // 16384 per hi / lo
constexpr int NSTAGE = 2;
constexpr int TM_D = 0, TM_A = 2 * 2 * TR, TM_COLS = 512;    // D: (buffer, M tile) at 64 (2 buf + mt) ; A at TM_A + (mt*2 + lo)*64 + k
constexpr int MAXF = 32;
static_assert(TM_A + 4 * H <= TM_COLS, "tensor memory budget");
}  // namespace xpk

static inline size_t xproj_smem_bytes(int Fin) {
    const size_t need = xpk::NSTAGE * 2 * (size_t)xpk::B_BYTES + sizeof(float) * (size_t)(Fin * xpk::H + xpk::H + xpk::G) + 128;
    return need < 120 * 1024 ? 120 * 1024 : need;      // the CTA owns all 512 TMEM columns: ask for more than half the SM's shared memory so that no second CTA is co-scheduled
}

template <int FMAX>
__global__ void __launch_bounds__(xpk::NT, 1)
lstm_xproj_tc_kernel(const float* __restrict__ u, const float* __restrict__ Wt, const float* __restrict__ bt,
                     const float* __restrict__ W_ih, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                     float* __restrict__ xproj, long long rows, int Fin) {
    using namespace xpk;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* Bst = smem;                                           // stage s: hi at s*2*B_BYTES, lo right after
    float* WtS = reinterpret_cast<float*>(smem + NSTAGE * 2 * B_BYTES);  // WtS[q*64 + k]
    float* bts = WtS + Fin * H;
    float* bsum = bts + H;                                               // bsum[p]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsum + G);
    uint64_t* b_full = bars, *b_empty = bars + 2, *acc_done = bars + 4, *acc_free = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < NP / 32;

    for (int i = tid; i < Fin * H; i += NT) { int q = i / H, k = i % H; WtS[i] = Wt[k * Fin + q]; }
    if (tid < H) bts[tid] = bt[tid];
    if (tid < G) { const int p = tid, n = (p & 3) * 64 + (p >> 2); bsum[p] = b_ih[n] + b_hh[n]; }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&b_full[i], NP / 32); tc::mbar_init(&b_empty[i], NI);
            tc::mbar_init(&acc_done[i], NI); tc::mbar_init(&acc_free[i], NP / 32);
        }
        tc::fence_mbar_init();
    }
    if (warp == NP / 32) tc::tmem_alloc(tmem_slot, TM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int q = warp & 3;                                              // TMEM lane quarter of this warp
#ifdef B200RL_EMUL
    const int my_lane = 32 * q + lane;
#else
    const int my_lane = 32 * q;
#endif
    // one-time: W_ih rows in the p order -> TF32 hi/lo in tensor memory; warps 0-3 fill M tile 0, warps 4-7 tile 1
    if (warp < 8) {
        const int mt = warp >> 2, p = mt * 128 + 32 * q + lane;
        const float* wrow = W_ih + (size_t)((p & 3) * 64 + (p >> 2)) * H;
        for (int k0 = 0; k0 < H; k0 += 16) {
            float hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
                const float4 w = ld4(wrow + k0 + e);
                tc::split_tf32(w.x, hi[e], lo[e]); tc::split_tf32(w.y, hi[e + 1], lo[e + 1]);
                tc::split_tf32(w.z, hi[e + 2], lo[e + 2]); tc::split_tf32(w.w, hi[e + 3], lo[e + 3]);
            }
            tc::tmem_st16(tc::tmem_addr(tmem, my_lane, TM_A + (mt * 2 + 0) * H + k0), hi);
            tc::tmem_st16(tc::tmem_addr(tmem, my_lane, TM_A + (mt * 2 + 1) * H + k0), lo);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, TR, false, false);
    const long long ntiles = (rows + TR - 1) / TR;

    if (!producer) {
        if (lane == 0) {
            const int mt = warp - NP / 32;
            const uint32_t Ah = tmem + TM_A + (mt * 2 + 0) * H, Al = tmem + TM_A + (mt * 2 + 1) * H;
            int n = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
                const int s = n & 1, use = n >> 1;
                unsigned char* Bh = Bst + s * 2 * B_BYTES, *Bl = Bh + B_BYTES;
                tc::mbar_wait(&b_full[s], (uint32_t)(use & 1));
                if (use >= 1) tc::mbar_wait(&acc_free[s], (uint32_t)((use - 1) & 1));
                tc::fence_after_sync();
                const uint32_t D = tmem + TM_D + (s * 2 + mt) * TR;
#pragma unroll
                for (int ks = 0; ks < H / 8; ++ks) {
                    const tc::SmemDesc bh = tc::make_desc(Bh + ks * 2 * B_LBO, B_LBO, B_SBO);
                    const tc::SmemDesc bl = tc::make_desc(Bl + ks * 2 * B_LBO, B_LBO, B_SBO);
                    tc::mma_tf32_ts(D, Ah + ks * 8, bh, idesc, ks > 0 ? 1u : 0u);
                    tc::mma_tf32_ts(D, Ah + ks * 8, bl, idesc, 1u);
                    tc::mma_tf32_ts(D, Al + ks * 8, bh, idesc, 1u);
                }
                tc::mma_commit(&b_empty[s]);
                tc::mma_commit(&acc_done[s]);
            }
        }
    } else {
        // producer thread: row r of the tile, k octant kq (8 of the 64 torso units)
        const int r = tid & 63, kq = tid >> 6;
        float uq[FMAX];
        auto load_u = [&](long long tile) {
            const long long row = tile * TR + r;
#pragma unroll
            for (int qq = 0; qq < FMAX; ++qq) uq[qq] = (qq < Fin && row < rows) ? __ldg(u + (size_t)row * Fin + qq) : 0.f;
        };
        auto fill = [&](long long tile, int n) {
            const int s = n & 1, use = n >> 1;
            unsigned char* Bh = Bst + s * 2 * B_BYTES, *Bl = Bh + B_BYTES;
            if (use >= 1) tc::mbar_wait(&b_empty[s], (uint32_t)((use - 1) & 1));
            const bool ok = tile * TR + r < rows;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
                const int k0 = kq * 8 + c4 * 4;
                float4 a = *reinterpret_cast<const float4*>(bts + k0);
#pragma unroll
                for (int qq = 0; qq < FMAX; ++qq) {
                    if (qq < Fin) {
                        const float4 w = *reinterpret_cast<const float4*>(WtS + qq * H + k0);
                        a.x = fmaf(uq[qq], w.x, a.x); a.y = fmaf(uq[qq], w.y, a.y); a.z = fmaf(uq[qq], w.z, a.z); a.w = fmaf(uq[qq], w.w, a.w);
                    }
                }
                const float xv[4] = {ok ? fmaxf(a.x, 0.f) : 0.f, ok ? fmaxf(a.y, 0.f) : 0.f, ok ? fmaxf(a.z, 0.f) : 0.f, ok ? fmaxf(a.w, 0.f) : 0.f};
                float4 hi, lo;
                tc::split_tf32(xv[0], hi.x, lo.x); tc::split_tf32(xv[1], hi.y, lo.y);
                tc::split_tf32(xv[2], hi.z, lo.z); tc::split_tf32(xv[3], hi.w, lo.w);
                const int o = (r & 7) * 16 + (r >> 3) * B_SBO + (k0 >> 2) * B_LBO;
                *reinterpret_cast<float4*>(Bh + o) = hi;
                *reinterpret_cast<float4*>(Bl + o) = lo;
            }
            tc::fence_async_smem();
            tc::mbar_arrive_warp(&b_full[s]);
        };
        const int ch = warp >> 2;                                        // which 16 of the tile's 64 rows this warp stores
        auto epilogue = [&](long long tile, int n) {
            const int s = n & 1, use = n >> 1;
            tc::mbar_wait(&acc_done[s], (uint32_t)(use & 1));
            tc::fence_after_sync();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float v[16];
                tc::tmem_ld16(tc::tmem_addr(tmem, my_lane, TM_D + (s * 2 + mt) * TR + ch * 16), v);
                const int p = mt * 128 + 32 * q + lane;
                const float bias = bsum[p];
                float* dst = xproj + (size_t)(tile * TR + ch * 16) * G + p;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (tile * TR + ch * 16 + i < rows) dst[(size_t)i * G] = v[i] + bias;      // 32 lanes = one 128-byte line
            }
            tc::fence_before_sync();
            tc::mbar_arrive_warp(&acc_free[s]);
        };
        int n = 0;
        long long prev = -1;
        if ((long long)blockIdx.x < ntiles) load_u(blockIdx.x);
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
            fill(tile, n);
            if (tile + gridDim.x < ntiles) load_u(tile + gridDim.x);
            if (prev >= 0) epilogue(prev, n - 1);
            prev = tile;
        }
        if (prev >= 0) epilogue(prev, n - 1);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NP / 32) tc::tmem_dealloc(tmem, TM_COLS);
}

int launch_xproj_tc(const float* u, const float* Wt, const float* bt, const float* W_ih, const float* b_ih, const float* b_hh,
                    float* xproj, int B, int T, int Fin, int sms, void* stream) {
    B200RL_REQUIRE(Fin >= 1 && Fin <= xpk::MAXF, "lstm_fwd(xproj): Fin=%d out of range", Fin);
    const long long rows = (long long)B * T;
    const long long ntiles = (rows + xpk::TR - 1) / xpk::TR;
    const size_t smem = xproj_smem_bytes(Fin);
    const int grid = (int)(ntiles < sms ? ntiles : sms);
#ifndef B200RL_EMUL
#define B200RL_XPROJ_ATTR(F)                                                                                             \
    {                                                                                                                    \
        cudaError_t e = cudaFuncSetAttribute(lstm_xproj_tc_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return set_error("lstm_fwd(xproj): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e)); \
    }
#else
#define B200RL_XPROJ_ATTR(F)
#endif
    if (Fin <= 4) {
        B200RL_XPROJ_ATTR(4)
        B200RL_LAUNCH(lstm_xproj_tc_kernel<4>, grid, xpk::NT, smem, stream, u, Wt, bt, W_ih, b_ih, b_hh, xproj, rows, Fin);
    } else if (Fin <= 8) {
        B200RL_XPROJ_ATTR(8)
        B200RL_LAUNCH(lstm_xproj_tc_kernel<8>, grid, xpk::NT, smem, stream, u, Wt, bt, W_ih, b_ih, b_hh, xproj, rows, Fin);
    } else {
        B200RL_XPROJ_ATTR(32)
        B200RL_LAUNCH(lstm_xproj_tc_kernel<32>, grid, xpk::NT, smem, stream, u, Wt, bt, W_ih, b_ih, b_hh, xproj, rows, Fin);
    }
#undef B200RL_XPROJ_ATTR
    return check_launch("lstm_fwd(xproj)");
}

}  // namespace b200rl
