// LSTM input projection for ALL timesteps at once on tcgen05 (sm_100a):
//
//   xproj[b,t,p] = b_ih[n(p)] + b_hh[n(p)] + sum_k ReLU(u[b,t,:] . Wt[k,:] + bt[k]) * W_ih[n(p)][k]
//
// This half of nn.LSTMCell's gate pre-activation (networks/models.py:67-75) does not depend on the
// recurrence, so it leaves the serial kernel: one [B*T x 64] . [64 x 256] GEMM with full parallelism
// instead of 64 of the 128 K-steps of every timestep of the persistent kernel (whose FFMA loop is
// shared-memory-wavefront bound, profiles/r1f_ncu_summary.md).  Output columns keep the gate-interleaved
// order p = j*4 + g of the saved activations; the recurrent kernel starts its accumulators from them.
//
// Structure (as dx_tc): 8 producer warps + 1 MMA issuer warp.  W_ih is split once into TF32 hi/lo and
// kept resident as the K-major B operand (N = 256, K = 64: 2 x 64 KB).  A producer thread owns one
// row of a 128-row tile: it recomputes x (4-wide obs), splits hi/lo and writes the K-major A operand
// with 16-byte stores.  3xTF32 (hi*hi + hi*lo + lo*hi) into one of TWO 128x256 fp32 accumulators in
// TMEM (all 512 columns), so the epilogue of tile i (tcgen05.ld, + bias, 128 KB of stores) overlaps the
// MMAs of tile i+1.  Bound: the 1 KB/row HBM write.  tcgen05.ld hands each thread one ROW, so direct
// stores touch 32 different 128-byte lines per warp instruction (ncu r1g: LSU wavefronts 72 %, all
// store traffic); the epilogue therefore transposes 32 x 16 blocks through a per-warp smem patch and
// writes 64-byte row segments.  The obs row of the next tile is loaded before the epilogue runs.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

namespace xpk {
constexpr int H = 64, G = 256, NP = 256, NT = NP + 32;
constexpr int TR = 128;
constexpr int W_LBO = 128, W_SBO = (H / 4) * 128;          // B: 16 core matrices along K per 8-column group (2048)
constexpr int W_BYTES = (G / 8) * W_SBO;                   // 65536 per hi / lo
constexpr int A_LBO = 128, A_SBO = (H / 4) * 128;          // A: 16 core matrices along K per 8-row group (2048)
constexpr int A_BYTES = (TR / 8) * A_SBO;                  // 32768 per hi / lo
constexpr int MAXF = 32;
constexpr int SC = 16, SST = SC + 1;                       // epilogue patch: 32 rows x 16 columns, stride 17
}  // namespace xpk

static inline size_t xproj_smem_bytes(int Fin) {
    return 2 * (size_t)xpk::W_BYTES + 2 * (size_t)xpk::A_BYTES + sizeof(float) * (size_t)(Fin * xpk::H + xpk::H + xpk::G + (xpk::NP / 32) * 32 * xpk::SST) + 96;
}

template <int FMAX>
__global__ void __launch_bounds__(xpk::NT, 1)
lstm_xproj_tc_kernel(const float* __restrict__ u, const float* __restrict__ Wt, const float* __restrict__ bt,
                     const float* __restrict__ W_ih, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                     float* __restrict__ xproj, long long rows, int Fin, int igfo) {
    using namespace xpk;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* W_hi = smem;
    unsigned char* W_lo = smem + W_BYTES;
    unsigned char* A_hi = smem + 2 * W_BYTES;
    unsigned char* A_lo = A_hi + A_BYTES;
    float* WtS = reinterpret_cast<float*>(A_lo + A_BYTES);              // WtS[q*64 + k]
    float* bts = WtS + Fin * H;
    float* bsum = bts + H;                                               // bsum[p]
    float* patch = bsum + G;                                             // [warp][32 * SST]
    uint64_t* bars = reinterpret_cast<uint64_t*>(patch + (NP / 32) * 32 * SST);   // a_full | a_empty | acc_done[2] | acc_free[2]
    uint64_t* a_full = bars, *a_empty = bars + 1, *acc_done = bars + 2, *acc_free = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < NP / 32;

    for (int i = tid; i < G * H; i += NT) {          // elem(p, k) = W_ih[n(p)][k], one core matrix (8 p x 4 k) per warp store
        const int k = ((i >> 5) & 15) * 4 + (i & 3), p = (i >> 9) * 8 + ((i >> 2) & 7);
        const int gate = igfo ? ((p & 1) * 2 + ((p >> 1) & 1)) : (p & 3);          // slot -> gate: (i,g,f,o) or (i,f,g,o)
        const float w = W_ih[(size_t)(gate * 64 + (p >> 2)) * H + k];
        float hi, lo;
        tc::split_tf32(w, hi, lo);
        const int o = (p & 7) * 16 + (p >> 3) * W_SBO + (k & 3) * 4 + (k >> 2) * W_LBO;
        *reinterpret_cast<float*>(W_hi + o) = hi;
        *reinterpret_cast<float*>(W_lo + o) = lo;
    }
    for (int i = tid; i < Fin * H; i += NT) { int q = i / H, k = i % H; WtS[i] = Wt[k * Fin + q]; }
    if (tid < H) bts[tid] = bt[tid];
    if (tid < G) { const int p = tid, gate = igfo ? ((p & 1) * 2 + ((p >> 1) & 1)) : (p & 3), n = gate * 64 + (p >> 2); bsum[p] = b_ih[n] + b_hh[n]; }
    if (tid == 0) {
        tc::mbar_init(a_full, NP / 32); tc::mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_done[i], 1); tc::mbar_init(&acc_free[i], NP / 32); }
        tc::fence_mbar_init();
    }
    if (warp == NP / 32) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = tc::make_idesc_tf32(TR, G, false, false);
    const long long ntiles = (rows + TR - 1) / TR;

    if (!producer) {
        if (lane == 0) {
            int n = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
                const int acc = n & 1, use = n >> 1;
                tc::mbar_wait(a_full, (uint32_t)(n & 1));
                if (use >= 1) tc::mbar_wait(&acc_free[acc], (uint32_t)((use - 1) & 1));
                tc::fence_after_sync();
#pragma unroll
                for (int s = 0; s < H / 8; ++s) {
                    const tc::SmemDesc ah = tc::make_desc(A_hi + s * 2 * A_LBO, A_LBO, A_SBO);
                    const tc::SmemDesc al = tc::make_desc(A_lo + s * 2 * A_LBO, A_LBO, A_SBO);
                    const tc::SmemDesc wh = tc::make_desc(W_hi + s * 2 * W_LBO, W_LBO, W_SBO);
                    const tc::SmemDesc wl = tc::make_desc(W_lo + s * 2 * W_LBO, W_LBO, W_SBO);
                    tc::mma_tf32(tmem + acc * G, ah, wh, idesc, s > 0 ? 1u : 0u);
                    tc::mma_tf32(tmem + acc * G, ah, wl, idesc, 1u);
                    tc::mma_tf32(tmem + acc * G, al, wh, idesc, 1u);
                }
                tc::mma_commit(a_empty);
                tc::mma_commit(&acc_done[acc]);
            }
        }
    } else {
        // producer thread: row r of the tile (threads 0..127) / the same rows again for the second half of K (128..255)
        const int r = tid & 127, khalf = tid >> 7;                          // k in [32*khalf, 32*khalf + 32)
        float uq[FMAX];
        auto load_u = [&](long long tile) {
            const long long row = tile * TR + r;
#pragma unroll
            for (int q = 0; q < FMAX; ++q) uq[q] = (q < Fin && row < rows) ? __ldg(u + (size_t)row * Fin + q) : 0.f;
        };
        auto fill = [&](long long tile, int n) {
            if (n >= 1) tc::mbar_wait(a_empty, (uint32_t)((n - 1) & 1));
            const bool ok = tile * TR + r < rows;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const int k0 = khalf * 32 + c4 * 4;
                float4 a = *reinterpret_cast<const float4*>(bts + k0);
#pragma unroll
                for (int q = 0; q < FMAX; ++q) {
                    if (q < Fin) {
                        const float4 w = *reinterpret_cast<const float4*>(WtS + q * H + k0);
                        a.x = fmaf(uq[q], w.x, a.x); a.y = fmaf(uq[q], w.y, a.y); a.z = fmaf(uq[q], w.z, a.z); a.w = fmaf(uq[q], w.w, a.w);
                    }
                }
                const float xv[4] = {ok ? fmaxf(a.x, 0.f) : 0.f, ok ? fmaxf(a.y, 0.f) : 0.f, ok ? fmaxf(a.z, 0.f) : 0.f, ok ? fmaxf(a.w, 0.f) : 0.f};
                float4 hi, lo;
                tc::split_tf32(xv[0], hi.x, lo.x); tc::split_tf32(xv[1], hi.y, lo.y);
                tc::split_tf32(xv[2], hi.z, lo.z); tc::split_tf32(xv[3], hi.w, lo.w);
                const int o = (r & 7) * 16 + (r >> 3) * A_SBO + (k0 >> 2) * A_LBO;
                *reinterpret_cast<float4*>(A_hi + o) = hi;
                *reinterpret_cast<float4*>(A_lo + o) = lo;
            }
            tc::fence_async_smem();
            tc::mbar_arrive_warp(a_full);
        };
        float* P = patch + warp * 32 * SST;
        auto epilogue = [&](long long tile, int n) {
            const int acc = n & 1, use = n >> 1;
            tc::mbar_wait(&acc_done[acc], (uint32_t)(use & 1));
            tc::fence_after_sync();
            const int rbase = 32 * (warp & 3), c_base = (warp >> 2) * 128;          // warps 0-3: columns 0..127, 4-7: 128..255
#ifdef B200RL_EMUL
            const int addr_lane = rbase + lane;
#else
            const int addr_lane = rbase;
#endif
            const int wr = lane >> 4, wc = lane & 15;                                // store role: 2 rows (16 apart) x 16 columns per instruction
            for (int c0 = c_base; c0 < c_base + 128; c0 += 32) {
                float v[32];
                tc::tmem_ld32(tc::tmem_addr(tmem, addr_lane, acc * G + c0), v);
#pragma unroll
                for (int hseg = 0; hseg < 2; ++hseg) {
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < SC; ++i) P[lane * SST + i] = v[hseg * SC + i];
                    __syncwarp();
                    const int col = c0 + hseg * SC + wc;
                    const float bias = bsum[col];
#pragma unroll 8
                    for (int rr = 0; rr < 16; ++rr) {                       // rows rr and rr+16: 16 x 17 = 16 (mod 32) banks apart
                        const long long row = tile * TR + rbase + rr + 16 * wr;
                        if (row < rows) xproj[(size_t)row * G + col] = P[(rr + 16 * wr) * SST + wc] + bias;
                    }
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive_warp(&acc_free[acc]);
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
    if (warp == NP / 32) tc::tmem_dealloc(tmem, 512);
}

int launch_xproj_tc(const float* u, const float* Wt, const float* bt, const float* W_ih, const float* b_ih, const float* b_hh,
                    float* xproj, int B, int T, int Fin, int igfo, int sms, void* stream) {
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
        B200RL_LAUNCH(lstm_xproj_tc_kernel<4>, grid, xpk::NT, smem, stream, u, Wt, bt, W_ih, b_ih, b_hh, xproj, rows, Fin, igfo);
    } else if (Fin <= 8) {
        B200RL_XPROJ_ATTR(8)
        B200RL_LAUNCH(lstm_xproj_tc_kernel<8>, grid, xpk::NT, smem, stream, u, Wt, bt, W_ih, b_ih, b_hh, xproj, rows, Fin, igfo);
    } else {
        B200RL_XPROJ_ATTR(32)
        B200RL_LAUNCH(lstm_xproj_tc_kernel<32>, grid, xpk::NT, smem, stream, u, Wt, bt, W_ih, b_ih, b_hh, xproj, rows, Fin, igfo);
    }
#undef B200RL_XPROJ_ATTR
    return check_launch("lstm_fwd(xproj)");
}

}  // namespace b200rl
