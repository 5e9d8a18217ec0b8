// LSTM weight gradients on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   dW_ih[n][k] = sum_{b,t} dpre[b,t,n] * x[b,t,k]        (x = ReLU torso output, recomputed)
//   dW_hh[n][k] = sum_{b,t} dpre[b,t,n] * h[b,t-1,k]
//   db_ih = db_hh = sum_{b,t} dpre[b,t,:]
//
// i.e. the autograd `AddmmBackward` of every nn.LSTMCell step of the reference
// (networks/models.py:71-75 replayed by loss.backward(), ppo/learning.py:92), as ONE
// reduction GEMM over all B*T rows:  dW^T[128 x 256] = A^T[128 x BT] . D[BT x 256],
// A = [x | h_prev].  Each CTA owns every `gridDim`-th stage of 32 rows, accumulates the
// full 128x256 fp32 tile in TMEM (256 columns) and writes one partial slab at the end.
//
// Precision: fp32 parity needs more than one TF32 pass, so both operands
// are split on the fly, v = hi + lo with hi, lo exactly representable in TF32, and three
// MMAs are issued per K-step: hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-22).
//
// Data path per stage (16 rows = 2 K-steps of 8): all 256 threads load fp32 rows from HBM --
// a quarter-warp takes one full 128-byte line of a row, a warp 4 rows (requests that touch half
// lines of 8 rows ran at 2.3 TB/s: the SM's outstanding-request budget, not bytes, was the limit,
// tools/dx_prof.cu) -- recompute x, split hi/lo and scatter the values with 4-byte stores into the
// canonical no-swizzle K-major UMMA layout, where the contraction index (the row r) is the fast one:
//     elem(mn, r) at (mn%8)*16 + (mn/8)*SBO + (r%4)*4 + (r/4)*LBO        [bytes]
// (measured on B200: kind::tf32 with MN-major descriptors returns zeros, K-major is exact
// -- tools/tc_probe.cu -- so the tile is transposed on the way into shared memory).
// SBO = 528 B (= 16 B modulo 128 B) makes the 32 lanes of every scatter store (4 rows x 8 float4
// columns, one element each) hit 32 distinct banks.
//
// Warp-specialised pipeline (no CTA-wide barrier inside the loop): warps 0-7 are PRODUCERS
// (load / split / scatter a stage, fence.proxy.async, arrive on full[stage]); warp 8 is the MMA
// ISSUER (one lane waits on full[stage], issues 6 tcgen05.mma, tcgen05.commit -> empty[stage]).
// A single thread needs ~50-100 cycles per tcgen05.mma issue (tools/tc_probe2.cu), so the
// issuer must not be one of the filling threads.  3-stage ring, 189 KB of smem.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

namespace wtc {
constexpr int H = 64, G = 256, KC = 128, NP = 256, NT = NP + 32;   // producers + issuer warp
constexpr int SR = 16;                          // rows per stage (2 MMA K-steps of 8)
constexpr int NSTAGE = 2;                         // two CTAs per SM (2 x 101 KB of stages, 2 x 256 TMEM columns): 4 producer warps per sub-partition
constexpr int LBO = 128;                        // K direction: next core matrix (4 rows)
constexpr int SBO = 528;                        // MN direction: next group of 8 mn (4 core matrices + 16 B skew) = 16 B mod 128 B
constexpr int A_TILE = (KC / 8) * SBO;          // 8448
constexpr int D_TILE = (G / 8) * SBO;           // 16896
constexpr int STAGE = 2 * A_TILE + 2 * D_TILE;  // A_hi | A_lo | D_hi | D_lo = 50688 B
constexpr int MAXF = 32;
static_assert((SR / 4 - 1) * LBO + 128 <= SBO, "core matrices of one mn-group must fit below SBO");
__device__ __forceinline__ int koff(int mn, int r) { return (mn & 7) * 16 + (mn >> 3) * SBO + (r & 3) * 4 + (r >> 2) * LBO; }
}  // namespace wtc

static inline size_t wgrad_tc_smem_bytes(int Fin) {
    return wtc::NSTAGE * (size_t)wtc::STAGE + sizeof(float) * (size_t)(Fin * wtc::H + wtc::H) + 96;
}

__global__ void __launch_bounds__(wtc::NT, 2)
lstm_wgrad_tc_kernel(const float* __restrict__ dpre, const float* __restrict__ h, const float* __restrict__ hx0,
                     const float* __restrict__ u, const float* __restrict__ Wt, const float* __restrict__ bt,
                     float* __restrict__ partials, long long P,
                     long long off_Wih, long long off_Whh, long long off_bih, long long off_bhh,
                     int B, int T, int Fin) {
    using namespace wtc;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* stage0 = smem;                                   // 1024-aligned by construction (offset 0)
    float* WtS = reinterpret_cast<float*>(smem + NSTAGE * STAGE);   // WtS[q*64 + j]
    float* bts = WtS + Fin * H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bts + H);          // full[NSTAGE] | empty[NSTAGE] | done
    uint64_t* full = bars, *empty = bars + NSTAGE, *done = bars + 2 * NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < NP / 32;
    const int rq = lane >> 3, cq = lane & 7;                         // row inside a group of 4, float4 inside a 128-byte line

    for (int i = tid; i < Fin * H; i += NT) { int q = i / H, j = i % H; WtS[i] = Wt[j * Fin + q]; }
    if (tid < H) bts[tid] = bt[tid];
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], NP / 32); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(done, 1);
        tc::fence_mbar_init();
    }
    if (warp == NP / 32) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = tc::make_idesc_tf32(KC, G, false, false);

    const long long rows = (long long)B * T;
    const long long nstages = (rows + SR - 1) / SR;
    float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f);               // bias gradient of gate columns (warp*8 + cq)*4 .. +3

    // Register-staged loads: two register sets; the HBM reads of the CTA's next stage are issued before the current one
    // is split and scattered (the co-resident second CTA keeps another stage in flight on the same SM).
    constexpr int UP = 4;                           // torso inputs prefetched in registers when Fin <= UP
    struct StageRegs { float4 d[4]; float4 h; float u[UP]; bool ok; };
    StageRegs ring[2];
    auto issue_loads = [&](StageRegs& R, long long stg) {
        const long long row0 = stg * SR;
        const bool live = stg < nstages;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const int c4 = warp * 8 + cq;
            const long long row = row0 + i4 * 4 + rq;
            R.d[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && row < rows) R.d[i4] = *reinterpret_cast<const float4*>(dpre + (size_t)row * G + c4 * 4);
        }
        const int c4 = (warp & 1) * 8 + cq;
        const long long row = row0 + (warp >> 1) * 4 + rq;
        R.h = make_float4(0.f, 0.f, 0.f, 0.f);
        R.ok = live && row < rows;
#pragma unroll
        for (int q = 0; q < UP; ++q) R.u[q] = 0.f;
        if (R.ok) {
            const unsigned urow = (unsigned)row, b = urow / (unsigned)T, t = urow - b * (unsigned)T;   // 32-bit: B*T < 2^31 is checked by the launcher
            const float* hp = (t == 0) ? hx0 + (size_t)b * H : h + (size_t)(row - 1) * H;
            R.h = *reinterpret_cast<const float4*>(hp + c4 * 4);
            if (Fin <= UP) {
#pragma unroll
                for (int q = 0; q < UP; ++q) if (q < Fin) R.u[q] = __ldg(u + (size_t)row * Fin + q);
            }
        }
    };

    int it = 0;
    auto process = [&](const StageRegs& R, long long st) {
        const int buf = it % NSTAGE, use = it / NSTAGE;
        unsigned char* sb = stage0 + (size_t)buf * STAGE;
        unsigned char* A_hi = sb, *A_lo = sb + A_TILE, *D_hi = sb + 2 * A_TILE, *D_lo = sb + 2 * A_TILE + D_TILE;
        const long long row0 = st * SR;
        if (use >= 1) tc::mbar_wait(&empty[buf], (uint32_t)((use - 1) & 1));  // MMAs of the previous use of this buffer are done

        // ---- D tile: 16 rows x 64 float4.  Thread: rows i4*4 + rq, float4 column warp*8 + cq
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const int c4 = warp * 8 + cq;
            const int r = i4 * 4 + rq;
            const float4 v = R.d[i4];
            dbacc.x += v.x; dbacc.y += v.y; dbacc.z += v.z; dbacc.w += v.w;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float hi, lo;
                tc::split_tf32(vv[e], hi, lo);
                const int o = koff(c4 * 4 + e, r);
                *reinterpret_cast<float*>(D_hi + o) = hi;
                *reinterpret_cast<float*>(D_lo + o) = lo;
            }
        }
        // ---- A tile: [x_t (k<64) | h_{t-1} (k>=64)], 16 float4 each per row; thread: row (warp>>1)*4 + rq
        {
            const int r = (warp >> 1) * 4 + rq, c4 = (warp & 1) * 8 + cq;                // c4 in 0..15
            const long long row = row0 + r;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 hv = R.h;
            if (R.ok) {
                float4 a = *reinterpret_cast<const float4*>(bts + c4 * 4);
                if (Fin <= UP) {
#pragma unroll
                    for (int q = 0; q < UP; ++q) {
                        if (q < Fin) {
                            const float4 w = *reinterpret_cast<const float4*>(WtS + q * H + c4 * 4);
                            a.x = fmaf(R.u[q], w.x, a.x); a.y = fmaf(R.u[q], w.y, a.y); a.z = fmaf(R.u[q], w.z, a.z); a.w = fmaf(R.u[q], w.w, a.w);
                        }
                    }
                } else {
                    const float* ur = u + (size_t)row * Fin;
                    for (int q = 0; q < Fin; ++q) {
                        const float uv = __ldg(ur + q);
                        const float4 w = *reinterpret_cast<const float4*>(WtS + q * H + c4 * 4);
                        a.x = fmaf(uv, w.x, a.x); a.y = fmaf(uv, w.y, a.y); a.z = fmaf(uv, w.z, a.z); a.w = fmaf(uv, w.w, a.w);
                    }
                }
                xv = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
            }
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, hs[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float hi, lo;
                tc::split_tf32(xs[e], hi, lo);
                int o = koff(c4 * 4 + e, r);
                *reinterpret_cast<float*>(A_hi + o) = hi;
                *reinterpret_cast<float*>(A_lo + o) = lo;
                tc::split_tf32(hs[e], hi, lo);
                o = koff(H + c4 * 4 + e, r);
                *reinterpret_cast<float*>(A_hi + o) = hi;
                *reinterpret_cast<float*>(A_lo + o) = lo;
            }
        }
        tc::fence_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
        tc::mbar_arrive_warp(&full[buf]);
        ++it;
    };

    if (producer) {
        const long long g = gridDim.x;
        long long st = blockIdx.x;
        issue_loads(ring[0], st);
        while (st < nstages) {
            issue_loads(ring[1], st + g); process(ring[0], st); st += g; if (st >= nstages) break;
            issue_loads(ring[0], st + g); process(ring[1], st); st += g;
        }
    } else if (lane == 0) {
        // ---- MMA issuer: stage descriptors differ only in their start address (low 14 bits, 16-byte units)
        int n = 0;
        for (long long st = blockIdx.x; st < nstages; st += gridDim.x, ++n) {
            const int buf = n % NSTAGE, use = n / NSTAGE;
            unsigned char* sb = stage0 + (size_t)buf * STAGE;
            unsigned char* A_hi = sb, *A_lo = sb + A_TILE, *D_hi = sb + 2 * A_TILE, *D_lo = sb + 2 * A_TILE + D_TILE;
            tc::mbar_wait(&full[buf], (uint32_t)(use & 1));
            tc::fence_after_sync();
#pragma unroll
            for (int s = 0; s < SR / 8; ++s) {
                const int ko = s * 2 * LBO;                                           // 8 rows = 2 core matrices along K
                const tc::SmemDesc ah = tc::make_desc(A_hi + ko, LBO, SBO);
                const tc::SmemDesc al = tc::make_desc(A_lo + ko, LBO, SBO);
                const tc::SmemDesc dh = tc::make_desc(D_hi + ko, LBO, SBO);
                const tc::SmemDesc dl = tc::make_desc(D_lo + ko, LBO, SBO);
                tc::mma_tf32(tmem, ah, dh, idesc, (n > 0 || s > 0) ? 1u : 0u);
                tc::mma_tf32(tmem, ah, dl, idesc, 1u);
                tc::mma_tf32(tmem, al, dh, idesc, 1u);
            }
            tc::mma_commit(&empty[buf]);
        }
        tc::mma_commit(done);
    }
    if (producer) {
    tc::mbar_wait(done, 0);
    tc::fence_after_sync();

    // ---- epilogue: TMEM lane = k (A column), TMEM column = p = j*4 + g ; PyTorch row n = g*64 + j
    float* out = partials + (size_t)blockIdx.x * P;
    {
        const int lane_base = 32 * (warp & 3);
        const int k = lane_base + lane;                    // 0..127
        float* dst = out + (k < H ? off_Wih + k : off_Whh + (k - H));
#ifdef B200RL_EMUL
        const int addr_lane = k;
#else
        const int addr_lane = lane_base;
#endif
        for (int c0 = (warp >> 2) * 128; c0 < (warp >> 2) * 128 + 128; c0 += 32) {
            float r[32];
            tc::tmem_ld32(tc::tmem_addr(tmem, addr_lane, c0), r);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int p = c0 + i, n = (p & 3) * H + (p >> 2);
                dst[(size_t)n * H] = it > 0 ? r[i] : 0.f;
            }
        }
    }
    // ---- bias gradient: reduce the 4 row-lanes that share a float4 column
    {
        float4 s = dbacc;
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            s.x += __shfl_xor_sync(kFull, s.x, o); s.y += __shfl_xor_sync(kFull, s.y, o);
            s.z += __shfl_xor_sync(kFull, s.z, o); s.w += __shfl_xor_sync(kFull, s.w, o);
        }
        if (rq == 0) {
            const int c4 = warp * 8 + cq;                  // p = c4*4 + g  ->  j = c4, n = g*64 + c4
            const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int g = 0; g < 4; ++g) { out[off_bih + g * H + c4] = v[g]; out[off_bhh + g * H + c4] = v[g]; }
        }
    }
    }   // producer
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NP / 32) tc::tmem_dealloc(tmem, 256);
}

int launch_wgrad_tc(const float* dpre, const float* h, const float* hx0, const float* u,
                    const float* Wt, const float* bt, float* partials, int nparts, long long P,
                    long long off_Wih, long long off_Whh, long long off_bih, long long off_bhh,
                    int B, int T, int Fin, void* stream) {
    B200RL_REQUIRE(Fin >= 1 && Fin <= wtc::MAXF, "lstm_wgrad(tc): Fin=%d out of range", Fin);
    B200RL_REQUIRE((long long)B * T < (1LL << 31), "lstm_wgrad(tc): B*T must be below 2^31");
    const size_t smem = wgrad_tc_smem_bytes(Fin);
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(lstm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lstm_wgrad(tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
    B200RL_LAUNCH(lstm_wgrad_tc_kernel, nparts, wtc::NT, smem, stream, dpre, h, hx0, u, Wt, bt, partials, P,
                  off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin);
    return check_launch("lstm_wgrad(tc)");
}

}  // namespace b200rl
