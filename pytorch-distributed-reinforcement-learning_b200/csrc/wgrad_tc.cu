// LSTM weight gradients on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   dW_ih[n][k] = sum_{b,t} dpre[b,t,n] * x[b,t,k]        (x = ReLU torso output, recomputed)
//   dW_hh[n][k] = sum_{b,t} dpre[b,t,n] * h[b,t-1,k]
//   db_ih = db_hh = sum_{b,t} dpre[b,t,:]
//
// i.e. the autograd `AddmmBackward` of every nn.LSTMCell step of the reference
// (networks/models.py:71-75 replayed by loss.backward(), ppo/learning.py:92), as ONE
// reduction GEMM over all B*T rows:  dW^T[128 x 256] = A^T[128 x BT] . D[BT x 256],
// A = [x | h_prev].  Each CTA owns every `gridDim`-th stage of 16 rows, accumulates the
// full 128x256 fp32 tile in TMEM (256 columns) and writes one partial slab at the end.
//
// Precision: fp32 parity needs more than one TF32 pass, so both operands are split on the fly,
// v = hi + lo with hi exactly representable in TF32, and three products are accumulated into the same fp32 tile:
//     hi*hi          as kind::tf32 (K = 8 per MMA: two per stage),
//     v*lo + lo*v    with all four factors rounded to bf16, as kind::f16 (K = 16 per MMA, twice the TF32 rate): ONE
//                    contraction of length 32 per stage over the concatenated operands  A' = [v | lo],  D' = [lo | v]
// -- the cross terms are 2^-11 of the product, so their bf16 rounding (2^-9) lands at 2^-20 of it, and they carry the
// dropped lo*lo term twice (2^-21).  Four tensor-pipe units per 16 rows instead of the six of three TF32 passes
// (ncu r2j: the tensor pipe was 53 % busy and the 6 MMAs read 72 KB of shared memory per stage; now 36 %, 48 KB).
//
// Data path per stage: all 256 producer threads load fp32 from HBM -- a warp reads 512 contiguous bytes of a dpre row or
// one full 128-byte line of an h row per instruction (4.7 TB/s for the bare load loop against 4.3 TB/s when a warp
// instruction touches four rows, and 2.3 TB/s with half lines: tools/wgrad_prof.cu, tools/dx_prof.cu) -- such that EVERY
// THREAD OWNS FOUR CONSECUTIVE ROWS OF ITS COLUMNS.  In the canonical no-swizzle K-major UMMA layouts the contraction index
// is the fast one and a 16-byte core-matrix row holds 4 tf32 or 8 bf16 of it:
//     tf32 tile:  elem(mn, r)  at (mn%8)*16 + (mn/8)*SBO + (r/4)*LBO + (r%4)*4          [bytes], r = row of the stage
//     bf16 tile:  elem(mn, kk) at (mn%8)*16 + (mn/8)*SBO + (kk/8)*LBO + (kk%8)*2
// and the order of the contraction index is free as long as both operands agree: the bf16 tiles use
//     kk = (r/4)*8 + (r%4)      for the first factor of a row ( v of A', lo of D'),   kk + 4 for the second (lo of A', v of D'),
// so the four rows of one column are ONE 16-byte store into the tf32 tile and ONE 16-byte store AT THE SAME OFFSET into the
// bf16 tile: 12 conflict-free 16-byte stores per thread and stage instead of 48 scalar scatters.
// (measured on B200: kind::tf32 with MN-major descriptors returns zeros, K-major is exact -- tools/tc_probe.cu -- so the
// tile is transposed on the way into shared memory; a raw TMA tile cannot feed these MMAs.)
// SBO = 528 B (= 16 B modulo 128 B) makes the 8 lanes of every store wavefront hit distinct 16-byte bank groups.
//
// Warp-specialised pipeline (no CTA-wide barrier inside the loop): warps 0-7 are PRODUCERS
// (load / split / store a stage, fence.proxy.async, arrive on full[stage]); warp 8 is the MMA
// ISSUER (one lane waits on full[stage], issues 4 tcgen05.mma, tcgen05.commit -> empty[stage]).
// A single thread needs ~50-100 cycles per tcgen05.mma issue (tools/tc_probe2.cu), so the
// issuer must not be one of the filling threads.  2-stage ring, two CTAs per SM.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

#ifndef B200RL_WGPROBE
#define B200RL_WGPROBE 0     // tools/wgrad_prof.cu only: 1 loads only, 2 no MMAs, 3 no HBM loads (timing experiments; results invalid)
#endif

#ifndef B200RL_WG_BF16X
#define B200RL_WG_BF16X 0    // 1: cross terms v*lo + lo*v as ONE bf16 contraction (4 tensor units per stage, 2^-20); 0: two more TF32 passes (6 units, 2^-22)
#endif

namespace wtc {
constexpr int kProbe = B200RL_WGPROBE;
constexpr bool kBf16X = B200RL_WG_BF16X != 0;
constexpr int H = 64, G = 256, KC = 128, NP = 256, NT = NP + 32;   // producers + issuer warp
constexpr int SR = 16;                          // rows per stage (2 TF32 K-steps of 8; 2 bf16 K-steps of 16 over [v | lo])
constexpr int NSTAGE = 2;                       // two CTAs per SM (2 x 101 KB of stages, 2 x 256 TMEM columns): 4 producer warps per sub-partition
constexpr int LBO = 128;                        // K direction: next core matrix (4 rows of the stage)
constexpr int SBO = 528;                        // MN direction: next group of 8 mn (4 core matrices + 16 B skew) = 16 B mod 128 B
constexpr int A_TILE = (KC / 8) * SBO;          // 8448
constexpr int D_TILE = (G / 8) * SBO;           // 16896
constexpr int O_AH = 0, O_DH = A_TILE, O_AB = O_DH + D_TILE, O_DB = O_AB + A_TILE;   // A_hi | D_hi (tf32) | A' | D' (bf16)
constexpr int STAGE = O_DB + D_TILE;            // 50688 B
constexpr int MAXF = 32;
// (two CTAs of 9 warps per SM: 96 registers per thread is the ceiling -- with 112 only one CTA is resident, measured)
constexpr int UP = 4;                           // torso widths up to UP: weight column in registers, inputs by shuffle; wider: inputs staged in shared memory
static_assert((SR / 4 - 1) * LBO + 128 <= SBO, "core matrices of one mn-group must fit below SBO");
static_assert(STAGE % 16 == 0 && O_DH % 16 == 0 && O_AB % 16 == 0 && O_DB % 16 == 0, "descriptor alignment");
__device__ __forceinline__ int koff(int mn, int r) { return (mn & 7) * 16 + (mn >> 3) * SBO + (r & 3) * 4 + (r >> 2) * LBO; }
}  // namespace wtc

static inline size_t wgrad_tc_smem_bytes(int Fin) {
    return wtc::NSTAGE * (size_t)wtc::STAGE + sizeof(float) * (size_t)(Fin * wtc::H + wtc::H + (Fin > wtc::UP ? 2 * wtc::SR * Fin : 0)) + 96;   // (WIDE <=> Fin > UP)
}

template <bool WIDE>          // WIDE: torso input wider than UP (staged through shared memory); else the 4 x Fin inputs of a row group ride in one register per lane
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
    float* usm = bts + H;                                           // WIDE: the stage's 16 x Fin torso inputs, double-buffered
    uint64_t* bars = reinterpret_cast<uint64_t*>(usm + (WIDE ? 2 * SR * Fin : 0));   // full[NSTAGE] | empty[NSTAGE] | done
    uint64_t* full = bars, *empty = bars + NSTAGE, *done = bars + 2 * NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < NP / 32;
    const int drg = (warp >> 1) & 3, dc4 = (warp & 1) * 32 + lane;   // D tile: row group of 4, float4 column (a warp = half a dpre row)
    const int ak = tid & 63, arg = (tid >> 6) & 3;                   // A tile: hidden column, row group of 4

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
    const uint32_t idescb = tc::make_idesc_bf16(KC, G);

    const long long rows = (long long)B * T;
    const long long nstages = (rows + SR - 1) / SR;
    float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f);               // bias gradient of gate columns dc4*4 .. +3 over this thread's rows
    constexpr bool ulane = !WIDE;                                  // the 4 x Fin torso inputs of a row group: one per lane
    const int nu = SR * Fin;                                       // wider torsos: the stage's 16 x Fin inputs, up to two per thread

    // Register-staged loads: two register sets, both in flight except while their own stage is processed (a set is refilled,
    // two stages ahead, as soon as it has been written to shared memory).  The common case -- a stage with all 16 rows in
    // range and no sequence start among them -- runs without a single bounds check or row/time division: three running
    // pointers, immediate offsets.  `mode`: 0 nothing, 1 fast, 2 generic.
    struct StageRegs { float4 d[4]; float hp[4]; float u, u2; int mode; };      // (u2: WIDE only, dead otherwise)
    StageRegs ring[2];
    const long long stride = (long long)gridDim.x * SR;            // rows between two stages of this CTA
    const unsigned dstep = (unsigned)(stride * G * 4), hstep = (unsigned)(stride * H * 4), ustep = (unsigned)(stride * Fin * 4);   // bytes (< 2^32: the launcher bounds the grid)
    const int nst = (int)nstages;
    const float* dptr = dpre + ((size_t)blockIdx.x * SR + drg * 4) * G + dc4 * 4;
    const float* hptr = h + ((long long)blockIdx.x * SR + arg * 4 - 1) * H + ak;
    const float* uptr = ulane ? u + ((size_t)blockIdx.x * SR + arg * 4) * Fin + lane : u + (size_t)blockIdx.x * SR * Fin + tid;
    const unsigned tstep = (unsigned)(stride % T);
    unsigned tpos = (unsigned)(((long long)blockIdx.x * SR) % T);  // time index of the stage's first row
    int lst = blockIdx.x;                                          // the next stage to load
    auto issue_loads = [&](StageRegs& R) {
        const long long row0 = (long long)lst * SR;
        const bool live = lst < nst;
        const bool whole = row0 + SR <= rows;
        const bool has_t0 = tpos == 0 || tpos + SR > (unsigned)T;  // a row with t == 0 takes h_{t-1} from hx0
        R.mode = !live ? 0 : (whole && !has_t0) ? 1 : 2;
        if (kProbe == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { R.d[j] = make_float4(1.f, 2.f, 3.f, 4.f); R.hp[j] = 1.f; }
            R.u = R.u2 = 1.f;
        } else if (R.mode == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) R.d[j] = *reinterpret_cast<const float4*>(dptr + j * G);
#pragma unroll
            for (int j = 0; j < 4; ++j) R.hp[j] = __ldg(hptr + j * H);
            if (ulane) R.u = lane < 4 * Fin ? __ldg(uptr) : 0.f;
            else { R.u = tid < nu ? __ldg(uptr) : 0.f; R.u2 = tid + NP < nu ? __ldg(uptr + NP) : 0.f; }
        } else {
            const long long r = row0 + drg * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                R.d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && r + j < rows) R.d[j] = *reinterpret_cast<const float4*>(dptr + j * G);
            }
            const long long ra = row0 + arg * 4;
            unsigned b = 0, t = 0;
            if (live && ra < rows) { const unsigned ur = (unsigned)ra; b = ur / (unsigned)T; t = ur - b * (unsigned)T; }   // 32-bit: B*T < 2^31 is checked by the launcher
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                R.hp[j] = 0.f;
                if (live && ra + j < rows) R.hp[j] = (t == 0) ? __ldg(hx0 + (size_t)b * H + ak) : __ldg(hptr + j * H);
                if (++t == (unsigned)T) { t = 0; ++b; }
            }
            R.u = R.u2 = 0.f;
            if (ulane) { if (live && lane < 4 * Fin && ra * Fin + lane < rows * Fin) R.u = __ldg(uptr); }
            else if (live) {
                if (tid < nu && row0 * Fin + tid < rows * Fin) R.u = __ldg(uptr);
                if (tid + NP < nu && row0 * Fin + tid + NP < rows * Fin) R.u2 = __ldg(uptr + NP);
            }
        }
        dptr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(dptr) + dstep);
        hptr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(hptr) + hstep);
        uptr = reinterpret_cast<const float*>(reinterpret_cast<const char*>(uptr) + ustep);
        lst += gridDim.x;
        tpos += tstep; if (tpos >= (unsigned)T) tpos -= (unsigned)T;
    };

    // per-thread tile offsets (constant over the stages; the same in the tf32 and in the bf16 tile)
    const int oD = koff(dc4 * 4, drg * 4);                         // gate columns dc4*4 + e are 16 bytes apart
    const int oA = koff(ak, arg * 4);
    constexpr int HK = (H >> 3) * SBO;                             // column H + ak is H/8 mn-groups further
    // four consecutive rows of one column: the TF32 parts into the tf32 tile; the bf16 roundings of the values and of their
    // remainders into the bf16 tile (`lo_first`: D' = [lo | v], else A' = [v | lo])
    auto put4 = [](unsigned char* p_hi, unsigned char* p_b, bool lo_first, float v0, float v1, float v2, float v3) {
        const float h0 = __uint_as_float(__float_as_uint(v0) & tc::kTf32Mask), h1 = __uint_as_float(__float_as_uint(v1) & tc::kTf32Mask);
        const float h2 = __uint_as_float(__float_as_uint(v2) & tc::kTf32Mask), h3 = __uint_as_float(__float_as_uint(v3) & tc::kTf32Mask);
        *reinterpret_cast<float4*>(p_hi) = make_float4(h0, h1, h2, h3);
        if (!kBf16X) { *reinterpret_cast<float4*>(p_b) = make_float4(v0 - h0, v1 - h1, v2 - h2, v3 - h3); return; }   // the lo tile (the tensor core drops the bits below TF32)
        const uint32_t a0 = tc::pack_bf16x2_rn(v0, v1), a1 = tc::pack_bf16x2_rn(v2, v3);
        const uint32_t l0 = tc::pack_bf16x2_rn(v0 - h0, v1 - h1), l1 = tc::pack_bf16x2_rn(v2 - h2, v3 - h3);
        float4 w;
        if (lo_first) w = make_float4(__uint_as_float(l0), __uint_as_float(l1), __uint_as_float(a0), __uint_as_float(a1));
        else w = make_float4(__uint_as_float(a0), __uint_as_float(a1), __uint_as_float(l0), __uint_as_float(l1));
        *reinterpret_cast<float4*>(p_b) = w;
    };

    const float bk = bts[ak];
    float wk[UP];
#pragma unroll
    for (int q = 0; q < UP; ++q) wk[q] = q < Fin ? WtS[q * H + ak] : 0.f;

    int nproc = 0;
    int pst = blockIdx.x;                                         // the next stage to process
    auto process = [&](const StageRegs& R, int buf) {
        const int use = nproc >> 1;
        unsigned char* sb = stage0 + (size_t)buf * STAGE;
        if (kProbe == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { dbacc.x += R.d[j].x + R.hp[j]; dbacc.y += R.d[j].y; dbacc.z += R.d[j].z; dbacc.w += R.d[j].w + R.u; }
            ++nproc; pst += gridDim.x;
            return;
        }
        if (use >= 1) tc::mbar_wait(&empty[buf], (uint32_t)((use - 1) & 1));  // MMAs of the previous use of this buffer are done

        // ---- D tile: thread = rows drg*4 .. +3 of the four gate columns dc4*4 .. +3
        {
            dbacc.x += (R.d[0].x + R.d[1].x) + (R.d[2].x + R.d[3].x);
            dbacc.y += (R.d[0].y + R.d[1].y) + (R.d[2].y + R.d[3].y);
            dbacc.z += (R.d[0].z + R.d[1].z) + (R.d[2].z + R.d[3].z);
            dbacc.w += (R.d[0].w + R.d[1].w) + (R.d[2].w + R.d[3].w);
            unsigned char* ph = sb + O_DH + oD, *pb = sb + O_DB + oD;
            put4(ph, pb, true, R.d[0].x, R.d[1].x, R.d[2].x, R.d[3].x);
            put4(ph + 16, pb + 16, true, R.d[0].y, R.d[1].y, R.d[2].y, R.d[3].y);
            put4(ph + 32, pb + 32, true, R.d[0].z, R.d[1].z, R.d[2].z, R.d[3].z);
            put4(ph + 48, pb + 48, true, R.d[0].w, R.d[1].w, R.d[2].w, R.d[3].w);
        }
        // ---- A tile: [x_t (k < 64) | h_{t-1} (k >= 64)]; thread = rows arg*4 .. +3 of column ak of both halves
        {
            float x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = bk;
            if (ulane) {                                           // torso column ak in registers, the 4 x Fin inputs by shuffle
#pragma unroll
                for (int q = 0; q < UP; ++q) {
                    if (q < Fin) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] = fmaf(__shfl_sync(kFull, R.u, j * Fin + q), wk[q], x[j]);
                    }
                }
            } else {                                               // the stage's inputs through shared memory (broadcast reads)
                float* us = usm + (nproc & 1) * nu;
                if (tid < nu) us[tid] = R.u;
                if (tid + NP < nu) us[tid + NP] = R.u2;
                tc::named_bar_sync(1, NP);
                const float* ur = us + arg * 4 * Fin;
#pragma unroll 4
                for (int q = 0; q < Fin; ++q) {
                    const float w = WtS[q * H + ak];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = fmaf(ur[j * Fin + q], w, x[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
            if (R.mode != 1) {
                const long long r = (long long)pst * SR + arg * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) if (r + j >= rows) x[j] = 0.f;
            }
            put4(sb + O_AH + oA, sb + O_AB + oA, false, x[0], x[1], x[2], x[3]);
            put4(sb + O_AH + oA + HK, sb + O_AB + oA + HK, false, R.hp[0], R.hp[1], R.hp[2], R.hp[3]);
        }
        tc::fence_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
        tc::mbar_arrive_warp(&full[buf]);
        ++nproc;
        pst += gridDim.x;
    };

    if (producer) {
        issue_loads(ring[0]);
        issue_loads(ring[1]);
        while (pst < nst) {
            process(ring[0], 0); issue_loads(ring[0]); if (pst >= nst) break;
            process(ring[1], 1); issue_loads(ring[1]);
        }
    } else if (lane == 0) {
        // ---- MMA issuer: stage descriptors differ only in their start address (low 14 bits, 16-byte units)
        int n = 0;
        for (long long st = blockIdx.x; st < nstages; st += gridDim.x, ++n) {
            const int buf = n % NSTAGE, use = n / NSTAGE;
            unsigned char* sb = stage0 + (size_t)buf * STAGE;
            if (kProbe == 1) continue;
            tc::mbar_wait(&full[buf], (uint32_t)(use & 1));
            tc::fence_after_sync();
            if (kProbe == 2) { tc::mbar_arrive(&empty[buf]); continue; }
#pragma unroll
            for (int s = 0; s < SR / 8; ++s) {
                const int ko = s * 2 * LBO;                        // 8 rows = 2 core matrices along K (tf32) = 16 of the 32 bf16 contraction steps
                tc::mma_tf32(tmem, tc::make_desc(sb + O_AH + ko, LBO, SBO), tc::make_desc(sb + O_DH + ko, LBO, SBO), idesc, (n > 0 || s > 0) ? 1u : 0u);
                if (kBf16X) {
                    tc::mma_bf16(tmem, tc::make_desc(sb + O_AB + ko, LBO, SBO), tc::make_desc(sb + O_DB + ko, LBO, SBO), idescb, 1u);
                } else {
                    tc::mma_tf32(tmem, tc::make_desc(sb + O_AH + ko, LBO, SBO), tc::make_desc(sb + O_DB + ko, LBO, SBO), idesc, 1u);
                    tc::mma_tf32(tmem, tc::make_desc(sb + O_AB + ko, LBO, SBO), tc::make_desc(sb + O_DH + ko, LBO, SBO), idesc, 1u);
                }
            }
            tc::mma_commit(&empty[buf]);
        }
        if (kProbe == 1 || kProbe == 2) tc::mbar_arrive(done); else
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
                dst[(size_t)n * H] = nproc > 0 ? r[i] : 0.f;
            }
        }
    }
    // ---- bias gradient: the four row groups of a gate column live in four warps -> through the (now idle) stage memory
    {
        float* red = reinterpret_cast<float*>(stage0);                   // red[drg][256 gate columns]
        *reinterpret_cast<float4*>(red + drg * G + dc4 * 4) = dbacc;
        tc::named_bar_sync(1, NP);
        const int p = tid;                                               // p = j*4 + g  ->  n = g*64 + j
        const float v = (red[p] + red[G + p]) + (red[2 * G + p] + red[3 * G + p]);
        const int n = (p & 3) * H + (p >> 2);
        out[off_bih + n] = v;
        out[off_bhh + n] = v;
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
    B200RL_REQUIRE(nparts >= 1 && nparts <= 65536, "lstm_wgrad(tc): nparts=%d out of range (32-bit byte strides)", nparts);
    const size_t smem = wgrad_tc_smem_bytes(Fin);
    if (Fin > wtc::UP) {
#ifndef B200RL_EMUL
        cudaError_t e = cudaFuncSetAttribute(lstm_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error("lstm_wgrad(tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
        B200RL_LAUNCH(lstm_wgrad_tc_kernel<true>, nparts, wtc::NT, smem, stream, dpre, h, hx0, u, Wt, bt, partials, P,
                      off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin);
    } else {
#ifndef B200RL_EMUL
        cudaError_t e = cudaFuncSetAttribute(lstm_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error("lstm_wgrad(tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
        B200RL_LAUNCH(lstm_wgrad_tc_kernel<false>, nparts, wtc::NT, smem, stream, dpre, h, hx0, u, Wt, bt, partials, P,
                      off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin);
    }
    return check_launch("lstm_wgrad(tc)");
}

}  // namespace b200rl
