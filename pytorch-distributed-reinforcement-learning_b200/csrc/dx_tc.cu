// Torso gradients through the LSTM input projection on tcgen05, and the head gradients.
//
//   dx[r][k]   = sum_p dpre[r][p] * W_ih[n(p)][k]        (r over all B*T rows, K = 256, N = 64)
//   dxr        = dx * [x > 0]                            (ReLU of the torso, x recomputed from u)
//   dWt[j][q]  = sum_r dxr[r][j] * u[r][q] ;  dbt[j] = sum_r dxr[r][j] ;  du[r][q] = sum_j dxr[r][j] Wt[j][q]
//   dW_head[c][k] = sum_r dhead[r][c] * h[r][k] ;  db_head[c] = sum_r dhead[r][c]
//
// (autograd of `x = body(obs)` feeding nn.LSTMCell and of the value/logits Linear layers,
//  networks/models.py:67-69,77-78, replayed by loss.backward(), ppo/learning.py:92.)
//
// dx kernel: a CTA owns 128-row tiles.  W_ih is split once into TF32 hi/lo and kept resident
// in shared memory as the K-major B operand (2 x 64 KB).  dpre is streamed in chunks of 32
// gate columns: every float4 of a row IS one 16-byte row of a K-major core matrix, so the A
// operand is written with conflict-free 16-byte stores (hi and lo), no transposition.
// 3xTF32: D_hi*W_hi + D_hi*W_lo + D_lo*W_hi accumulate into a 128x64 fp32 tile in TMEM.
// Epilogue: tcgen05.ld -> ReLU mask -> shared tile -> per-column reductions on FFMA.
// Warp-specialised like wgrad_tc: warps 0-7 produce stages and run the epilogue, warp 8 issues the MMAs;
// full/empty mbarriers per stage, tile_done / tmem_free per tile, a named barrier among the producers.
#include "common.cuh"
#include "tc05.cuh"
#include "launchers.cuh"

namespace b200rl {

// phase timeline for tools/dx_prof.cu (never compiled into the product library)
#ifdef B200RL_DXPROF
__device__ unsigned long long g_dxprof[16];
#define DXP_INIT unsigned long long dxp_prev = clock64();
#define DXP(i) do { if (tid == 0) { unsigned long long now_ = clock64(); atomicAdd(&g_dxprof[i], now_ - dxp_prev); dxp_prev = now_; } } while (0)
#else
#define DXP_INIT
#define DXP(i) do { } while (0)
#endif

namespace dxk {
constexpr int H = 64, G = 256, NP = 256, NT = NP + 32;     // 8 producer warps + the MMA issuer warp
constexpr int TR = 128;                     // rows per tile (UMMA M)
constexpr int CK = 32;                      // gate columns per chunk (4 MMA K-steps)
constexpr int NCH = G / CK;                 // 8 chunks per tile
constexpr int W_LBO = 128, W_SBO = (G / 4) * 128;          // B: 64 core matrices along K per 8-k group (8192)
constexpr int W_BYTES = (H / 8) * W_SBO;                   // 65536 per hi / lo
// A: 8 core matrices along K per 8-row group.  K-adjacent core matrices sit 144 bytes apart (128 + 16) so that the
// eight float4 of one row's 128-byte global segment -- one quarter-warp -- land in eight different bank groups:
// full-line global requests AND conflict-free 16-byte stores.
constexpr int D_LBO = 144, D_SBO = (CK / 4) * D_LBO;       // 1152
constexpr int D_BYTES = (TR / 8) * D_SBO;                  // 18432 per hi / lo
constexpr int STAGE = 2 * D_BYTES;                         // 32 KB
constexpr int NSTAGE = 2;
constexpr int SS = H + 1;                                  // epilogue tile row stride (floats)
constexpr int MAXF = 32;
}  // namespace dxk

static inline size_t dx_tc_smem_bytes(int Fin) {
    return 2 * (size_t)dxk::W_BYTES + dxk::NSTAGE * (size_t)dxk::STAGE + sizeof(float) * (size_t)(Fin * dxk::H + dxk::H + dxk::H * Fin) + 96;
}

template <int FMAX>            // compile-time bound on the torso input width (register accumulators)
__global__ void __launch_bounds__(dxk::NT, 1)
lstm_dx_tc_kernel(const float* __restrict__ dpre, const float* __restrict__ u,
                  const float* __restrict__ W_ih, const float* __restrict__ Wt, const float* __restrict__ bt,
                  float* __restrict__ partials, long long P, long long off_Wt, long long off_bt,
                  float* __restrict__ du, long long rows, int Fin) {
    using namespace dxk;
    B200RL_DYN_SMEM(unsigned char, smem);
    unsigned char* W_hi = smem;
    unsigned char* W_lo = smem + W_BYTES;
    unsigned char* stage0 = smem + 2 * W_BYTES;
    float* WtS = reinterpret_cast<float*>(stage0 + NSTAGE * STAGE);     // WtS[q*64 + j]
    float* bts = WtS + Fin * H;
    float* WtR = bts + H;                                                // WtR[j*Fin + q]
    uint64_t* bars = reinterpret_cast<uint64_t*>(WtR + H * Fin);         // full[2] | empty[2] | tile_done | tmem_free
    uint64_t* full = bars, *empty = bars + NSTAGE, *tile_done = bars + 2 * NSTAGE, *tmem_free = bars + 2 * NSTAGE + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
    // the epilogue reuses the (then idle) stage memory: S[128][65] floats, then U[128][Fin]
    float* S = reinterpret_cast<float*>(stage0);
    float* Us = S + TR * SS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < NP / 32;

    // ---- one-time: W_ih -> TF32 hi/lo, K-major B operand: elem(k, p) = W_ih[(p&3)*64 + (p>>2)][k]
    // lanes cover one 128-byte core matrix (8 k x 4 p) per store -> conflict-free
    for (int i = tid; i < G * H; i += NT) {          // (any stride: every i is handled exactly once)
        const int p = ((i >> 5) & 63) * 4 + (i & 3), k = (i >> 11) * 8 + ((i >> 2) & 7);
        const float w = W_ih[(size_t)((p & 3) * 64 + (p >> 2)) * H + k];
        float hi, lo;
        tc::split_tf32(w, hi, lo);
        const int o = (k & 7) * 16 + (k >> 3) * W_SBO + (p & 3) * 4 + (p >> 2) * W_LBO;
        *reinterpret_cast<float*>(W_hi + o) = hi;
        *reinterpret_cast<float*>(W_lo + o) = lo;
    }
    for (int i = tid; i < Fin * H; i += NT) { int q = i / H, j = i % H; WtS[i] = Wt[j * Fin + q]; WtR[j * Fin + q] = Wt[j * Fin + q]; }
    if (tid < H) bts[tid] = bt[tid];
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], NP / 32); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(tile_done, 1); tc::mbar_init(tmem_free, NP / 32);
        tc::fence_mbar_init();
    }
    if (warp == NP / 32) tc::tmem_alloc(tmem_slot, 64);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = tc::make_idesc_tf32(TR, H, false, false);

    const long long ntiles = (rows + TR - 1) / TR;
    if (!producer) {
        if (lane == 0) {
            int g = 0, tdone = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tdone) {
                if (tdone > 0) { tc::mbar_wait(tmem_free, (uint32_t)((tdone - 1) & 1)); tc::fence_after_sync(); }
                for (int ch = 0; ch < NCH; ++ch, ++g) {
                    const int buf = g & 1, use = g >> 1;
                    unsigned char* D_hi = stage0 + (size_t)buf * STAGE, *D_lo = D_hi + D_BYTES;
                    tc::mbar_wait(&full[buf], (uint32_t)(use & 1));
                    tc::fence_after_sync();
#pragma unroll
                    for (int s = 0; s < CK / 8; ++s) {
                        const tc::SmemDesc ah = tc::make_desc(D_hi + s * 2 * D_LBO, D_LBO, D_SBO);
                        const tc::SmemDesc al = tc::make_desc(D_lo + s * 2 * D_LBO, D_LBO, D_SBO);
                        const int wo = (ch * (CK / 4) + s * 2) * W_LBO;
                        const tc::SmemDesc wh = tc::make_desc(W_hi + wo, W_LBO, W_SBO);
                        const tc::SmemDesc wl = tc::make_desc(W_lo + wo, W_LBO, W_SBO);
                        tc::mma_tf32(tmem, ah, wh, idesc, (ch > 0 || s > 0) ? 1u : 0u);
                        tc::mma_tf32(tmem, ah, wl, idesc, 1u);
                        tc::mma_tf32(tmem, al, wh, idesc, 1u);
                    }
                    tc::mma_commit(&empty[buf]);
                    if (ch == NCH - 1) tc::mma_commit(tile_done);
                }
            }
        }
    } else {
    // reduction ownership: column j = tid % 64, row quarter part = tid / 64
    const int rj = tid & 63, part = tid >> 6;
    float acc_bt = 0.f, acc_wt[FMAX];
#pragma unroll
    for (int q = 0; q < FMAX; ++q) acc_wt[q] = 0.f;

    // register-staged chunk loads: a ring of 4 slots, loads issued 3 chunks (~3 barriers) ahead of
    // their use so HBM latency overlaps the split/store/MMA work of the chunks in between
    constexpr int RING = 4, AHEAD = 3;
    static_assert(NCH % RING == 0, "ring slots must be compile-time constants inside the unrolled chunk loop");
    float4 dv[RING][4];
    auto issue_loads = [&](float4 (&dst)[4], long long tile, int ch) {
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const int r = warp * 16 + i4 * 4 + (lane >> 3), c4 = ch * (CK / 4) + (lane & 7);
            const long long row = tile * TR + r;
            dst[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tile < ntiles && row < rows) dst[i4] = *reinterpret_cast<const float4*>(dpre + (size_t)row * G + c4 * 4);
        }
    };
#pragma unroll
    for (int c = 0; c < AHEAD; ++c) issue_loads(dv[c], blockIdx.x, c);

    int g = 0;                                            // chunk counter across tiles (stage ring position)
    int tdone = 0;                                        // tiles completed by this CTA (tile-done parity)
    DXP_INIT
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tdone) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch, ++g) {
            const int buf = g & 1, use = g >> 1;
            unsigned char* D_hi = stage0 + (size_t)buf * STAGE, *D_lo = D_hi + D_BYTES;
            float4 cd[4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) cd[i4] = dv[ch % RING][i4];
            if (ch + AHEAD < NCH) issue_loads(dv[(ch + AHEAD) % RING], tile, ch + AHEAD);
            else issue_loads(dv[(ch + AHEAD) % RING], tile + gridDim.x, ch + AHEAD - NCH);
#ifdef B200RL_DXPROF
            if (cd[0].x == 1234.5f || cd[1].x == 1234.5f || cd[2].x == 1234.5f || cd[3].x == 1234.5f) acc_bt += 1.f;
            DXP(8);
#endif
            if (use >= 1) tc::mbar_wait(&empty[buf], (uint32_t)((use - 1) & 1));
            DXP(9);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const int r = warp * 16 + i4 * 4 + (lane >> 3), c4l = lane & 7;
                float4 hi, lo;
                tc::split_tf32(cd[i4].x, hi.x, lo.x); tc::split_tf32(cd[i4].y, hi.y, lo.y);
                tc::split_tf32(cd[i4].z, hi.z, lo.z); tc::split_tf32(cd[i4].w, hi.w, lo.w);
                const int o = (r & 7) * 16 + (r >> 3) * D_SBO + c4l * D_LBO;
                *reinterpret_cast<float4*>(D_hi + o) = hi;
                *reinterpret_cast<float4*>(D_lo + o) = lo;
            }
            tc::fence_async_smem();
            tc::mbar_arrive_warp(&full[buf]);
            DXP(10);
        }
        // ---- epilogue of this tile
        DXP(0);
        tc::mbar_wait(tile_done, (uint32_t)(tdone & 1));
        tc::fence_after_sync();
        DXP(1);
        {
            const int r = 32 * (warp & 3) + lane, c0 = (warp >> 2) * 32;         // row, first of 32 columns
            const long long row = tile * TR + r;
#ifdef B200RL_EMUL
            const int addr_lane = r;
#else
            const int addr_lane = 32 * (warp & 3);
#endif
            float dxv[32];
            tc::tmem_ld32(tc::tmem_addr(tmem, addr_lane, c0), dxv);
            const bool ok = row < rows;
            const float* ur = u + (size_t)(ok ? row : 0) * Fin;
            float uq[FMAX];                                   // this row's torso inputs, loaded once
#pragma unroll
            for (int q = 0; q < FMAX; ++q) uq[q] = (ok && q < Fin) ? __ldg(ur + q) : 0.f;
#ifdef B200RL_DXPROF
            if (uq[0] == 1234.5f) dxv[0] = 0.f;
            DXP(2);
#endif
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const int k = c0 + i;
                float a = bts[k];
#pragma unroll
                for (int q = 0; q < FMAX; ++q) if (q < Fin) a = fmaf(uq[q], WtS[q * H + k], a);
                S[r * SS + k] = (ok && a > 0.f) ? dxv[i] : 0.f;
            }
            if (warp < 4) {
#pragma unroll
                for (int q = 0; q < FMAX; ++q) if (q < Fin) Us[r * Fin + q] = uq[q];
            }
        }
        tc::fence_before_sync();
        tc::mbar_arrive_warp(tmem_free);                        // accumulator drained: the issuer may start the next tile
        DXP(3);
        tc::named_bar_sync(1, NP);
        DXP(4);
        {
            float sb = 0.f;
#pragma unroll 4
            for (int r = part * 32; r < part * 32 + 32; ++r) {
                const float s = S[r * SS + rj];
                sb += s;
#pragma unroll
                for (int q = 0; q < FMAX; ++q) if (q < Fin) acc_wt[q] = fmaf(s, Us[r * Fin + q], acc_wt[q]);
            }
            acc_bt += sb;
        }
        DXP(5);
        if (du != nullptr && tid < TR) {
            const long long row = tile * TR + tid;
            if (row < rows)
                for (int q = 0; q < Fin; ++q) {
                    float s = 0.f;
                    for (int j = 0; j < H; ++j) s = fmaf(S[tid * SS + j], WtR[j * Fin + q], s);
                    du[(size_t)row * Fin + q] = s;
                }
        }
        DXP(6);
        tc::named_bar_sync(1, NP);                         // S / Us (stage memory) free for the next tile
        DXP(7);
    }
    // ---- combine the four row quarters and write this CTA's slab entries
    float* red = reinterpret_cast<float*>(stage0);         // [4][64][Fin+1]
    const int RW = Fin + 1;
    red[(part * H + rj) * RW + Fin] = acc_bt;
#pragma unroll
    for (int q = 0; q < FMAX; ++q) if (q < Fin) red[(part * H + rj) * RW + q] = acc_wt[q];
    tc::named_bar_sync(1, NP);
    float* out = partials + (size_t)blockIdx.x * P;
    for (int i = tid; i < H * RW; i += NP) {
        const int j = i / RW, q = i % RW;
        const float s = red[(0 * H + j) * RW + q] + red[(1 * H + j) * RW + q] + red[(2 * H + j) * RW + q] + red[(3 * H + j) * RW + q];
        if (q == Fin) out[off_bt + j] = s; else out[off_Wt + (size_t)j * Fin + q] = s;
    }
    }   // producer
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NP / 32) tc::tmem_dealloc(tmem, 64);
}

// --------------------------------------------------------------------------- head gradients
// Bandwidth-bound over h (256 B per row): 1024 threads = 4 row groups x 256 so that each SM has
// 32 warps of loads in flight; the groups are combined through shared memory at the end.
// TILED (NH <= 4, the actor-critic heads): a thread owns (head c, 4 consecutive k, 16 of the chunk's 64 rows) and does 4 FMA
// per scalar + 16-byte shared-memory load; one FMA per two scalar loads made the reduction LDS-bound (ncu r1h: mio_throttle 3.2).
constexpr int HG_R = 64, HG_G = 4, HG_S = 68;       // rows per chunk, row groups, smem row stride (16-byte aligned rows)
template <bool TILED>
__global__ void __launch_bounds__(256 * HG_G)
head_grad_kernel(const float* __restrict__ h, const float* __restrict__ dhead, float* __restrict__ partials,
                 long long P, long long off_Whead, long long off_bhead, long long rows, int NH) {
    B200RL_DYN_SMEM(float, smem);
    const int grp = threadIdx.x >> 8, tid = threadIdx.x & 255;
    float* S = smem + grp * (HG_R * HG_S + HG_R * B200RL_MAX_HEADS);
    float* Dh = S + HG_R * HG_S;
    constexpr int E = TILED ? 1 : (B200RL_MAX_HEADS * 64 + 255) / 256;
    float acc[E], accb = 0.f;
    float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    const int items = NH * 16, slice = tid / items, item = tid - slice * items, tc_ = item >> 4, tk4 = item & 15;   // TILED roles
    const long long nch = (rows + HG_R - 1) / HG_R;
    const long long nloop = (nch + (long long)gridDim.x * HG_G - 1) / ((long long)gridDim.x * HG_G);
    // software pipeline: the HBM loads of chunk it+1 are in flight (registers) while chunk it is reduced from shared
    // memory -- without it every chunk paid a full DRAM round trip between two barriers (80 us for 140 MB)
    constexpr int HV = HG_R * 16 / 256, DV = (HG_R * B200RL_MAX_HEADS + 255) / 256;
    float4 hv[HV]; float dv[DV];
    auto fetch = [&](long long it) {
        const long long row0 = ((it * gridDim.x + blockIdx.x) * HG_G + grp) * HG_R;
#pragma unroll
        for (int e = 0; e < HV; ++e) {
            const int i = e * 256 + tid, r = i >> 4, c4 = i & 15;
            hv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < nloop && row0 + r < rows) hv[e] = *reinterpret_cast<const float4*>(h + (size_t)(row0 + r) * 64 + c4 * 4);
        }
#pragma unroll
        for (int e = 0; e < DV; ++e) {
            const int i = e * 256 + tid;
            dv[e] = (it < nloop && i < HG_R * NH && row0 + i / NH < rows) ? dhead[(size_t)row0 * NH + i] : 0.f;
        }
    };
    fetch(0);
    for (long long it = 0; it < nloop; ++it) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < HV; ++e) {
            const int i = e * 256 + tid, r = i >> 4, c4 = i & 15;
            st4(S + r * HG_S + c4 * 4, hv[e]);
        }
#pragma unroll
        for (int e = 0; e < DV; ++e) {
            const int i = e * 256 + tid;
            if (i < HG_R * NH) Dh[i] = dv[e];
        }
        __syncthreads();
        fetch(it + 1);
        if (TILED) {
            if (slice < 4) {
#pragma unroll 8
                for (int r = slice * 16; r < slice * 16 + 16; ++r) {
                    const float d = Dh[r * NH + tc_];
                    const float4 v = ld4(S + r * HG_S + tk4 * 4);
                    acc4.x = fmaf(d, v.x, acc4.x); acc4.y = fmaf(d, v.y, acc4.y); acc4.z = fmaf(d, v.z, acc4.z); acc4.w = fmaf(d, v.w, acc4.w);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int o = e * 256 + tid;
                if (o < NH * 64) {
                    const int c = o >> 6, k = o & 63;
                    float s = 0.f;
#pragma unroll 8
                    for (int r = 0; r < HG_R; ++r) s = fmaf(Dh[r * NH + c], S[r * HG_S + k], s);
                    acc[e] += s;
                }
            }
        }
        if (tid < NH) {
            float s = 0.f;
            for (int r = 0; r < HG_R; ++r) s += Dh[r * NH + tid];
            accb += s;
        }
    }
    // combine the row groups (and row slices): red[part][NH*64 + NH]
    __syncthreads();
    float* red = smem;
    const int RW = NH * 64 + NH, RS = (RW + 3) & ~3;        // row stride of `red`: 16-byte aligned for the float4 stores
    const int nparts_red = TILED ? HG_G * 4 : HG_G;
    if (TILED) {
        for (int i = threadIdx.x; i < nparts_red * RS; i += 256 * HG_G) red[i] = 0.f;     // (bias slots are written by one slice only)
        __syncthreads();
        if (slice < 4) st4(red + (grp * 4 + slice) * RS + tc_ * 64 + tk4 * 4, acc4);
        if (tid < NH) red[(grp * 4) * RS + NH * 64 + tid] = accb;
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int o = e * 256 + tid;
            if (o < NH * 64) red[grp * RS + o] = acc[e];
        }
        if (tid < NH) red[grp * RS + NH * 64 + tid] = accb;
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < RW; i += 256 * HG_G) {
        float s = 0.f;
        for (int gg = 0; gg < nparts_red; ++gg) s += red[gg * RS + i];
        if (i < NH * 64) out[off_Whead + i] = s; else out[off_bhead + (i - NH * 64)] = s;
    }
}

static int launch_head_grad(const float* h, const float* dhead, float* partials, int nparts, long long P,
                            long long off_Whead, long long off_bhead, long long rows, int NH, void* stream);

// --------------------------------------------------------------------------- torso gradients from a stored dx
// When the backward kernel already produced dx = dpre . W_ih (bwd_tc.cu fuses that GEMM into the recurrent one),
// what is left is a bandwidth-bound pass over dx (256 B per row): ReLU mask (x recomputed from u), dbt, dWt.
// 1024 threads: 16 threads x float4 per row, 64 rows per CTA pass.
template <int FMAX, int TG_T>                      // (8 torso inputs need more registers: half the threads)
__global__ void __launch_bounds__(TG_T)
torso_grad_kernel(const float* __restrict__ dxg, const float* __restrict__ u, const float* __restrict__ Wt,
                  const float* __restrict__ bt, float* __restrict__ partials, long long P, long long off_Wt, long long off_bt,
                  long long rows, int Fin) {
    B200RL_DYN_SMEM(float, smem);                     // red[32 warps][(Fin+1) * 64]
    const int tid = threadIdx.x, c4 = tid & 15, part = tid >> 4, lane = tid & 31, warp = tid >> 5;
    float4 w[FMAX];                                   // Wt[4 c4 + e][q], e = x,y,z,w
#pragma unroll
    for (int q = 0; q < FMAX; ++q)
        w[q] = q < Fin ? make_float4(Wt[(c4 * 4 + 0) * Fin + q], Wt[(c4 * 4 + 1) * Fin + q], Wt[(c4 * 4 + 2) * Fin + q], Wt[(c4 * 4 + 3) * Fin + q])
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b4 = *reinterpret_cast<const float4*>(bt + c4 * 4);
    float4 ab = make_float4(0.f, 0.f, 0.f, 0.f), aw[FMAX];
#pragma unroll
    for (int q = 0; q < FMAX; ++q) aw[q] = ab;
#pragma unroll 2
    for (long long r = (long long)blockIdx.x * (TG_T / 16) + part; r < rows; r += (long long)gridDim.x * (TG_T / 16)) {
        const float4 d = *reinterpret_cast<const float4*>(dxg + (size_t)r * 64 + c4 * 4);
        float uq[FMAX];
#pragma unroll
        for (int q = 0; q < FMAX; ++q) uq[q] = q < Fin ? __ldg(u + (size_t)r * Fin + q) : 0.f;
        float4 x = b4;
#pragma unroll
        for (int q = 0; q < FMAX; ++q) { x.x = fmaf(uq[q], w[q].x, x.x); x.y = fmaf(uq[q], w[q].y, x.y); x.z = fmaf(uq[q], w[q].z, x.z); x.w = fmaf(uq[q], w[q].w, x.w); }
        const float4 sv = make_float4(x.x > 0.f ? d.x : 0.f, x.y > 0.f ? d.y : 0.f, x.z > 0.f ? d.z : 0.f, x.w > 0.f ? d.w : 0.f);
        ab.x += sv.x; ab.y += sv.y; ab.z += sv.z; ab.w += sv.w;
#pragma unroll
        for (int q = 0; q < FMAX; ++q) {
            aw[q].x = fmaf(sv.x, uq[q], aw[q].x); aw[q].y = fmaf(sv.y, uq[q], aw[q].y);
            aw[q].z = fmaf(sv.z, uq[q], aw[q].z); aw[q].w = fmaf(sv.w, uq[q], aw[q].w);
        }
    }
    // the two row lanes of a warp, then the 32 warps through shared memory
    auto fold = [&](float4& v) {
        v.x += __shfl_xor_sync(kFull, v.x, 16); v.y += __shfl_xor_sync(kFull, v.y, 16);
        v.z += __shfl_xor_sync(kFull, v.z, 16); v.w += __shfl_xor_sync(kFull, v.w, 16);
    };
    fold(ab);
#pragma unroll
    for (int q = 0; q < FMAX; ++q) fold(aw[q]);
    const int RW = (Fin + 1) * 64;
    if (lane < 16) {
        float* d = smem + warp * RW;
        *reinterpret_cast<float4*>(d + Fin * 64 + c4 * 4) = ab;
#pragma unroll
        for (int q = 0; q < FMAX; ++q) if (q < Fin) *reinterpret_cast<float4*>(d + q * 64 + c4 * 4) = aw[q];
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * P;
    for (int i = tid; i < RW; i += TG_T) {
        float sum = 0.f;
        for (int wv = 0; wv < TG_T / 32; ++wv) sum += smem[wv * RW + i];
        const int q = i >> 6, jj = i & 63;
        if (q == Fin) out[off_bt + jj] = sum; else out[off_Wt + (size_t)jj * Fin + q] = sum;
    }
}

template <typename K>
static int dx_set_smem(K kern, size_t smem) {
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("lstm_wgrad(dx tc): cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
#endif
    return 0;
}

static int launch_head_grad(const float* h, const float* dhead, float* partials, int nparts, long long P,
                            long long off_Whead, long long off_bhead, long long rows, int NH, void* stream) {
    const size_t hsm = sizeof(float) * HG_G * (HG_R * HG_S + HG_R * B200RL_MAX_HEADS);
    if (NH <= 4) {
        if (dx_set_smem(head_grad_kernel<true>, hsm)) return 1;
        B200RL_LAUNCH(head_grad_kernel<true>, nparts, 256 * HG_G, hsm, stream, h, dhead, partials, P, off_Whead, off_bhead, rows, NH);
    } else {
        if (dx_set_smem(head_grad_kernel<false>, hsm)) return 1;
        B200RL_LAUNCH(head_grad_kernel<false>, nparts, 256 * HG_G, hsm, stream, h, dhead, partials, P, off_Whead, off_bhead, rows, NH);
    }
    return check_launch("lstm_wgrad(head grad)");
}

int launch_dx_tc(const float* dpre, const float* u, const float* W_ih, const float* Wt, const float* bt,
                 const float* h, const float* dhead, float* partials, int nparts, long long P,
                 long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                 float* du, int B, int T, int Fin, int NH, void* stream) {
    B200RL_REQUIRE(Fin >= 1 && Fin <= dxk::MAXF, "lstm_wgrad(dx tc): Fin=%d out of range", Fin);
    const long long rows = (long long)B * T;
    const size_t smem = dx_tc_smem_bytes(Fin);
#define B200RL_DX_LAUNCH(FM)                                                                                          \
    do {                                                                                                              \
        if (dx_set_smem(lstm_dx_tc_kernel<FM>, smem)) return 1;                                                       \
        B200RL_LAUNCH(lstm_dx_tc_kernel<FM>, nparts, dxk::NT, smem, stream, dpre, u, W_ih, Wt, bt, partials, P, off_Wt, off_bt, du, rows, Fin); \
    } while (0)
    if (Fin <= 4) B200RL_DX_LAUNCH(4);
    else if (Fin <= 8) B200RL_DX_LAUNCH(8);
    else if (Fin <= 16) B200RL_DX_LAUNCH(16);
    else B200RL_DX_LAUNCH(32);
#undef B200RL_DX_LAUNCH
    if (check_launch("lstm_wgrad(dx tc)")) return 1;
    return launch_head_grad(h, dhead, partials, nparts, P, off_Whead, off_bhead, rows, NH, stream);
}

int launch_torso_from_dx(const float* dx, const float* u, const float* Wt, const float* bt,
                         const float* h, const float* dhead, float* partials, int nparts, long long P,
                         long long off_Wt, long long off_bt, long long off_Whead, long long off_bhead,
                         int B, int T, int Fin, int NH, void* stream) {
    B200RL_REQUIRE(Fin >= 1 && Fin <= 8, "lstm_wgrad(torso from dx): Fin=%d out of range [1,8]", Fin);
    const long long rows = (long long)B * T;
    if (Fin <= 4) {
        const size_t tsm = sizeof(float) * (size_t)(1024 / 32) * (Fin + 1) * 64;
        auto kern = torso_grad_kernel<4, 1024>;
        if (dx_set_smem(kern, tsm)) return 1;
        B200RL_LAUNCH(kern, nparts, 1024, tsm, stream, dx, u, Wt, bt, partials, P, off_Wt, off_bt, rows, Fin);
    } else {
        const size_t tsm = sizeof(float) * (size_t)(512 / 32) * (Fin + 1) * 64;
        auto kern = torso_grad_kernel<8, 512>;
        if (dx_set_smem(kern, tsm)) return 1;
        B200RL_LAUNCH(kern, nparts, 512, tsm, stream, dx, u, Wt, bt, partials, P, off_Wt, off_bt, rows, Fin);
    }
    if (check_launch("lstm_wgrad(torso from dx)")) return 1;
    if (h == nullptr) return 0;                          // (the head gradients were accumulated by the backward kernel)
    return launch_head_grad(h, dhead, partials, nparts, P, off_Whead, off_bhead, rows, NH, stream);
}

}  // namespace b200rl
