// Torso + LSTMCell-over-T + linear heads: forward, BPTT and weight gradients (sm_100a).
//
// Replaces the reference's `MlpLSTM*.forward` python loop over T
// (networks/models.py:63-100,135-322: Linear+ReLU torso, nn.LSTMCell unrolled over the
// window, Linear heads) and the autograd graph replay of `loss.backward()`
// (agents/learner_module/ppo/learning.py:92) with four kernels:
//
//   lstm_fwd    persistent: a CTA owns 32 sequences for all T steps; [W_ih|W_hh]
//               (128x256 fp32, gate-interleaved) stays in shared memory for the whole
//               launch, c stays in registers, [x_t|h_{t-1}] ping-pongs in shared
//               memory; one __syncthreads per timestep.  Saves gates/c/h for BPTT.
//   head_fwd    head_out = h . W_head^T + b (tiny N; bandwidth-bound over h).
//   lstm_bwd    persistent reverse-time: LSTM cell backward + dh_{t-1} = dpre . W_hh,
//               overwrites the saved gates with d(pre-activation) in place.
//   lstm_wgrad  parallel over (b,t): dW_ih, dW_hh, db, dW_head, db_head, torso grads as
//               per-CTA partial slabs (deterministic; reduced by reduce_partials).
//
// All arithmetic is fp32 FFMA: the losses this path feeds are near-zero means of signed
// terms and single-pass TF32 misses the rtol 1e-4 parity bar.
//
// Saved-activation layout (private): gates (B,T,64,4) = per hidden unit j the four
// post-activation gates (i,f,g,o) as one float4; c, h (B,T,64).
#include "common.cuh"
#include "launchers.cuh"
#include "tc05.cuh"

namespace b200rl {

constexpr int H = 64;
constexpr int G = 4 * H;      // 256 gate columns
constexpr int KC = 2 * H;     // 128 = [x | h] concatenated K
constexpr int TM = 32;        // sequences per CTA tile
constexpr int NT = 256;       // threads per CTA
constexpr int DS = 36;        // row stride of r-minor (transposed) tiles: 16B aligned, bank-skewed
constexpr int HM = TM / 2;    // sequences per independent half-CTA (backward kernel)
constexpr int MAXF = 32;      // max torso input width

// thread -> (4 rows, 2 hidden units).  A shared-memory LDS.128 is served one quarter-warp
// (8 lanes) per wavefront and identical quarter-warps merge, so the operand that is read
// with 128-bit loads in the hot loop should vary with lane&7 and repeat across quarters.
//   MapR: lane&7 -> row group  (r-minor tiles: backward D_T, wgrad)   [rows vary in-quarter]
//   MapJ: lane&7 -> hidden unit (forward: the two W float4 loads per k dominate)
struct MapR {
    int r0, j0, j1;
    __device__ MapR() {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        r0 = (lane & 7) * 4;
        j0 = warp * 4 + (lane >> 3);
        j1 = j0 + 32;
    }
};
struct MapJ {
    int r0, j0, j1;
    __device__ MapJ() {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        r0 = ((warp & 1) * 4 + (lane >> 3)) * 4;
        j0 = (warp >> 1) * 8 + (lane & 7);
        j1 = j0 + 32;
    }
};


// =============================================================================== forward
// shared (floats): Wp[128][256] | AT[2][128][32] | bsum[256] | bt[64] | WtS[Fin][64] | U[2][32][Fin]
static inline size_t fwd_smem_bytes(int Fin) {
    return sizeof(float) * (size_t)(KC * G + 2 * KC * TM + G + H + Fin * H + 2 * TM * Fin);
}

__global__ void __launch_bounds__(NT, 1)
lstm_fwd_kernel(const float* __restrict__ u, const float* __restrict__ hx0, const float* __restrict__ cx0,
                const float* __restrict__ Wt, const float* __restrict__ bt,
                const float* __restrict__ W_ih, const float* __restrict__ W_hh,
                const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                float* __restrict__ gates, float* __restrict__ c_out, float* __restrict__ h_out,
                int B, int T, int Fin) {
    B200RL_DYN_SMEM(float, smem);
    float* Wp = smem;                       // Wp[k*256 + j*4 + g]
    float* AT = Wp + KC * G;                // AT[buf][k*32 + r]
    float* bsum = AT + 2 * KC * TM;         // bsum[j*4 + g]
    float* bts = bsum + G;
    float* WtS = bts + H;                   // WtS[q*64 + j]
    float* U = WtS + Fin * H;               // U[buf][r*Fin + q]
    const int tid = threadIdx.x;
    const MapJ m;

    // ---- one-time: stage the weights (PyTorch layout (4H, H), gate order i,f,g,o)
    {
        const int n = tid, g = n >> 6, j = n & 63, pc = j * 4 + g;
        for (int k4 = 0; k4 < H; k4 += 4) {
            float4 a = ld4(W_ih + (size_t)n * H + k4);
            float4 b = ld4(W_hh + (size_t)n * H + k4);
            Wp[(k4 + 0) * G + pc] = a.x; Wp[(k4 + 1) * G + pc] = a.y;
            Wp[(k4 + 2) * G + pc] = a.z; Wp[(k4 + 3) * G + pc] = a.w;
            Wp[(H + k4 + 0) * G + pc] = b.x; Wp[(H + k4 + 1) * G + pc] = b.y;
            Wp[(H + k4 + 2) * G + pc] = b.z; Wp[(H + k4 + 3) * G + pc] = b.w;
        }
        bsum[pc] = b_ih[n] + b_hh[n];
        if (tid < H) bts[tid] = bt[tid];
        for (int i = tid; i < Fin * H; i += NT) { int q = i / H, jj = i % H; WtS[i] = Wt[jj * Fin + q]; }
    }

    const int ntiles = (B + TM - 1) / TM;
    const int nU = TM * Fin;                 // floats per U tile
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b0 = tile * TM;
        __syncthreads();                     // weights staged / previous tile finished with AT, U
        // rows of this thread (clamped for the ragged last tile; writes are predicated)
        int brow[4]; bool bok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { int b = b0 + m.r0 + i; bok[i] = b < B; brow[i] = bok[i] ? b : B - 1; }

        float cst[4][2];
        {
            float hv0[4], hv1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cst[i][0] = cx0[(size_t)brow[i] * H + m.j0]; cst[i][1] = cx0[(size_t)brow[i] * H + m.j1];
                hv0[i] = hx0[(size_t)brow[i] * H + m.j0];    hv1[i] = hx0[(size_t)brow[i] * H + m.j1];
            }
            st4(AT + (H + m.j0) * TM + m.r0, make_float4(hv0[0], hv0[1], hv0[2], hv0[3]));
            st4(AT + (H + m.j1) * TM + m.r0, make_float4(hv1[0], hv1[1], hv1[2], hv1[3]));
        }
        // U[0] <- u_0, U[1] <- u_1
        for (int i = tid; i < nU; i += NT) {
            int r = i / Fin, q = i % Fin; int b = b0 + r < B ? b0 + r : B - 1;
            U[i] = u[((size_t)b * T + 0) * Fin + q];
            if (T > 1) U[nU + i] = u[((size_t)b * T + 1) * Fin + q];
        }
        __syncthreads();
        // x_0 -> AT[0]
        {
            float x0[4], x1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a0 = bts[m.j0], a1 = bts[m.j1];
                for (int q = 0; q < Fin; ++q) {
                    float uv = U[(m.r0 + i) * Fin + q];
                    a0 = fmaf(uv, WtS[q * H + m.j0], a0); a1 = fmaf(uv, WtS[q * H + m.j1], a1);
                }
                x0[i] = fmaxf(a0, 0.f); x1[i] = fmaxf(a1, 0.f);
            }
            st4(AT + m.j0 * TM + m.r0, make_float4(x0[0], x0[1], x0[2], x0[3]));
            st4(AT + m.j1 * TM + m.r0, make_float4(x1[0], x1[1], x1[2], x1[3]));
        }
        __syncthreads();

        for (int t = 0; t < T; ++t) {
            const float* Ac = AT + (t & 1) * KC * TM;
            float* An = AT + ((t + 1) & 1) * KC * TM;
            // (a) prefetch u_{t+2} into registers (lands in U[t&1] at the end of the step)
            float upre[(TM * MAXF + NT - 1) / NT];
            const bool do_pre = t + 2 < T;
            if (do_pre) {
#pragma unroll
                for (int e = 0; e < (TM * MAXF + NT - 1) / NT; ++e) {
                    int i = e * NT + tid;
                    if (i < nU) {
                        int r = i / Fin, q = i % Fin; int b = b0 + r < B ? b0 + r : B - 1;
                        upre[e] = __ldg(u + ((size_t)b * T + (t + 2)) * Fin + q);
                    }
                }
            }
            // (b) gates = bias + [x_t | h_{t-1}] . Wp       (32 x 256 x 128 per CTA), packed FFMA2:
            //     accumulator pairs run along the gate index, the activation is broadcast into both halves
            f2 acc2[4][4];
            {
                f2 b00, b01, b10, b11;
                f2_ld2(bsum + m.j0 * 4, b00, b01); f2_ld2(bsum + m.j1 * 4, b10, b11);
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc2[i][0] = b00; acc2[i][1] = b01; acc2[i][2] = b10; acc2[i][3] = b11; }
            }
#pragma unroll 4
            for (int k = 0; k < KC; ++k) {
                const float4 a = ld4(Ac + k * TM + m.r0);
                f2 w00, w01, w10, w11;
                f2_ld2(Wp + k * G + m.j0 * 4, w00, w01);
                f2_ld2(Wp + k * G + m.j1 * 4, w10, w11);
                const f2 av[4] = {f2_make(a.x, a.x), f2_make(a.y, a.y), f2_make(a.z, a.z), f2_make(a.w, a.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    f2_fma(acc2[i][0], av[i], w00); f2_fma(acc2[i][1], av[i], w01);
                    f2_fma(acc2[i][2], av[i], w10); f2_fma(acc2[i][3], av[i], w11);
                }
            }
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) { acc[i][2 * c] = f2_lo(acc2[i][c]); acc[i][2 * c + 1] = f2_hi(acc2[i][c]); }
            // (c) cell update, save activations, publish h_t for the next step
            float hn[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    float ig = sigm(acc[i][jj * 4 + 0]);
                    float fg = sigm(acc[i][jj * 4 + 1]);
                    float gg = tanh_(acc[i][jj * 4 + 2]);
                    float og = sigm(acc[i][jj * 4 + 3]);
                    float cn = fmaf(fg, cst[i][jj], ig * gg);
                    float hh = og * tanh_(cn);
                    cst[i][jj] = cn; hn[i][jj] = hh;
                    if (bok[i]) {
                        const int j = jj ? m.j1 : m.j0;
                        const size_t o = ((size_t)brow[i] * T + t) * H + j;
                        st4(gates + o * 4, make_float4(ig, fg, gg, og));
                        c_out[o] = cn; h_out[o] = hh;
                    }
                }
            }
            st4(An + (H + m.j0) * TM + m.r0, make_float4(hn[0][0], hn[1][0], hn[2][0], hn[3][0]));
            st4(An + (H + m.j1) * TM + m.r0, make_float4(hn[0][1], hn[1][1], hn[2][1], hn[3][1]));
            // (d) x_{t+1} from U[(t+1)&1]
            if (t + 1 < T) {
                const float* Un = U + ((t + 1) & 1) * nU;
                float x0[4], x1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a0 = bts[m.j0], a1 = bts[m.j1];
                    for (int q = 0; q < Fin; ++q) {
                        float uv = Un[(m.r0 + i) * Fin + q];
                        a0 = fmaf(uv, WtS[q * H + m.j0], a0); a1 = fmaf(uv, WtS[q * H + m.j1], a1);
                    }
                    x0[i] = fmaxf(a0, 0.f); x1[i] = fmaxf(a1, 0.f);
                }
                st4(An + m.j0 * TM + m.r0, make_float4(x0[0], x0[1], x0[2], x0[3]));
                st4(An + m.j1 * TM + m.r0, make_float4(x1[0], x1[1], x1[2], x1[3]));
            }
            if (do_pre) {
                float* Ud = U + (t & 1) * nU;
#pragma unroll
                for (int e = 0; e < (TM * MAXF + NT - 1) / NT; ++e) {
                    int i = e * NT + tid;
                    if (i < nU) Ud[i] = upre[e];
                }
            }
            __syncthreads();
        }
    }
}

// ============================================================================== head_fwd
// head_out[row][c] = b[c] + sum_k h[row][k] W[c][k];   128 rows per CTA pass, staged through smem; the next pass's
// rows are already in flight (registers) while this one is reduced.
// TILED (NH <= 4, the actor-critic heads): four threads share a row, each holds its 16 k-columns of W_head in registers
// (chunks kp, kp+4, kp+8, kp+12 of the row's sixteen float4) and reads h with conflict-free 16-byte loads (row stride 80
// floats: a quarter-warp = 2 rows x 4 chunks covers eight bank groups); two shuffles finish the dot products.  The scalar
// form needed two shared-memory loads per FMA and was LDS-bound (ncu r1h: short_scoreboard 8.1).
constexpr int HR = 128, HFS = 80;
template <bool TILED>
__global__ void __launch_bounds__(HR)
head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ W, const float* __restrict__ b,
                float* __restrict__ out, long long rows, int NH) {
    B200RL_DYN_SMEM(float, smem);
    float* Ht = smem;                       // [HR][HFS]
    float* Ws = Ht + HR * HFS;              // [32][H]
    float* bs = Ws + 32 * H;                // [32]
    const int tid = threadIdx.x;
    for (int i = tid; i < NH * H; i += HR) Ws[i] = W[i];
    if (tid < NH) bs[tid] = b[tid];
    const long long nblk = (rows + HR - 1) / HR;
    constexpr int NV = HR * H / 4 / HR;                   // float4 per thread per pass (16)
    float4 pv[NV];
    auto fetch = [&](long long blk) {
        const long long r0 = blk * HR;
#pragma unroll
        for (int it = 0; it < NV; ++it) {
            const int idx = it * HR + tid, r = idx >> 4, c4 = idx & 15;
            pv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (blk < nblk && r0 + r < rows) pv[it] = ld4(h + (size_t)(r0 + r) * H + c4 * 4);
        }
    };
    fetch(blockIdx.x);
    __syncthreads();
    const int kp = tid & 3, rsub = tid >> 2;              // TILED roles: k-chunk set, row inside a 32-row sub-pass
    float4 wreg[TILED ? 4 : 1][4];                        // W[c][chunk kp + 4 e]
    float breg[TILED ? 4 : 1];
    if (TILED) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            breg[c] = c < NH ? bs[c] : 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) wreg[c][e] = c < NH ? ld4(Ws + c * H + (e * 4 + kp) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long r0 = blk * HR;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < NV; ++it) {
            const int idx = it * HR + tid, r = idx >> 4, c4 = idx & 15;
            st4(Ht + r * HFS + c4 * 4, pv[it]);
        }
        __syncthreads();
        fetch(blk + gridDim.x);
        if (TILED) {
#pragma unroll
            for (int sub = 0; sub < HR / 32; ++sub) {
                const int r = sub * 32 + rsub;
                float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 v = ld4(Ht + r * HFS + (e * 4 + kp) * 4);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        a[c] = fmaf(v.x, wreg[c][e].x, fmaf(v.y, wreg[c][e].y, fmaf(v.z, wreg[c][e].z, fmaf(v.w, wreg[c][e].w, a[c]))));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    a[c] += __shfl_xor_sync(kFull, a[c], 1);
                    a[c] += __shfl_xor_sync(kFull, a[c], 2);
                }
                if (kp < NH && r0 + r < rows) {          // lane kp of the four writes head kp
                    const float v = kp == 0 ? a[0] + breg[0] : kp == 1 ? a[1] + breg[1] : kp == 2 ? a[2] + breg[2] : a[3] + breg[3];
                    out[(size_t)(r0 + r) * NH + kp] = v;
                }
            }
        } else if (r0 + tid < rows) {
            const float* hr = Ht + tid * HFS;
            for (int c = 0; c < NH; ++c) {
                float a = bs[c];
#pragma unroll 16
                for (int k = 0; k < H; ++k) a = fmaf(hr[k], Ws[c * H + k], a);
                out[(size_t)(r0 + tid) * NH + c] = a;
            }
        }
    }
}

// ======================================================================= backward (serial)
// The CTA is two independent halves of 128 threads x 16 sequences (as the recurrent forward): the halves
// share the read-only W_hh / W_head tiles and synchronise only with themselves (named barriers), so one
// half's cell phase (MUFU, global loads/stores) overlaps the other's dh GEMM (FFMA2 + LDS).
// shared (floats): WhT[256][64] | DT[half][2][256*16] | WhS[NH][64] | dHd[half][2][16][NH] | Rp[half][4][16*64]
static inline size_t bwd_smem_bytes(int NH) {
    return sizeof(float) * (size_t)(G * H + 2 * 2 * G * HM + NH * H + 2 * 2 * HM * NH + 4 + 2 * 4 * HM * H);
}
// D tile address of gate column p (16 rows each): neighbouring units j, j+1 (p differs by 4) land in
// different 64-byte halves of a 128-byte line, so a quarter-warp's cell-phase stores are conflict-free
__device__ __forceinline__ int bwd_daddr(int p) { return (p >> 1) * 32 + (((p & 1) ^ ((p >> 2) & 1)) * HM); }

__global__ void __launch_bounds__(NT, 1)
lstm_bwd_kernel(const float* __restrict__ dhead, const float* __restrict__ W_head,
                const float* __restrict__ W_hh, float* __restrict__ gates,
                const float* __restrict__ c_sav, const float* __restrict__ cx0,
                int B, int T, int NH) {
    B200RL_DYN_SMEM(float, smem);
    const int tid = threadIdx.x, hf = tid >> 7, ht = tid & 127, lane = tid & 31, hw = ht >> 5;
    const int nD = HM * NH;
    float* WhT = smem;                                   // WhT[p*64 + k] = W_hh[n(p)][k],  p = j*4+g, n = g*64+j
    float* DT = WhT + G * H + hf * 2 * G * HM;           // DT[buf][daddr(p) + r]
    float* WhS = WhT + G * H + 2 * 2 * G * HM;           // WhS[c*64 + k]
    float* dHd = WhS + NH * H + hf * 2 * nD;             // dHd[buf][r*NH + c]
    float* Rp = WhS + NH * H + 2 * 2 * nD;
    Rp += ((4 - ((NH * H + 2 * 2 * nD) & 3)) & 3) + hf * 4 * HM * H;   // Rp[slice][row*64 + k']: split-K partial sums (16-byte aligned)
    // cell ownership: 4 rows x 2 units
    const int r0 = (lane & 3) * 4, j0 = hw * 8 + (lane >> 2), j1 = j0 + 32;
    // dh GEMM ownership: K (the 256 gate columns) is split over the half's 4 warps; within a slice a thread owns
    // 4 rows x 8 k: 12 shared-memory wavefronts per 32 FMA (the kernel is wavefront-bound)
    const int gs = hw, grg = lane & 3, kq = lane >> 2, gk0 = kq * 8;
    {
        const int n = tid, g = n >> 6, j = n & 63, p = j * 4 + g;
        for (int k4 = 0; k4 < H; k4 += 4) st4(WhT + p * H + k4, ld4(W_hh + (size_t)n * H + k4));
        for (int i = tid; i < NH * H; i += NT) WhS[i] = W_head[i];
    }
    __syncthreads();
    const int ntiles = (B + TM - 1) / TM;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b0 = tile * TM + hf * HM;
        tc::named_bar_sync(1 + hf, 128);
        int brow[4]; bool bok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { int b = b0 + r0 + i; bok[i] = b < B; brow[i] = bok[i] ? b : B - 1; }
        // software pipeline (global latency must never sit on the per-step critical path):
        // registers hold gates_t, c_t, c_{t-1} for the step being processed; the loads of
        // gates_{t-1} and c_{t-2} are issued at the top of step t and first used one full step
        // later; the dhead tile of step t-2 is loaded at step t and parked in shared memory
        // at the end of step t-1.
        constexpr int NPF = (HM * B200RL_MAX_HEADS + 127) / 128;
        auto load_dhead = [&](int t, float (&dst)[NPF]) {
#pragma unroll
            for (int e = 0; e < NPF; ++e) {
                int i = e * 128 + ht;
                if (i < nD) {
                    int r = i / NH, c = i % NH; int b = b0 + r < B ? b0 + r : B - 1;
                    dst[e] = __ldg(dhead + ((size_t)b * T + t) * NH + c);
                }
            }
        };
        auto park_dhead = [&](int t, const float (&src)[NPF]) {
            float* Hd = dHd + (t & 1) * nD;
#pragma unroll
            for (int e = 0; e < NPF; ++e) {
                int i = e * 128 + ht;
                if (i < nD) Hd[i] = src[e];
            }
        };
        float dh_next[NPF];                       // dhead tile of step t-1 while step t runs
        {
            float tmp[NPF];
            load_dhead(T - 1, tmp);
            park_dhead(T - 1, tmp);
            if (T > 1) load_dhead(T - 2, dh_next);
        }
        float dc[4][2], dhr[4][2], ccur[4][2], cprev[4][2];
        float4 gcur[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = jj ? j1 : j0;
                const size_t o = ((size_t)brow[i] * T + (T - 1)) * H + j;
                dc[i][jj] = 0.f; dhr[i][jj] = 0.f;
                ccur[i][jj] = c_sav[o];
                cprev[i][jj] = T > 1 ? c_sav[o - H] : cx0[(size_t)brow[i] * H + j];
                gcur[i][jj] = ld4(gates + o * 4);
            }
        tc::named_bar_sync(1 + hf, 128);

        for (int t = T - 1; t >= 0; --t) {
            float* Dc = DT + (t & 1) * G * HM;
            // (a) issue loads whose first use is one full step away
            float4 gnx[4][2]; float cpp[4][2];
            float dh_new[NPF];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = jj ? j1 : j0;
                    gnx[i][jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                    cpp[i][jj] = 0.f;
                    if (t > 0) gnx[i][jj] = ld4(gates + (((size_t)brow[i] * T + (t - 1)) * H + j) * 4);
                    if (t > 1) cpp[i][jj] = c_sav[((size_t)brow[i] * T + (t - 2)) * H + j];
                    else if (t == 1) cpp[i][jj] = cx0[(size_t)brow[i] * H + j];
                }
            if (t > 1) load_dhead(t - 2, dh_new);
            // (b) dh = dh_rec + dhead_t . W_head ; LSTM cell backward
            const float* dH = dHd + (t & 1) * nD;
            float4 dp[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = jj ? j1 : j0;
                    float dh = dhr[i][jj];
                    for (int c = 0; c < NH; ++c) dh = fmaf(dH[(r0 + i) * NH + c], WhS[c * H + j], dh);
                    const float4 gt = gcur[i][jj];          // i, f, g, o (post-activation)
                    const float tc = tanh_(ccur[i][jj]);
                    const float d_o = dh * tc;
                    const float dct = fmaf(dh * gt.w, 1.f - tc * tc, dc[i][jj]);
                    const float d_i = dct * gt.z, d_g = dct * gt.x, d_f = dct * cprev[i][jj];
                    dc[i][jj] = dct * gt.y;
                    dp[i][jj] = make_float4(d_i * gt.x * (1.f - gt.x), d_f * gt.y * (1.f - gt.y),
                                            d_g * (1.f - gt.z * gt.z), d_o * gt.w * (1.f - gt.w));
                    if (bok[i]) st4(gates + (((size_t)brow[i] * T + t) * H + j) * 4, dp[i][jj]);
                }
            }
            // r-minor tile for the dh GEMM: one float4 (4 rows) per gate column
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int p0 = (jj ? j1 : j0) * 4;
                st4(Dc + bwd_daddr(p0 + 0) + r0, make_float4(dp[0][jj].x, dp[1][jj].x, dp[2][jj].x, dp[3][jj].x));
                st4(Dc + bwd_daddr(p0 + 1) + r0, make_float4(dp[0][jj].y, dp[1][jj].y, dp[2][jj].y, dp[3][jj].y));
                st4(Dc + bwd_daddr(p0 + 2) + r0, make_float4(dp[0][jj].z, dp[1][jj].z, dp[2][jj].z, dp[3][jj].z));
                st4(Dc + bwd_daddr(p0 + 3) + r0, make_float4(dp[0][jj].w, dp[1][jj].w, dp[2][jj].w, dp[3][jj].w));
            }
            // (d) park the dhead tile of step t-1 (loaded a step ago); visible after the barrier
            if (t > 0) park_dhead(t - 1, dh_next);
            tc::named_bar_sync(1 + hf, 128);
            // (c) dh_rec[r][k] = sum_p dpre[r][p] W_hh[n(p)][k]        (16 x 64 x 256 per half)
            if (t > 0) {
                f2 acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = f2_make(0.f, 0.f);
#pragma unroll 4
                for (int pp = 0; pp < G / 4; ++pp) {
                    const int p = gs * (G / 4) + pp;
                    const float4 d = ld4(Dc + bwd_daddr(p) + grg * 4);
                    f2 w0, w1, w2, w3;
                    f2_ld2(WhT + p * H + gk0, w0, w1);
                    f2_ld2(WhT + p * H + gk0 + 4, w2, w3);
                    const f2 dv[4] = {f2_make(d.x, d.x), f2_make(d.y, d.y), f2_make(d.z, d.z), f2_make(d.w, d.w)};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        f2_fma(acc[i][0], dv[i], w0); f2_fma(acc[i][1], dv[i], w1);
                        f2_fma(acc[i][2], dv[i], w2); f2_fma(acc[i][3], dv[i], w3);
                    }
                }
                // exchange through Rp, 16-byte granules XOR-swizzled by the row group and written in swapped order by
                // odd k-octets: both the stores here and the scalar reads below are bank-conflict-free (unswizzled,
                // this exchange was 41 % of all shared-memory wavefronts, ncu r1g)
                float* Rs = Rp + gs * HM * H;
                const bool odd = kq & 1;
                const int oa = (gk0 + (odd ? 4 : 0)) ^ (grg << 3), ob = (gk0 + (odd ? 0 : 4)) ^ (grg << 3);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float* dst = Rs + (grg * 4 + i) * H;
                    const float4 lo4 = make_float4(f2_lo(acc[i][0]), f2_hi(acc[i][0]), f2_lo(acc[i][1]), f2_hi(acc[i][1]));
                    const float4 hi4 = make_float4(f2_lo(acc[i][2]), f2_hi(acc[i][2]), f2_lo(acc[i][3]), f2_hi(acc[i][3]));
                    st4(dst + oa, odd ? hi4 : lo4);
                    st4(dst + ob, odd ? lo4 : hi4);
                }
                tc::named_bar_sync(1 + hf, 128);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) {
                        s0 += Rp[sl * HM * H + (r0 + i) * H + (j0 ^ (r0 << 1))];       // r0 << 1 = row group << 3
                        s1 += Rp[sl * HM * H + (r0 + i) * H + (j1 ^ (r0 << 1))];
                    }
                    dhr[i][0] = s0; dhr[i][1] = s1;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) { ccur[i][jj] = cprev[i][jj]; cprev[i][jj] = cpp[i][jj]; gcur[i][jj] = gnx[i][jj]; }
#pragma unroll
            for (int e = 0; e < NPF; ++e) dh_next[e] = dh_new[e];
        }
    }
}

// =================================================================== weight gradients
// shared (floats): DtT[256][DS] | AtT[128][DS] | WiT[256][64] | HtT[64][DS] | DxT[64][DS]
//                  | dHd[32][NH] | U[32][Fin] | WtS[Fin][64] | bts[64] | Wt_row[64][Fin]
static inline size_t wgrad_smem_bytes(int Fin, int NH) {
    return sizeof(float) * (size_t)(G * DS + KC * DS + G * H + H * DS + H * DS + TM * NH + TM * Fin + Fin * H + H + H * Fin);
}
constexpr int WG_E = 8;   // per-thread slots for the small (head / torso) gradient accumulators

template <bool WITH_DW>
__global__ void __launch_bounds__(NT, 1)
lstm_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ h, const float* __restrict__ hx0,
                  const float* __restrict__ u, const float* __restrict__ dhead,
                  const float* __restrict__ Wt, const float* __restrict__ bt, const float* __restrict__ W_ih,
                  float* __restrict__ partials, long long P,
                  long long off_Wt, long long off_bt, long long off_Wih, long long off_Whh,
                  long long off_bih, long long off_bhh, long long off_Whead, long long off_bhead,
                  float* __restrict__ du, int B, int T, int Fin, int NH) {
    B200RL_DYN_SMEM(float, smem);
    float* DtT = smem;                       // DtT[p*DS + r]
    float* AtT = DtT + G * DS;               // AtT[k*DS + r]: k<64 x_t (post-ReLU), k>=64 h_{t-1}
    float* WiT = AtT + KC * DS;              // WiT[p*64 + k] = W_ih[n(p)][k]
    float* HtT = WiT + G * H;                // HtT[k*DS + r] = h_t
    float* DxT = HtT + H * DS;               // DxT[k*DS + r] = dLoss/dx_t masked by ReLU
    float* dHd = DxT + H * DS;               // dHd[r*NH + c]
    float* U = dHd + TM * NH;                // U[r*Fin + q]
    float* WtS = U + TM * Fin;               // WtS[q*64 + j]
    float* bts = WtS + Fin * H;
    float* WtR = bts + H;                    // WtR[j*Fin + q]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const MapR m;
    {
        const int n = tid, g = n >> 6, j = n & 63, p = j * 4 + g;
        for (int k4 = 0; k4 < H; k4 += 4) st4(WiT + p * H + k4, ld4(W_ih + (size_t)n * H + k4));
        if (tid < H) bts[tid] = bt[tid];
        for (int i = tid; i < Fin * H; i += NT) { int q = i / H, jj = i % H; WtS[i] = Wt[jj * Fin + q]; WtR[jj * Fin + q] = Wt[jj * Fin + q]; }
    }
    // dW ownership: 16 gate columns (p) x 8 k.  pg = tid/16 -> p = pg*16..+15 ; kg = tid%16 -> k = kg + 16*e (e<4: x_t, e>=4: h_{t-1});
    // lane-consecutive k rows of the stride-36 tile hit disjoint bank groups (no conflicts)
    const int pg = tid >> 4, kg = tid & 15;
    float acc[16][8];
#pragma unroll
    for (int a = 0; a < 16; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    float acc_b = 0.f, acc_bt = 0.f;          // db[p = tid], dbt[j = tid < 64]
    float acc_hw[WG_E], acc_tw[WG_E], acc_hb = 0.f;
#pragma unroll
    for (int e = 0; e < WG_E; ++e) { acc_hw[e] = 0.f; acc_tw[e] = 0.f; }

    const long long rows = (long long)B * T;
    const long long nchunks = (rows + TM - 1) / TM;
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long row0 = ch * TM;
        __syncthreads();
        // ---- load tiles; lane <-> row so the transposed stores are conflict-free
        {
            const long long row = row0 + lane;
            const bool ok = row < rows;
            const long long rr = ok ? row : rows - 1;
            const int b = (int)(rr / T), t = (int)(rr % T);
#pragma unroll
            for (int it = 0; it < 8; ++it) {          // dpre: 64 float4 per row, 8 per warp-iteration
                const int c4 = warp * 8 + it;
                float4 v = ok ? ld4(dpre + (size_t)rr * G + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                DtT[(c4 * 4 + 0) * DS + lane] = v.x; DtT[(c4 * 4 + 1) * DS + lane] = v.y;
                DtT[(c4 * 4 + 2) * DS + lane] = v.z; DtT[(c4 * 4 + 3) * DS + lane] = v.w;
            }
            const float* hp = (t == 0) ? hx0 + (size_t)b * H : h + (size_t)(rr - 1) * H;
#pragma unroll
            for (int it = 0; it < 2; ++it) {          // h_{t-1} and h_t: 16 float4 per row each
                const int c4 = warp * 2 + it;
                float4 v = ok ? ld4(hp + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                AtT[(H + c4 * 4 + 0) * DS + lane] = v.x; AtT[(H + c4 * 4 + 1) * DS + lane] = v.y;
                AtT[(H + c4 * 4 + 2) * DS + lane] = v.z; AtT[(H + c4 * 4 + 3) * DS + lane] = v.w;
                float4 w = ok ? ld4(h + (size_t)rr * H + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                HtT[(c4 * 4 + 0) * DS + lane] = w.x; HtT[(c4 * 4 + 1) * DS + lane] = w.y;
                HtT[(c4 * 4 + 2) * DS + lane] = w.z; HtT[(c4 * 4 + 3) * DS + lane] = w.w;
            }
            for (int i = tid; i < TM * Fin; i += NT) {
                long long r2 = row0 + i / Fin;
                U[i] = r2 < rows ? u[(size_t)r2 * Fin + i % Fin] : 0.f;
            }
            for (int i = tid; i < TM * NH; i += NT) {
                long long r2 = row0 + i / NH;
                dHd[i] = r2 < rows ? dhead[(size_t)r2 * NH + i % NH] : 0.f;
            }
        }
        __syncthreads();
        // ---- recompute x_t = relu(u . Wt^T + bt) into AtT[0..63]   (rows beyond `rows` -> 0 via zero dpre)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = warp * 8 + e;
            float a = bts[j];
            for (int q = 0; q < Fin; ++q) a = fmaf(U[lane * Fin + q], WtS[q * H + j], a);
            AtT[j * DS + lane] = fmaxf(a, 0.f);
        }
        __syncthreads();
        if constexpr (WITH_DW) {
        // ---- dW[p][k] += sum_r dpre[r][p] * A[r][k]        (256 x 128 x 32 per chunk)
#pragma unroll 1
        for (int r4 = 0; r4 < TM; r4 += 4) {
            float4 av[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                av[e] = ld4(AtT + (kg + 16 * e) * DS + r4);
                av[4 + e] = ld4(AtT + (H + kg + 16 * e) * DS + r4);
            }
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                const float4 d = ld4(DtT + (pg * 16 + a) * DS + r4);
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    acc[a][b] = fmaf(d.x, av[b].x, acc[a][b]); acc[a][b] = fmaf(d.y, av[b].y, acc[a][b]);
                    acc[a][b] = fmaf(d.z, av[b].z, acc[a][b]); acc[a][b] = fmaf(d.w, av[b].w, acc[a][b]);
                }
            }
        }
        }
        if constexpr (WITH_DW) {
        // ---- db[p] (both b_ih and b_hh), p = tid
        {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < TM; ++r) s += DtT[tid * DS + r];
            acc_b += s;
        }
        }
        // ---- dx[r][k] = sum_p dpre[r][p] W_ih[n(p)][k], masked by x > 0   (32 x 64 x 256 per chunk)
        {
            float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
            for (int p = 0; p < G; ++p) {
                float4 d = ld4(DtT + p * DS + m.r0);
                float w0 = WiT[p * H + m.j0], w1 = WiT[p * H + m.j1];
                a0[0] = fmaf(d.x, w0, a0[0]); a0[1] = fmaf(d.y, w0, a0[1]);
                a0[2] = fmaf(d.z, w0, a0[2]); a0[3] = fmaf(d.w, w0, a0[3]);
                a1[0] = fmaf(d.x, w1, a1[0]); a1[1] = fmaf(d.y, w1, a1[1]);
                a1[2] = fmaf(d.z, w1, a1[2]); a1[3] = fmaf(d.w, w1, a1[3]);
            }
            float4 x0 = ld4(AtT + m.j0 * DS + m.r0), x1 = ld4(AtT + m.j1 * DS + m.r0);
            st4(DxT + m.j0 * DS + m.r0, make_float4(x0.x > 0.f ? a0[0] : 0.f, x0.y > 0.f ? a0[1] : 0.f,
                                                    x0.z > 0.f ? a0[2] : 0.f, x0.w > 0.f ? a0[3] : 0.f));
            st4(DxT + m.j1 * DS + m.r0, make_float4(x1.x > 0.f ? a1[0] : 0.f, x1.y > 0.f ? a1[1] : 0.f,
                                                    x1.z > 0.f ? a1[2] : 0.f, x1.w > 0.f ? a1[3] : 0.f));
        }
        // ---- head grads: dW_head[c][k] += sum_r dhead[r][c] h_t[r][k] ; db_head[c]
#pragma unroll
        for (int e = 0; e < WG_E; ++e) {
            const int o = e * NT + tid;
            if (o < NH * H) {
                const int c = o / H, k = o % H;
                float s = 0.f;
#pragma unroll 8
                for (int r = 0; r < TM; ++r) s = fmaf(dHd[r * NH + c], HtT[k * DS + r], s);
                acc_hw[e] += s;
            }
        }
        if (tid < NH) {
            float s = 0.f;
            for (int r = 0; r < TM; ++r) s += dHd[r * NH + tid];
            acc_hb += s;
        }
        __syncthreads();
        // ---- torso grads from DxT: dWt[j][q] += sum_r dxr[r][j] u[r][q] ; dbt[j] ; du[r][q]
#pragma unroll
        for (int e = 0; e < WG_E; ++e) {
            const int o = e * NT + tid;
            if (o < H * Fin) {
                const int j = o % H, q = o / H;
                float s = 0.f;
#pragma unroll 8
                for (int r = 0; r < TM; ++r) s = fmaf(DxT[j * DS + r], U[r * Fin + q], s);
                acc_tw[e] += s;
            }
        }
        if (tid < H) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < TM; ++r) s += DxT[tid * DS + r];
            acc_bt += s;
        }
        if (du != nullptr) {
            for (int i = tid; i < TM * Fin; i += NT) {
                const int r = i / Fin, q = i % Fin;
                if (row0 + r < rows) {
                    float s = 0.f;
                    for (int j = 0; j < H; ++j) s = fmaf(DxT[j * DS + r], WtR[j * Fin + q], s);
                    du[(size_t)(row0 + r) * Fin + q] = s;
                }
            }
        }
    }
    // ---- write this CTA's partial slab (PyTorch layouts: W_* (4H,H) with n = g*64+j)
    float* out = partials + (size_t)blockIdx.x * P;
    if constexpr (WITH_DW) {
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        const int p = pg * 16 + a, j = p >> 2, g = p & 3, n = g * H + j;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            out[off_Wih + (size_t)n * H + kg + 16 * e] = acc[a][e];
            out[off_Whh + (size_t)n * H + kg + 16 * e] = acc[a][4 + e];
        }
    }
    {
        const int p = tid, j = p >> 2, g = p & 3, n = g * H + j;
        out[off_bih + n] = acc_b; out[off_bhh + n] = acc_b;
    }
    }
    if (tid < H) out[off_bt + tid] = acc_bt;
#pragma unroll
    for (int e = 0; e < WG_E; ++e) {
        const int o = e * NT + tid;
        if (o < NH * H) out[off_Whead + o] = acc_hw[e];
        if (o < H * Fin) { const int j = o % H, q = o / H; out[off_Wt + (size_t)j * Fin + q] = acc_tw[e]; }
    }
    if (tid < NH) out[off_bhead + tid] = acc_hb;
}

// grad[i] = sum over the slabs: a CTA owns 32 consecutive parameters (one 128-byte line per slab row), its 8 warps take the
// slabs y, y + 8, ... with four independent loads in flight each, and the eight partial sums are added in a fixed order
// (deterministic).  (One thread per parameter walking all ~300 slabs serially was a 300-deep dependent-load chain: 16 us.)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, float* __restrict__ grad, int nparts, long long P) {
#ifdef B200RL_EMUL
    // (host emulator: a 256-thread barrier costs a millisecond of context switches -- same sums, the same order per slab
    // group, without the block-wide exchange)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        float part[8];
        for (int y = 0; y < 8; ++y) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int p = y;
            for (; p + 24 < nparts; p += 32) {
                s0 += partials[(size_t)p * P + i]; s1 += partials[(size_t)(p + 8) * P + i];
                s2 += partials[(size_t)(p + 16) * P + i]; s3 += partials[(size_t)(p + 24) * P + i];
            }
            for (; p < nparts; p += 8) s0 += partials[(size_t)p * P + i];
            part[y] = (s0 + s1) + (s2 + s3);
        }
        float s = part[0];
        for (int k = 1; k < 8; ++k) s += part[k];
        grad[i] = s;
    }
    return;
#endif
    __shared__ float red[8][32];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    for (long long i0 = (long long)blockIdx.x * 32; i0 < P; i0 += (long long)gridDim.x * 32) {
        const long long i = i0 + x;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        if (i < P) {
            int p = y;
            for (; p + 24 < nparts; p += 32) {
                s0 += partials[(size_t)p * P + i]; s1 += partials[(size_t)(p + 8) * P + i];
                s2 += partials[(size_t)(p + 16) * P + i]; s3 += partials[(size_t)(p + 24) * P + i];
            }
            for (; p < nparts; p += 8) s0 += partials[(size_t)p * P + i];
        }
        red[y][x] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (y == 0 && i < P) {
            float s = red[0][x];
#pragma unroll
            for (int k = 1; k < 8; ++k) s += red[k][x];
            grad[i] = s;
        }
        __syncthreads();
    }
}

#ifndef B200RL_EMUL
template <typename K>
static int set_smem(K kern, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error("%s: cudaFuncSetAttribute(%zu B): %s", what, bytes, cudaGetErrorString(e));
    return 0;
}
static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
#else
template <typename K> static int set_smem(K, size_t, const char*) { return 0; }
static int sm_count() { return 4; }
#endif

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_lstm_fwd(const float* u, const float* hx0, const float* cx0,
                               const float* Wt, const float* bt,
                               const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh,
                               float* gates, float* c, float* h,
                               int B, int T, int Fin, int variant, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1, "lstm_fwd: bad shape B=%d T=%d", B, T);
    B200RL_REQUIRE(Fin >= 1 && Fin <= MAXF, "lstm_fwd: Fin=%d out of range [1,%d]", Fin, MAXF);
    B200RL_REQUIRE(variant == 0 || variant == 3, "lstm_fwd: unknown variant %d (0 = FFMA reference kernel, 3 = tcgen05 projection + tcgen05 recurrence)", variant);
    const int ntiles = (B + TM - 1) / TM;
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    if (variant == 3) {
        // input projection AND the recurrent GEMM on tcgen05 (two kernels; variant 5 = b200rl_lstm_fwd_fused is ONE)
        if (launch_xproj_tc(u, Wt, bt, W_ih, b_ih, b_hh, gates, B, T, Fin, sm_count(), stream)) return 1;
        return launch_fwd_tc(gates, hx0, cx0, W_hh, c, h, B, T, sm_count(), stream);
    }
    const size_t smem = fwd_smem_bytes(Fin);
    if (set_smem(lstm_fwd_kernel, smem, "lstm_fwd")) return 1;
    B200RL_LAUNCH(lstm_fwd_kernel, grid, NT, smem, stream, u, hx0, cx0, Wt, bt, W_ih, W_hh, b_ih, b_hh, gates, c, h, B, T, Fin);
    return check_launch("lstm_fwd");
}

extern "C" int b200rl_head_fwd(const float* h, const float* W_head, const float* b_head,
                               float* head_out, long long rows, int NH, void* stream) {
    B200RL_REQUIRE(rows >= 1 && NH >= 1 && NH <= B200RL_MAX_HEADS, "head_fwd: bad shape rows=%lld NH=%d", rows, NH);
    const long long nblk = (rows + HR - 1) / HR;
    const long long cap = (long long)sm_count() * 4;
    const size_t hsm = sizeof(float) * (size_t)(HR * HFS + 32 * H + 32);
    if (NH <= 4) {
        if (set_smem(head_fwd_kernel<true>, hsm, "head_fwd")) return 1;
        B200RL_LAUNCH(head_fwd_kernel<true>, (int)(nblk < cap ? nblk : cap), HR, hsm, stream, h, W_head, b_head, head_out, rows, NH);
    } else {
        if (set_smem(head_fwd_kernel<false>, hsm, "head_fwd")) return 1;
        B200RL_LAUNCH(head_fwd_kernel<false>, (int)(nblk < cap ? nblk : cap), HR, hsm, stream, h, W_head, b_head, head_out, rows, NH);
    }
    return check_launch("head_fwd");
}

extern "C" int b200rl_lstm_bwd(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih,
                               float* gates, const float* c, const float* cx0, float* dx,
                               int B, int T, int NH, int variant, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1, "lstm_bwd: bad shape B=%d T=%d", B, T);
    B200RL_REQUIRE(NH >= 1 && NH <= B200RL_MAX_HEADS, "lstm_bwd: NH=%d out of range", NH);
    B200RL_REQUIRE(variant == 0 || variant == 1, "lstm_bwd: unknown variant %d", variant);
    if (variant == 1) return launch_bwd_tc(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, B, T, NH, sm_count(), stream);
    const size_t smem = bwd_smem_bytes(NH);
    if (set_smem(lstm_bwd_kernel, smem, "lstm_bwd")) return 1;
    const int ntiles = (B + TM - 1) / TM;
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    B200RL_LAUNCH(lstm_bwd_kernel, grid, NT, smem, stream, dhead, W_head, W_hh, gates, c, cx0, B, T, NH);
    return check_launch("lstm_bwd");
}

extern "C" int b200rl_lstm_bwd_fused(const float* dhead, const float* W_head, const float* W_hh, const float* W_ih,
                                     float* gates, const float* c, const float* cx0, const float* u, const float* Wt, const float* bt,
                                     float* partials, int nparts, long long P, long long off_Wt, long long off_bt,
                                     long long off_Whead, long long off_bhead, float* dx, int B, int T, int Fin, int NH, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1, "lstm_bwd_fused: bad shape B=%d T=%d", B, T);
    B200RL_REQUIRE(partials != nullptr, "lstm_bwd_fused: needs the partial-gradient slabs");
    const int ntiles = (B + 31) / 32;
    B200RL_REQUIRE(nparts >= (ntiles < sm_count() ? ntiles : sm_count()), "lstm_bwd_fused: nparts=%d below the grid size", nparts);
    return launch_bwd_tc_fused(dhead, W_head, W_hh, W_ih, gates, c, cx0, dx, u, Wt, bt, partials, P, off_Wt, off_bt, off_Whead, off_bhead,
                               B, T, Fin, NH, sm_count(), stream);
}

extern "C" int b200rl_lstm_wgrad(const float* dpre, const float* h, const float* hx0, const float* u,
                                 const float* dhead, const float* Wt, const float* bt, const float* W_ih,
                                 float* partials, int nparts, long long P,
                                 long long off_Wt, long long off_bt, long long off_Wih, long long off_Whh,
                                 long long off_bih, long long off_bhh, long long off_Whead, long long off_bhead,
                                 float* du, const float* dx, int B, int T, int Fin, int NH, int variant, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1, "lstm_wgrad: bad shape B=%d T=%d", B, T);
    B200RL_REQUIRE(Fin >= 1 && Fin <= MAXF && NH >= 1 && NH <= B200RL_MAX_HEADS, "lstm_wgrad: Fin=%d NH=%d out of range", Fin, NH);
    B200RL_REQUIRE(H * Fin <= WG_E * NT && NH * H <= WG_E * NT, "lstm_wgrad: head/torso too wide");
    B200RL_REQUIRE(variant >= 0 && variant <= 5, "lstm_wgrad: unknown variant %d", variant);
    B200RL_REQUIRE(nparts >= 1, "lstm_wgrad: nparts=%d", nparts);
    B200RL_REQUIRE((off_Wih % 4) == 0 && (off_Whh % 4) == 0 && (P % 4) == 0, "lstm_wgrad: W_ih/W_hh offsets and P must be multiples of 4 floats");
    const size_t smem = wgrad_smem_bytes(Fin, NH);
    // wgrad_tc runs two CTAs per SM and takes every slab; the other kernels of the stage fill an SM with one CTA and use the
    // first sm_count slabs (their entries in the remaining slabs stay zero, as allocated)
    const int np_misc = nparts < sm_count() ? nparts : sm_count();
    if (variant == 4) {
        // lstm_bwd_fused already accumulated the head and torso gradients: only dW_ih / dW_hh / db_ih / db_hh are left
        B200RL_REQUIRE((long long)nparts * 32 <= (long long)B * T + 31, "lstm_wgrad: nparts=%d exceeds the number of 32-row stages", nparts);
        return launch_wgrad_tc(dpre, h, hx0, u, Wt, bt, partials, nparts, P, off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin, stream);
    }
    if (variant == 5) {
        // lstm_bwd_fused accumulated the head gradients and stored dx: tcgen05 dW + the torso pass over dx
        B200RL_REQUIRE(dx != nullptr && du == nullptr, "lstm_wgrad: variant 5 needs dx and does not produce du");
        B200RL_REQUIRE((long long)nparts * 32 <= (long long)B * T + 31, "lstm_wgrad: nparts=%d exceeds the number of 32-row stages", nparts);
        if (launch_wgrad_tc(dpre, h, hx0, u, Wt, bt, partials, nparts, P, off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin, stream)) return 1;
        return launch_torso_from_dx(dx, u, Wt, bt, nullptr, dhead, partials, np_misc, P, off_Wt, off_bt, off_Whead, off_bhead, B, T, Fin, NH, stream);
    }
    if (variant == 3) {
        // dx was already produced by lstm_bwd variant 1: tcgen05 dW + a bandwidth-bound torso pass over dx
        B200RL_REQUIRE(dx != nullptr && du == nullptr, "lstm_wgrad: variant 3 needs dx (from lstm_bwd variant 1) and does not produce du");
        B200RL_REQUIRE((long long)nparts * 32 <= (long long)B * T + 31, "lstm_wgrad: nparts=%d exceeds the number of 32-row stages", nparts);
        if (launch_wgrad_tc(dpre, h, hx0, u, Wt, bt, partials, nparts, P, off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin, stream)) return 1;
        return launch_torso_from_dx(dx, u, Wt, bt, h, dhead, partials, np_misc, P, off_Wt, off_bt, off_Whead, off_bhead, B, T, Fin, NH, stream);
    }
    if (variant == 1) {
        // tensor-core path: dW_ih / dW_hh / db (wgrad_tc.cu) and dx -> torso grads (dx_tc.cu) on tcgen05
        B200RL_REQUIRE((long long)nparts * 32 <= (long long)B * T + 31, "lstm_wgrad: nparts=%d exceeds the number of 32-row stages", nparts);
        if (launch_wgrad_tc(dpre, h, hx0, u, Wt, bt, partials, nparts, P, off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin, stream)) return 1;
        return launch_dx_tc(dpre, u, W_ih, Wt, bt, h, dhead, partials, np_misc, P, off_Wt, off_bt, off_Whead, off_bhead,
                            du, B, T, Fin, NH, stream);
    }
    if (variant == 2) {   // tcgen05 dW + FFMA remainder (kept for A/B measurements)
        if (launch_wgrad_tc(dpre, h, hx0, u, Wt, bt, partials, nparts, P, off_Wih, off_Whh, off_bih, off_bhh, B, T, Fin, stream)) return 1;
        if (set_smem(lstm_wgrad_kernel<false>, smem, "lstm_wgrad")) return 1;
        B200RL_LAUNCH(lstm_wgrad_kernel<false>, nparts, NT, smem, stream, dpre, h, hx0, u, dhead, Wt, bt, W_ih, partials, P,
                      off_Wt, off_bt, off_Wih, off_Whh, off_bih, off_bhh, off_Whead, off_bhead, du, B, T, Fin, NH);
        return check_launch("lstm_wgrad(misc)");
    }
    if (set_smem(lstm_wgrad_kernel<true>, smem, "lstm_wgrad")) return 1;
    B200RL_LAUNCH(lstm_wgrad_kernel<true>, nparts, NT, smem, stream, dpre, h, hx0, u, dhead, Wt, bt, W_ih, partials, P,
                  off_Wt, off_bt, off_Wih, off_Whh, off_bih, off_bhh, off_Whead, off_bhead, du, B, T, Fin, NH);
    return check_launch("lstm_wgrad");
}

extern "C" int b200rl_reduce_partials(const float* partials, float* grad, int nparts, long long P, void* stream) {
    B200RL_REQUIRE(nparts >= 1 && P >= 1, "reduce_partials: bad shape");
    long long g = (P + 31) / 32;
#ifdef B200RL_EMUL
    if (g > 4) g = 4;                   // (the host emulator spawns a thread per CUDA thread: keep the grid small)
#endif
    B200RL_LAUNCH(reduce_partials_kernel, (int)(g > 148 * 8 ? 148 * 8 : g), 256, 0, stream, partials, grad, nparts, P);
    return check_launch("reduce_partials");
}
