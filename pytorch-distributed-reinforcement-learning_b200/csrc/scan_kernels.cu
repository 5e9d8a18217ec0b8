// Per-timestep recurrences of the learner update as warp-scan kernels (sm_100a).
//
//   compute_gae      agents/learner_module/compute_loss.py:7-19
//   td + GAE fused   agents/learner_module/ppo/learning.py:48-57
//   compute_v_trace  agents/learner_module/compute_loss.py:22-66
//   kldivergence     agents/learner_module/compute_loss.py:74-77
//   soft_update      agents/learner_module/compute_loss.py:69-71
//
// Layout: (B, T, 1) fp32, rows contiguous along T.  One warp owns one row; lanes run
// along T in slabs of 32 consecutive timesteps so every global load/store of a warp is
// one coalesced 128-byte line.  The recurrences are first-order linear
// (y[t] = a[t]*y[t+1] + b[t]) and are evaluated as a suffix scan of affine maps with
// warp shuffles (5 steps per slab), slabs processed from the end of the row with a
// scalar carry.  HBM-bound: 8 B/elem (gae), 20 B/elem (td+gae), 32 B/elem (v-trace).
#include "common.cuh"
#include "scan.cuh"

namespace b200rl {

constexpr int kScanThreads = 256;
constexpr int kSuper = 4;  // slabs whose loads are issued together (memory-level parallelism)

static inline int scan_grid(int rows) {
    int warps_per_cta = kScanThreads / 32;
    long long ctas = ((long long)rows + warps_per_cta - 1) / warps_per_cta;
    long long cap = 148LL * 8;
    return (int)(ctas < 1 ? 1 : (ctas > cap ? cap : ctas));
}

// ------------------------------------------------------------------------------ GAE
__global__ void __launch_bounds__(kScanThreads)
gae_kernel(const float* __restrict__ deltas, float* __restrict__ gae, int B, int n, float k) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int nslab = (n + 31) >> 5;
    for (int row = wid; row < B; row += nw) {
        const float* d = deltas + (size_t)row * n;
        float* g = gae + (size_t)row * n;
        float carry = 0.0f;
        for (int s_hi = nslab; s_hi > 0; s_hi -= kSuper) {
            float b[kSuper];
#pragma unroll
            for (int j = 0; j < kSuper; ++j) {
                int s = s_hi - 1 - j, t = s * 32 + lane;
                b[j] = (s >= 0 && t < n) ? __ldg(d + t) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < kSuper; ++j) {
                int s = s_hi - 1 - j, t = s * 32 + lane;
                if (s < 0) break;                       // warp-uniform
                bool ok = t < n;
                float y = affine_suffix_scan(ok ? k : 1.0f, b[j], carry);
                if (ok) g[t] = y;
                carry = __shfl_sync(kFull, y, 0);
            }
        }
    }
}

// ----------------------------------------------------------------- td target + GAE
__global__ void __launch_bounds__(kScanThreads)
td_gae_kernel(const float* __restrict__ rew, const float* __restrict__ is_fir,
              const float* __restrict__ value, float* __restrict__ td_out,
              float* __restrict__ gae, int B, int T, float gamma, float lam) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = T - 1;
    const int nslab = (n + 31) >> 5;
    const float k = gamma * lam;
    for (int row = wid; row < B; row += nw) {
        const float* r = rew + (size_t)row * T;
        const float* f = is_fir + (size_t)row * T;
        const float* v = value + (size_t)row * T;
        float* tdo = td_out + (size_t)row * n;
        float* g = gae + (size_t)row * n;
        float carry = 0.0f;
        for (int s_hi = nslab; s_hi > 0; s_hi -= kSuper) {
            float td[kSuper], dl[kSuper];
#pragma unroll
            for (int j = 0; j < kSuper; ++j) {
                int s = s_hi - 1 - j, t = s * 32 + lane;
                td[j] = 0.0f; dl[j] = 0.0f;
                if (s >= 0 && t < n) {
                    float rt = __ldg(r + t), f1 = __ldg(f + t + 1), vt = __ldg(v + t), v1 = __ldg(v + t + 1);
                    td[j] = rt + (gamma * (1.0f - f1)) * v1;       // ppo/learning.py:49-52
                    dl[j] = td[j] - vt;
                }
            }
#pragma unroll
            for (int j = 0; j < kSuper; ++j) {
                int s = s_hi - 1 - j, t = s * 32 + lane;
                if (s < 0) break;
                bool ok = t < n;
                float y = affine_suffix_scan(ok ? k : 1.0f, dl[j], carry);
                if (ok) { g[t] = y; tdo[t] = td[j]; }
                carry = __shfl_sync(kFull, y, 0);
            }
        }
    }
}

// ------------------------------------------- td target + GAE, four timesteps per lane
// The shuffle scan costs ~45 instructions per 32 elements, which is what bounds td_gae_kernel at large B*T (ncu r2j: 3.2 TB/s,
// issue-bound).  Here a lane owns FOUR consecutive timesteps (one 16-byte load per input row segment), folds them into one
// affine map serially, and the warp scans 128 timesteps with ONE 5-step shuffle scan.  The (B, T-1) outputs have 4-byte-
// aligned rows, so the results are transposed through a 1 KB per-warp shared-memory tile and stored as full 128-byte lines.
// Needs T % 4 == 0 and 16-byte aligned inputs (the launcher falls back to td_gae_kernel otherwise).
__global__ void __launch_bounds__(kScanThreads)
td_gae4_kernel(const float* __restrict__ rew, const float* __restrict__ is_fir,
               const float* __restrict__ value, float* __restrict__ td_out,
               float* __restrict__ gae, int B, int T, float gamma, float lam) {
    __shared__ __align__(16) float stage[kScanThreads / 32][2][128];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = T - 1;
    const int nchunk = (T + 127) >> 7;
    const float k = gamma * lam;
    float* s_td = stage[wl][0];
    float* s_g = stage[wl][1];
    for (int row = wid; row < B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        float carry = 0.0f;
        for (int c = nchunk - 1; c >= 0; --c) {
            const int t0 = c << 7, t = t0 + 4 * lane;
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f), f4 = r4, v4 = r4;
            if (t < T) { r4 = ld4(rew + o + t); f4 = ld4(is_fir + o + t); v4 = ld4(value + o + t); }
            float vn = __shfl_down_sync(kFull, v4.x, 1), fn = __shfl_down_sync(kFull, f4.x, 1);
            if (lane == 31) {
                const bool more = t0 + 128 < T;
                vn = more ? __ldg(value + o + t0 + 128) : 0.f;
                fn = more ? __ldg(is_fir + o + t0 + 128) : 0.f;
            }
            const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, vv[5] = {v4.x, v4.y, v4.z, v4.w, vn}, ff[5] = {f4.x, f4.y, f4.z, f4.w, fn};
            float td[4], dl[4];
            bool ok[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ok[e] = t + e < n;
                td[e] = ok[e] ? rr[e] + (gamma * (1.0f - ff[e + 1])) * vv[e + 1] : 0.f;      // ppo/learning.py:49-52
                dl[e] = ok[e] ? td[e] - vv[e] : 0.f;
            }
            float a = 1.f, b = 0.f;                           // y[t] = a * y[t+4] + b for this lane's four steps
#pragma unroll
            for (int e = 3; e >= 0; --e) if (ok[e]) { b = fmaf(k, b, dl[e]); a *= k; }
            const float y0 = affine_suffix_scan(a, b, carry);
            float Y = __shfl_down_sync(kFull, y0, 1);
            if (lane == 31) Y = carry;
            float g[4];
#pragma unroll
            for (int e = 3; e >= 0; --e) { Y = ok[e] ? fmaf(k, Y, dl[e]) : Y; g[e] = Y; }
            carry = __shfl_sync(kFull, y0, 0);
            __syncwarp();
            st4(s_td + 4 * lane, make_float4(td[0], td[1], td[2], td[3]));
            st4(s_g + 4 * lane, make_float4(g[0], g[1], g[2], g[3]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = i * 32 + lane;
                if (t0 + idx < n) { td_out[on + t0 + idx] = s_td[idx]; gae[on + t0 + idx] = s_g[idx]; }
            }
        }
    }
}

// ----------------------------------------------------- GAE, four timesteps per lane
// compute_gae's rows have n = T-1 elements (4-byte aligned), so both directions go through the per-warp shared-memory tile:
// coalesced 4-byte loads in, 16 bytes per lane out of the tile, one shuffle scan per 128 steps, and back the same way.
__global__ void __launch_bounds__(kScanThreads)
gae4_kernel(const float* __restrict__ deltas, float* __restrict__ gae, int B, int n, float k) {
    __shared__ __align__(16) float stage[kScanThreads / 32][128];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int nchunk = (n + 127) >> 7;
    float* st = stage[wl];
    for (int row = wid; row < B; row += nw) {
        const float* d = deltas + (size_t)row * n;
        float* g = gae + (size_t)row * n;
        float carry = 0.0f;
        for (int c = nchunk - 1; c >= 0; --c) {
            const int t0 = c << 7;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) { const int idx = i * 32 + lane; st[idx] = t0 + idx < n ? __ldg(d + t0 + idx) : 0.f; }
            __syncwarp();
            const float4 d4 = ld4(st + 4 * lane);
            const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
            const int t = t0 + 4 * lane;
            float a = 1.f, b = 0.f;
#pragma unroll
            for (int e = 3; e >= 0; --e) if (t + e < n) { b = fmaf(k, b, dl[e]); a *= k; }
            const float y0 = affine_suffix_scan(a, b, carry);
            float Y = __shfl_down_sync(kFull, y0, 1);
            if (lane == 31) Y = carry;
            float y[4];
#pragma unroll
            for (int e = 3; e >= 0; --e) { Y = (t + e < n) ? fmaf(k, Y, dl[e]) : Y; y[e] = Y; }
            carry = __shfl_sync(kFull, y0, 0);
            __syncwarp();
            st4(st + 4 * lane, make_float4(y[0], y[1], y[2], y[3]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) { const int idx = i * 32 + lane; if (t0 + idx < n) g[t0 + idx] = st[idx]; }
        }
    }
}

// ------------------------------------------------- V-trace, four timesteps per lane (T % 4 == 0, 16-byte aligned inputs)
__global__ void __launch_bounds__(kScanThreads)
vtrace4_kernel(const float* __restrict__ blp, const float* __restrict__ tlp,
               const float* __restrict__ is_fir, const float* __restrict__ rew,
               const float* __restrict__ values, float* __restrict__ rho_out,
               float* __restrict__ adv_out, float* __restrict__ vs_out,
               int B, int T, float gamma, float rho_bar, float c_bar) {
    __shared__ __align__(16) float stage[kScanThreads / 32][2][128];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = T - 1;
    const int nchunk = (T + 127) >> 7;
    float* s_rho = stage[wl][0];
    float* s_adv = stage[wl][1];
    for (int row = wid; row < B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        float carry = 0.0f;                                   // acc at the first step of the chunk processed before (later in time)
        for (int c = nchunk - 1; c >= 0; --c) {
            const int t0 = c << 7, t = t0 + 4 * lane;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 b4 = z, l4 = z, f4 = z, r4 = z, v4 = z;
            if (t < T) { b4 = ld4(blp + o + t); l4 = ld4(tlp + o + t); f4 = ld4(is_fir + o + t); r4 = ld4(rew + o + t); v4 = ld4(values + o + t); }
            float vn = __shfl_down_sync(kFull, v4.x, 1), fn = __shfl_down_sync(kFull, f4.x, 1);
            if (lane == 31) {
                const bool more = t0 + 128 < T;
                vn = more ? __ldg(values + o + t0 + 128) : 0.f;
                fn = more ? __ldg(is_fir + o + t0 + 128) : 0.f;
            }
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
            const float vv[5] = {v4.x, v4.y, v4.z, v4.w, vn}, ff[5] = {f4.x, f4.y, f4.z, f4.w, fn};
            float rc[4], am[4], disc[4], dl[4];
            bool ok[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ok[e] = t + e < n;                            // scan / rho / adv are defined for t < T-1
                const float rho = expf(ll[e] - bb[e]);
                rc[e] = fminf(fmaxf(rho, 0.1f), rho_bar);     // compute_loss.py:37
                disc[e] = gamma * (1.0f - ff[e + 1]);
                am[e] = fminf(rho, c_bar) * disc[e];          // compute_loss.py:43
                dl[e] = rc[e] * (rr[e] + disc[e] * vv[e + 1] - vv[e]);
            }
            float a = 1.f, b = 0.f;
#pragma unroll
            for (int e = 3; e >= 0; --e) if (ok[e]) { b = fmaf(am[e], b, dl[e]); a *= am[e]; }
            const float y0 = affine_suffix_scan(a, b, carry);
            float Y = __shfl_down_sync(kFull, y0, 1);
            if (lane == 31) Y = carry;
            float acc[5];
            acc[4] = Y;                                       // acc at t + 4
#pragma unroll
            for (int e = 3; e >= 0; --e) acc[e] = ok[e] ? fmaf(am[e], acc[e + 1], dl[e]) : (t + e == n ? 0.0f : acc[e + 1]);   // acc[T-1] = 0
            carry = __shfl_sync(kFull, y0, 0);
            float adv[4], vs[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                vs[e] = vv[e] + (t + e >= n ? 0.0f : acc[e]);
                const float vs1 = vv[e + 1] + ((t + e + 1 >= n) ? 0.0f : acc[e + 1]);
                adv[e] = rc[e] * (rr[e] + disc[e] * vs1 - vv[e]);
            }
            if (t < T) st4(vs_out + o + t, make_float4(vs[0], vs[1], vs[2], vs[3]));
            __syncwarp();
            st4(s_rho + 4 * lane, make_float4(rc[0], rc[1], rc[2], rc[3]));
            st4(s_adv + 4 * lane, make_float4(adv[0], adv[1], adv[2], adv[3]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = i * 32 + lane;
                if (t0 + idx < n) { rho_out[on + t0 + idx] = s_rho[idx]; adv_out[on + t0 + idx] = s_adv[idx]; }
            }
        }
    }
}

// -------------------------------------------------------------------------- V-trace
__global__ void __launch_bounds__(kScanThreads)
vtrace_kernel(const float* __restrict__ blp, const float* __restrict__ tlp,
              const float* __restrict__ is_fir, const float* __restrict__ rew,
              const float* __restrict__ values, float* __restrict__ rho_out,
              float* __restrict__ adv_out, float* __restrict__ vs_out,
              int B, int T, float gamma, float rho_bar, float c_bar) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = T - 1;
    const int nslab = (T + 31) >> 5;   // slabs cover t = 0..T-1 (vs has T entries)
    for (int row = wid; row < B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        float carry = 0.0f;            // acc[t0+32] of the slab processed before (later in time)
        for (int s = nslab - 1; s >= 0; --s) {
            const int t = s * 32 + lane;
            const bool ok = t < n;     // scan / rho / adv are defined for t < T-1
            float rc = 0.f, cc = 1.f, disc1 = 0.f, rt = 0.f, vt = 0.f, v1 = 0.f, delta = 0.f;
            if (t < T) vt = __ldg(values + o + t);
            if (ok) {
                float rho = expf(__ldg(tlp + o + t) - __ldg(blp + o + t));
                rc = fminf(fmaxf(rho, 0.1f), rho_bar);            // compute_loss.py:37
                cc = fminf(rho, c_bar);                           // compute_loss.py:43
                disc1 = gamma * (1.0f - __ldg(is_fir + o + t + 1));
                rt = __ldg(rew + o + t);
                v1 = __ldg(values + o + t + 1);
                delta = rc * (rt + disc1 * v1 - vt);
            }
            float acc = affine_suffix_scan(ok ? cc * disc1 : 1.0f, delta, carry);
            if (!ok) acc = (t == n) ? 0.0f : acc;                 // acc[T-1] = 0
            float acc_next = __shfl_down_sync(kFull, acc, 1);
            if (lane == 31) acc_next = carry;
            if (t < T) vs_out[o + t] = vt + (t == n ? 0.0f : acc);
            if (ok) {
                float vs1 = v1 + ((t + 1 == n) ? 0.0f : acc_next);
                rho_out[on + t] = rc;
                adv_out[on + t] = rc * (rt + disc1 * vs1 - vt);
            }
            carry = __shfl_sync(kFull, acc, 0);
        }
    }
}

// -------------------------------------------------------------- KL(Cat p || Cat q)
__global__ void __launch_bounds__(256)
kl_kernel(const float* __restrict__ lp, const float* __restrict__ lq, float* __restrict__ out,
          long long rows, int A) {
    const float eps = 1.1920929e-07f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows;
         i += (long long)gridDim.x * blockDim.x) {
        const float* a = lp + i * A;
        const float* b = lq + i * A;
        float ma = -INFINITY, mb = -INFINITY;
        for (int k = 0; k < A; ++k) { ma = fmaxf(ma, a[k]); mb = fmaxf(mb, b[k]); }
        float sa = 0.f, sb = 0.f;
        for (int k = 0; k < A; ++k) { sa += expf(a[k] - ma); sb += expf(b[k] - mb); }
        // softmax, then Categorical(probs=.) renormalises and clamps before the log
        float na = 0.f, nb = 0.f;
        for (int k = 0; k < A; ++k) { na += expf(a[k] - ma) / sa; nb += expf(b[k] - mb) / sb; }
        float kl = 0.f;
        for (int k = 0; k < A; ++k) {
            float p = expf(a[k] - ma) / sa / na, q = expf(b[k] - mb) / sb / nb;
            float t = p * (logf(fminf(fmaxf(p, eps), 1.f - eps)) - logf(fminf(fmaxf(q, eps), 1.f - eps)));
            if (q == 0.f) t = INFINITY;
            if (p == 0.f) t = 0.f;
            kl += t;
        }
        out[i] = kl;
    }
}

__global__ void __launch_bounds__(256)
soft_update_kernel(const float* __restrict__ src, float* __restrict__ tgt, long long n, float tau) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float t = tgt[i], s = src[i];
        tgt[i] = (1.0f - tau) * t + tau * s;               // compute_loss.py:71
    }
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_compute_gae(const float* deltas, float* gae, int B, int n,
                                  float gamma, float lambda_, void* stream) {
    B200RL_REQUIRE(B >= 0 && n >= 0, "compute_gae: bad shape B=%d n=%d", B, n);
    if (B == 0 || n == 0) return 0;
    if (n >= 64) { B200RL_LAUNCH(gae4_kernel, scan_grid(B), kScanThreads, 0, stream, deltas, gae, B, n, gamma * lambda_); }
    else { B200RL_LAUNCH(gae_kernel, scan_grid(B), kScanThreads, 0, stream, deltas, gae, B, n, gamma * lambda_); }
    return check_launch("compute_gae");
}

extern "C" int b200rl_td_gae(const float* rew, const float* is_fir, const float* value,
                             float* td_target, float* gae, int B, int T,
                             float gamma, float lambda_, void* stream) {
    B200RL_REQUIRE(B >= 0 && T >= 1, "td_gae: bad shape B=%d T=%d", B, T);
    if (B == 0 || T == 1) return 0;
    const bool vec = (T % 4) == 0 && (((uintptr_t)rew | (uintptr_t)is_fir | (uintptr_t)value) & 15) == 0;
    if (vec) { B200RL_LAUNCH(td_gae4_kernel, scan_grid(B), kScanThreads, 0, stream, rew, is_fir, value, td_target, gae, B, T, gamma, lambda_); }
    else { B200RL_LAUNCH(td_gae_kernel, scan_grid(B), kScanThreads, 0, stream, rew, is_fir, value, td_target, gae, B, T, gamma, lambda_); }
    return check_launch("td_gae");
}

extern "C" int b200rl_compute_v_trace(const float* behav_logp, const float* target_logp,
                                      const float* is_fir, const float* rew, const float* values,
                                      float* rho, float* adv, float* vs, int B, int T,
                                      float gamma, float rho_bar, float c_bar, void* stream) {
    B200RL_REQUIRE(B >= 0 && T >= 1, "compute_v_trace: bad shape B=%d T=%d", B, T);
    if (B == 0) return 0;
    const bool vec = (T % 4) == 0 && (((uintptr_t)behav_logp | (uintptr_t)target_logp | (uintptr_t)is_fir | (uintptr_t)rew |
                                       (uintptr_t)values | (uintptr_t)vs) & 15) == 0;
    if (vec) {
        B200RL_LAUNCH(vtrace4_kernel, scan_grid(B), kScanThreads, 0, stream, behav_logp, target_logp, is_fir, rew, values,
                      rho, adv, vs, B, T, gamma, rho_bar, c_bar);
    } else {
        B200RL_LAUNCH(vtrace_kernel, scan_grid(B), kScanThreads, 0, stream, behav_logp, target_logp, is_fir, rew, values,
                      rho, adv, vs, B, T, gamma, rho_bar, c_bar);
    }
    return check_launch("compute_v_trace");
}

extern "C" int b200rl_kldivergence(const float* logits_p, const float* logits_q, float* kl,
                                   long long rows, int A, void* stream) {
    B200RL_REQUIRE(rows >= 0 && A >= 1, "kldivergence: bad shape");
    if (rows == 0) return 0;
    long long g = (rows + 255) / 256;
    B200RL_LAUNCH(kl_kernel, (int)(g > 148 * 16 ? 148 * 16 : g), 256, 0, stream, logits_p, logits_q, kl, rows, A);
    return check_launch("kldivergence");
}

extern "C" int b200rl_soft_update(const float* src, float* target, long long n, float tau, void* stream) {
    if (n <= 0) return 0;
    long long g = (n + 255) / 256;
    B200RL_LAUNCH(soft_update_kernel, (int)(g > 148 * 16 ? 148 * 16 : g), 256, 0, stream, src, target, n, tau);
    return check_launch("soft_update");
}
