// V-MPO loss head (sm_100a).
//
// Replaces the epoch body of agents/learner_module/v_mpo/learning.py:44-99 and the dual
// variables / their RMSprop groups of agents/learner.py:320-348:
//   gae            = compute_gae(td - V)                                  (no grad)
//   top_gae, idx   = topk(gae, ceil(B/2), dim=0)      per TIMESTEP, across the whole batch
//   psi            = softmax(flatten(top_gae / (eta + 1e-7)))             over all selected
//   L_pi           = -sum psi * logp[idx]         L_V = mean smooth_l1(V - td)
//   L_eta          = eta*coef_eta + eta*log(mean(exp(top_gae/eta)))
//   L_alpha        = mean(alpha*(eps_alpha - KL.detach()) + alpha.detach()*KL),
//                    KL = KL(Cat(softmax(behaviour_logits)) || Cat(softmax(logits)))
//
// Four small launches over (B, T) scalars:
//   vmpo_gae     one warp per sequence: distribution-free part (V, td, GAE scan), value
//                loss + dV, writes the (B, T-1) advantage matrix;
//   vmpo_select  one CTA per timestep column: 4-pass radix select of the k-th largest
//                advantage over the batch (keys in shared memory, 256-bin histograms), the
//                column max and the column's sum of exp((g - colmax)/eta);
//   vmpo_grad    one warp per sequence: psi from the column statistics, policy + KL
//                gradients w.r.t. the head pre-activations, loss sums in fp64;
//   vmpo_dual    one thread: temperature / KL-multiplier losses, their gradients and the
//                RMSprop step of log_eta, log_alpha in fp64 (the reference keeps them fp64).
// When the batch is sharded over GPUs the advantage matrices are all-gathered between
// vmpo_gae and vmpo_select, so every rank selects against the global batch.
#include "common.cuh"
#include "scan.cuh"
#include "dist.cuh"

namespace b200rl {

__global__ void __launch_bounds__(kHeadThreads)
vmpo_gae_kernel(const float* __restrict__ head_out, const float* __restrict__ rew, const float* __restrict__ is_fir,
                float* __restrict__ gae_out, float* __restrict__ dhead, double* __restrict__ scalars,
                int B, int T, int NH, float gamma, float lam, float c_val, float invN) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int n = T - 1, nslab = (T + 31) >> 5;
    const float k = gamma * lam;
    double s_val = 0;
    for (int row = wid; row < B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        float carry = 0.f;
        for (int s = nslab - 1; s >= 0; --s) {
            const int t = s * 32 + lane;
            const bool in = t < T, ok = t < n;
            float v = 0.f, td = 0.f, delta = 0.f, dv = 0.f;
            if (in) v = __ldg(head_out + (o + t) * NH);
            if (ok) {
                const float v1 = __ldg(head_out + (o + t + 1) * NH);
                td = __ldg(rew + o + t) + (gamma * (1.f - __ldg(is_fir + o + t + 1))) * v1;
                delta = td - v;
            }
            const float g = affine_suffix_scan(ok ? k : 1.f, delta, carry);
            carry = __shfl_sync(kFull, g, 0);
            if (ok) {
                const float dd = v - td;
                dv = c_val * invN * smooth_l1_grad(dd);
                s_val += (double)smooth_l1(dd);
                gae_out[on + t] = g;
            }
            if (in) {
                float* dh = dhead + (o + t) * NH;
                dh[0] = dv;
                for (int q = 1; q < NH; ++q) dh[q] = 0.f;
            }
        }
    }
    __shared__ double red[kHeadThreads / 32];
    s_val = warp_sum(s_val);
    if (lane == 0) red[threadIdx.x >> 5] = s_val;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int w = 0; w < kHeadThreads / 32; ++w) a += red[w];
        atomicAdd(scalars + 1, a);
    }
}

__device__ __forceinline__ unsigned umax_(unsigned a, unsigned b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned f2key(float f) { unsigned b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// one CTA per timestep column t: thr[t] = k-th largest of gae[:, t]; colstat[t] = max, colstat[n+t] = sum exp((g-max)*inv_eta) over selected.
// Exactly k rows are selected per column, as torch.topk does: rows above the threshold, plus -- when several rows TIE at the
// threshold value -- the lowest-indexed of them up to tie_cut[t] (INT_MAX when every tied row is taken, the usual case).
__global__ void __launch_bounds__(256)
vmpo_select_kernel(const float* __restrict__ gae, int Bg, int n, int k, const double* __restrict__ duals,
                   float* __restrict__ thr, float* __restrict__ colstat, int* __restrict__ tie_cut) {
    B200RL_DYN_SMEM(unsigned, keys);
    __shared__ int hist[256];
    __shared__ unsigned s_prefix;
    __shared__ int s_k;
    __shared__ double red[8];
    const int t = blockIdx.x, tid = threadIdx.x;
    const float inv = 1.0f / ((float)exp(duals[0]) + 1e-7f);         // ratio = top_gae / (eta + 1e-7), v_mpo/learning.py:66
    unsigned kmax = 0;
    for (int b = tid; b < Bg; b += 256) { unsigned key = f2key(gae[(size_t)b * n + t]); keys[b] = key; kmax = umax_(kmax, key); }
    if (tid == 0) { s_prefix = 0; s_k = k; }
    unsigned mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int b = tid; b < Bg; b += 256) {
            const unsigned key = keys[b];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int kk = s_k, bin = 255;
            for (; bin > 0; --bin) { if (hist[bin] >= kk) break; kk -= hist[bin]; }
            s_prefix = prefix | ((unsigned)bin << shift);
            s_k = kk;
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const unsigned tkey = s_prefix;
    // ties at the threshold: s_k of them belong to the top k
    __shared__ int s_ties, s_cut;
    if (tid == 0) { s_ties = 0; s_cut = 0x7fffffff; }
    __syncthreads();
    {
        int mine = 0;
        for (int b = tid; b < Bg; b += 256) mine += keys[b] == tkey;
        if (mine) atomicAdd(&s_ties, mine);
    }
    __syncthreads();
    if (tid == 0 && s_ties > s_k) {                       // rare (duplicate / saturated advantages): serial scan in index order
        int left = s_k, b = 0;
        for (; b < Bg && left > 0; ++b) left -= keys[b] == tkey;
        s_cut = b - 1;
    }
    __syncthreads();
    const int cut = s_cut;
    // column max
    for (int o = 16; o > 0; o >>= 1) kmax = umax_(kmax, __shfl_xor_sync(kFull, kmax, o));
    __shared__ unsigned wmax[8];
    if ((tid & 31) == 0) wmax[tid >> 5] = kmax;
    __syncthreads();
    unsigned cm = wmax[0];
    for (int w = 1; w < 8; ++w) cm = umax_(cm, wmax[w]);
    const float colmax = key2f(cm);
    double s = 0;
    for (int b = tid; b < Bg; b += 256) {
        const unsigned key = keys[b];
        if (key > tkey || (key == tkey && b <= cut)) s += (double)expf((key2f(key) - colmax) * inv);
    }
    s = warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double a = 0;
        for (int w = 0; w < 8; ++w) a += red[w];
        thr[t] = key2f(tkey); colstat[t] = colmax; colstat[n + t] = (float)a; tie_cut[t] = cut;
    }
}

struct VmpoGradArgs {
    const float* head_out; const float* act; const float* blogits; const float* gae; const float* thr; const float* colstat;
    const double* duals; float* dhead; double* scalars; const int* tie_cut;
    int B, T, A, NH, row_offset;
    float c_pol, invN;
};

template <int AT>                    // AT > 0: A == AT at compile time (dist.cuh)
__global__ void __launch_bounds__(kHeadThreads)
vmpo_grad_kernel(VmpoGradArgs a) {
    constexpr int NZ = AT > 0 ? AT : MAXA;
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int T = a.T, n = T - 1, NH = a.NH, A = AT > 0 ? AT : a.A;
    const int nslab = (n + 31) >> 5;
    const float inv = 1.0f / ((float)exp(a.duals[0]) + 1e-7f);
    const float alpha32 = (float)exp(a.duals[1]);
    // global softmax normaliser from the per-column statistics (every CTA recomputes it: n <= a few hundred)
    __shared__ double sh_m, sh_Z;
    __shared__ double red[kHeadThreads / 32];
    {
        float m = -INFINITY;
        for (int t = threadIdx.x; t < n; t += blockDim.x) m = fmaxf(m, a.colstat[t] * inv);
        m = warp_max(m);
        if (lane == 0) red[threadIdx.x >> 5] = (double)m;
        __syncthreads();
        if (threadIdx.x == 0) { double mm = red[0]; for (int w = 1; w < kHeadThreads / 32; ++w) mm = fmax(mm, red[w]); sh_m = mm; }
        __syncthreads();
        const float mf = (float)sh_m;
        double z = 0;
        for (int t = threadIdx.x; t < n; t += blockDim.x) z += (double)a.colstat[n + t] * (double)expf(a.colstat[t] * inv - mf);
        z = warp_sum(z);
        __syncthreads();
        if (lane == 0) red[threadIdx.x >> 5] = z;
        __syncthreads();
        if (threadIdx.x == 0) {
            double zz = 0; for (int w = 0; w < kHeadThreads / 32; ++w) zz += red[w];
            sh_Z = zz;
            if (blockIdx.x == 0) { a.scalars[6] = sh_m; a.scalars[7] = zz; }
        }
        __syncthreads();
    }
    const float mf = (float)sh_m;
    const float invZ = (float)(1.0 / sh_Z);
    double s_pol = 0, s_kl = 0, s_cnt = 0;
    for (int row = wid; row < a.B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        for (int s = 0; s < nslab; ++s) {
            const int t = s * 32 + lane;
            if (t >= n) continue;
            const float* ho = a.head_out + (o + t) * NH;
            float z[MAXA];
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) z[q] = __ldg(ho + 1 + q);
            const float actv = __ldg(a.act + o + t);
            Dist d;
            dist_forward<false, AT>(z, A, actv, d);
            // behaviour policy: Categorical(softmax(behaviour_logits)) restated (compute_loss.py:74-77)
            float bl[MAXA], pp[MAXA], bm = -INFINITY, bs = 0.f, bS = 0.f;
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) { bl[q] = __ldg(a.blogits + (o + t) * A + q); bm = fmaxf(bm, bl[q]); }
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) { pp[q] = expf(bl[q] - bm); bs += pp[q]; }
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) { pp[q] = pp[q] / bs; bS += pp[q]; }
            // current policy as the reference sees it: softmax of the normalised log-probs it returns as "logits"
            float lm = -INFINITY, ls = 0.f, q2[MAXA], qS = 0.f;
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) lm = fmaxf(lm, d.l[q]);
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) { q2[q] = expf(d.l[q] - lm); ls += q2[q]; }
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) { q2[q] = q2[q] / ls; qS += q2[q]; }
            float kl = 0.f;
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) {
                const float p = pp[q] / bS, qq = q2[q] / qS;
                float term = p * (logf(fminf(fmaxf(p, kEps), 1.f - kEps)) - logf(fminf(fmaxf(qq, kEps), 1.f - kEps)));
                if (qq == 0.f) term = INFINITY;
                if (p == 0.f) term = 0.f;
                kl += term;
            }
            s_kl += (double)kl;
            const float g = __ldg(a.gae + on + t);
            float psi = 0.f;
            const float th = __ldg(a.thr + t);
            if (g > th || (g == th && a.row_offset + row <= __ldg(a.tie_cut + t))) {
                psi = expf(g * inv - mf) * invZ;
                s_pol += (double)(-psi * d.logp);
                s_cnt += 1.0;
            }
            float dz[MAXA];
#pragma unroll
            for (int q = 0; q < NZ; ++q) dz[q] = 0.f;
            dist_backward<false, AT>(d, A, actv, -a.c_pol * psi, 0.f, dz);
            const float kc = alpha32 * a.invN;                     // d(L_alpha)/dKL = alpha.detach() / N
            float* dh = a.dhead + (o + t) * NH;
#pragma unroll
            for (int q = 0; q < NZ; ++q) if (q < A) dh[1 + q] = dz[q] + kc * (d.p[q] / d.S - pp[q] / bS);
        }
    }
    s_pol = warp_sum(s_pol); s_kl = warp_sum(s_kl); s_cnt = warp_sum(s_cnt);
    __shared__ double r3[3][kHeadThreads / 32];
    if (lane == 0) { r3[0][threadIdx.x >> 5] = s_pol; r3[1][threadIdx.x >> 5] = s_kl; r3[2][threadIdx.x >> 5] = s_cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0, y = 0, c = 0;
        for (int w = 0; w < kHeadThreads / 32; ++w) { x += r3[0][w]; y += r3[1][w]; c += r3[2][w]; }
        atomicAdd(a.scalars + 0, x); atomicAdd(a.scalars + 2, y); atomicAdd(a.scalars + 3, c);
    }
}

// duals = {log_eta, log_alpha}, dual_sq = their RMSprop second moments; losses_out[6] =
// {L_pi, L_V, L_eta, L_alpha, total, KL mean}
__global__ void vmpo_dual_kernel(double* duals, double* dual_sq, const double* scalars, double* losses_out,
                                 double coef_eta, double coef_alpha, double lr, double rms_alpha, double rms_eps,
                                 double N, double n_selected, double c_pol, double c_val) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double eta = exp(duals[0]), alpha = exp(duals[1]);
    const double lme = scalars[6] + log(scalars[7]) - log(n_selected);           // log(mean(exp(ratio)))
    const double L_eta = eta * coef_eta + eta * lme;
    const double klm = scalars[2] / N;
    const double L_alpha = alpha * coef_alpha;                                    // alpha*(eps - KL) + alpha*KL, mean
    const double g_eta = eta * (coef_eta + lme);
    const double g_alpha = alpha * (coef_alpha - klm);
    const double L_pi = scalars[0], L_v = scalars[1] / N;
    losses_out[0] = L_pi; losses_out[1] = L_v; losses_out[2] = L_eta; losses_out[3] = L_alpha;
    losses_out[4] = c_pol * L_pi + c_val * L_v + L_eta + L_alpha; losses_out[5] = klm;
    const double g[2] = {g_eta, g_alpha};
    for (int i = 0; i < 2; ++i) {                                                  // RMSprop, same group options as the model (learner.py:331-338)
        dual_sq[i] = rms_alpha * dual_sq[i] + (1.0 - rms_alpha) * g[i] * g[i];
        duals[i] -= lr * g[i] / (sqrt(dual_sq[i]) + rms_eps);
    }
}

static inline int vm_grid(int rows) {
    long long ctas = ((long long)rows + kHeadThreads / 32 - 1) / (kHeadThreads / 32);
    long long cap = 148LL * 4;
    return (int)(ctas < 1 ? 1 : (ctas > cap ? cap : ctas));
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_vmpo_gae(const float* head_out, const float* rew, const float* is_fir,
                               float* gae, float* dhead, double* scalars, int B, int T, int A,
                               float gamma, float lambda_, float c_value, double N_global, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2 && A >= 1 && A <= MAXA, "vmpo_gae: bad shape B=%d T=%d A=%d", B, T, A);
    B200RL_LAUNCH(vmpo_gae_kernel, vm_grid(B), kHeadThreads, 0, stream, head_out, rew, is_fir, gae, dhead, scalars, B, T, 1 + A,
                  gamma, lambda_, c_value, (float)(1.0 / N_global));
    return check_launch("vmpo_gae");
}

extern "C" int b200rl_vmpo_select(const float* gae_all, int Bg, int n, int k, const double* duals,
                                  float* thr, float* colstat, int* tie_cut, void* stream) {
    B200RL_REQUIRE(Bg >= 1 && n >= 1 && k >= 1 && k <= Bg, "vmpo_select: bad shape Bg=%d n=%d k=%d", Bg, n, k);
    const size_t smem = sizeof(unsigned) * (size_t)Bg;
    B200RL_REQUIRE(smem <= 200 * 1024, "vmpo_select: batch of %d sequences does not fit the shared-memory key buffer", Bg);
#ifndef B200RL_EMUL
    cudaError_t e = cudaFuncSetAttribute(vmpo_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error("vmpo_select: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
#endif
    B200RL_LAUNCH(vmpo_select_kernel, n, 256, smem, stream, gae_all, Bg, n, k, duals, thr, colstat, tie_cut);
    return check_launch("vmpo_select");
}

extern "C" int b200rl_vmpo_grad(const float* head_out, const float* act, const float* behav_logits, const float* gae,
                                const float* thr, const float* colstat, const int* tie_cut, const double* duals,
                                float* dhead, double* scalars, int B, int T, int A, int row_offset,
                                float c_policy, double N_global, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2 && A >= 1 && A <= MAXA, "vmpo_grad: bad shape B=%d T=%d A=%d", B, T, A);
    VmpoGradArgs a{head_out, act, behav_logits, gae, thr, colstat, duals, dhead, scalars, tie_cut, B, T, A, 1 + A, row_offset, c_policy, (float)(1.0 / N_global)};
    auto kern = A == 2 ? vmpo_grad_kernel<2> : A == 3 ? vmpo_grad_kernel<3> : A == 4 ? vmpo_grad_kernel<4> : vmpo_grad_kernel<0>;
    B200RL_LAUNCH(kern, vm_grid(B), kHeadThreads, 0, stream, a);
    return check_launch("vmpo_grad");
}

extern "C" int b200rl_vmpo_dual_step(double* duals, double* dual_sq, const double* scalars, double* losses_out,
                                     double coef_eta, double coef_alpha, double lr, double rms_alpha, double rms_eps,
                                     double N_global, double n_selected, double c_policy, double c_value, void* stream) {
    B200RL_LAUNCH(vmpo_dual_kernel, 1, 32, 0, stream, duals, dual_sq, scalars, losses_out, coef_eta, coef_alpha, lr, rms_alpha, rms_eps,
                  N_global, n_selected, c_policy, c_value);
    return check_launch("vmpo_dual_step");
}
