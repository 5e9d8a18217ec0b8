// Loss heads: distribution math + per-timestep recurrence + loss terms + the analytic
// gradient w.r.t. the head pre-activations, one pass over (B, T) (sm_100a).
//
//   PPO     agents/learner_module/ppo/learning.py:43-81     (clipped surrogate, GAE)
//   IMPALA  agents/learner_module/impala/learning.py:43-69  (V-trace policy gradient)
//   distributions: networks/models.py:58-61,86-87 (Categorical(probs=softmax)),
//                  networks/models.py:114-118 (Normal(tanh(mu), softplus(std)))
//
// One warp owns one sequence; lanes run along T in slabs of 32 timesteps processed from
// the end of the row (the recurrences run backwards in time), using the warp affine
// suffix scan of scan.cuh.  Per (b,t) the kernel reads NH + 4 floats and writes NH floats:
// HBM-bound and tiny next to the LSTM kernels.  Loss sums are accumulated in fp64
// (warp shuffle -> shared -> one atomic per CTA) so the result does not depend on the
// launch geometry beyond fp64 rounding.
//
// Gradients follow autograd of the reference expressions exactly:
//   torch.min tie -> grad/2 each side; clamp passes gradient on the closed interval;
//   smooth_l1 (beta=1) switches at |d| < 1; Categorical(probs=p): q = p/sum(p),
//   logits = log(clamp(q, eps, 1-eps)), entropy = -sum q*logits.
#include "common.cuh"
#include "scan.cuh"
#include "dist.cuh"

namespace b200rl {

// scalars: [0] sum policy  [1] sum value  [2] sum entropy  [3] sum ratio  [4] min ratio  [5] max ratio
__device__ __forceinline__ void flush_scalars(double s_pol, double s_val, double s_ent, double s_rat,
                                              float rmin, float rmax, double* scalars) {
    __shared__ double red[6][kHeadThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_pol = warp_sum(s_pol); s_val = warp_sum(s_val); s_ent = warp_sum(s_ent); s_rat = warp_sum(s_rat);
    rmin = warp_min(rmin); rmax = warp_max(rmax);
    if (lane == 0) {
        red[0][warp] = s_pol; red[1][warp] = s_val; red[2][warp] = s_ent; red[3][warp] = s_rat;
        red[4][warp] = (double)rmin; red[5][warp] = (double)rmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, d = 0, mn = INFINITY, mx = -INFINITY;
        for (int w = 0; w < kHeadThreads / 32; ++w) {
            a += red[0][w]; b += red[1][w]; c += red[2][w]; d += red[3][w];
            mn = fmin(mn, red[4][w]); mx = fmax(mx, red[5][w]);
        }
        atomicAdd(scalars + 0, a); atomicAdd(scalars + 1, b); atomicAdd(scalars + 2, c); atomicAdd(scalars + 3, d);
        atomic_min_double(scalars + 4, mn); atomic_max_double(scalars + 5, mx);
    }
}

struct HeadArgs {
    const float* head_out; const float* act; const float* rew; const float* is_fir; const float* blogp;
    float* dhead; double* scalars;
    float* o_logits; float* o_logp; float* o_ent; float* o_a; float* o_b; float* o_c;   // optional outputs
    int B, T, A, NH;
    float gamma, lam, eps_clip, rho_bar, c_bar, c_pol, c_val, c_ent;
};

// ------------------------------------------------------------------------------- PPO
template <bool CONT, int AT>          // AT > 0: A == AT (discrete) at compile time
__global__ void __launch_bounds__(kHeadThreads)
ppo_head_kernel(HeadArgs a) {
    constexpr int NZ = CONT ? 2 : (AT > 0 ? AT : MAXA);     // head pre-activations after the value
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int T = a.T, n = T - 1, NH = a.NH, A = a.A;
    const int nslab = (T + 31) >> 5;
    const float k = a.gamma * a.lam;
    const float invN = 1.0f / ((float)a.B * (float)n);
    double s_pol = 0, s_val = 0, s_ent = 0, s_rat = 0;
    float rmin = INFINITY, rmax = -INFINITY;
    for (int row = wid; row < a.B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        float carry = 0.f;
        for (int s = nslab - 1; s >= 0; --s) {
            const int t = s * 32 + lane;
            const bool in = t < T, ok = t < n;
            float z[MAXA], v = 0.f, v1 = 0.f, actv = 0.f, td = 0.f, delta = 0.f;
            Dist d;
            d.logp = 0.f; d.ent = 0.f;
            if (in) {
                const float* ho = a.head_out + (o + t) * NH;
                v = __ldg(ho);
#pragma unroll
                for (int q = 0; q < NZ; ++q) if (q < NH - 1) z[q] = __ldg(ho + 1 + q);
                actv = __ldg(a.act + o + t);
                dist_forward<CONT, AT>(z, A, actv, d);
                if (a.o_logp) a.o_logp[o + t] = d.logp;
                if (a.o_ent) a.o_ent[o + t] = d.ent;
                if (!CONT && a.o_logits)
                    for (int q = 0; q < NZ; ++q) if (q < A) a.o_logits[(o + t) * A + q] = d.l[q];
            }
            if (ok) {
                v1 = __ldg(a.head_out + (o + t + 1) * NH);
                td = __ldg(a.rew + o + t) + (a.gamma * (1.f - __ldg(a.is_fir + o + t + 1))) * v1;
                delta = td - v;
            }
            const float gae = affine_suffix_scan(ok ? k : 1.f, delta, carry);
            carry = __shfl_sync(kFull, gae, 0);
            float dz[MAXA], dv = 0.f;
#pragma unroll
            for (int q = 0; q < NZ; ++q) dz[q] = 0.f;
            if (ok) {
                const float ratio = expf(d.logp - __ldg(a.blogp + o + t));
                const float lo = 1.f - a.eps_clip, hi = 1.f + a.eps_clip;
                const float s1 = ratio * gae;
                const float s2 = fminf(fmaxf(ratio, lo), hi) * gae;
                const float inclamp = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
                const float g1 = s1 < s2 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
                const float g2 = s2 < s1 ? 1.f : (s1 == s2 ? 0.5f : 0.f);
                const float dratio = -(g1 * gae + g2 * gae * inclamp) * invN;
                const float dlogp = a.c_pol * dratio * ratio;
                const float dd = v - td;
                dv = a.c_val * invN * smooth_l1_grad(dd);
                dist_backward<CONT, AT>(d, A, actv, dlogp, -a.c_ent * invN, dz);
                s_pol += (double)(-fminf(s1, s2));
                s_val += (double)smooth_l1(dd);
                s_ent += (double)d.ent;
                s_rat += (double)ratio;
                rmin = fminf(rmin, ratio); rmax = fmaxf(rmax, ratio);
                if (a.o_a) a.o_a[on + t] = td;
                if (a.o_b) a.o_b[on + t] = gae;
                if (a.o_c) a.o_c[on + t] = ratio;
            }
            if (in) {
                float* dh = a.dhead + (o + t) * NH;
                dh[0] = dv;
#pragma unroll
                for (int q = 0; q < NZ; ++q) if (q < NH - 1) dh[1 + q] = dz[q];
            }
        }
    }
    flush_scalars(s_pol, s_val, s_ent, s_rat, rmin, rmax, a.scalars);
}

// ---------------------------------------------------------------------------- IMPALA
template <int AT>
__global__ void __launch_bounds__(kHeadThreads)
impala_head_kernel(HeadArgs a) {
    constexpr int NZ = AT > 0 ? AT : MAXA;
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const int T = a.T, n = T - 1, NH = a.NH, A = a.A;
    const int nslab = (T + 31) >> 5;
    const float invN = 1.0f / ((float)a.B * (float)n);
    double s_pol = 0, s_val = 0, s_ent = 0, s_rat = 0;
    float rmin = INFINITY, rmax = -INFINITY;
    for (int row = wid; row < a.B; row += nw) {
        const size_t o = (size_t)row * T, on = (size_t)row * n;
        float carry = 0.f;
        for (int s = nslab - 1; s >= 0; --s) {
            const int t = s * 32 + lane;
            const bool in = t < T, ok = t < n;
            float z[MAXA], v = 0.f, v1 = 0.f, actv = 0.f, rc = 0.f, cc = 1.f, disc1 = 0.f, rt = 0.f, delta = 0.f;
            Dist d;
            d.logp = 0.f; d.ent = 0.f;
            if (in) {
                const float* ho = a.head_out + (o + t) * NH;
                v = __ldg(ho);
#pragma unroll
                for (int q = 0; q < NZ; ++q) if (q < NH - 1) z[q] = __ldg(ho + 1 + q);
                actv = __ldg(a.act + o + t);
                dist_forward<false, AT>(z, A, actv, d);
                if (a.o_logp) a.o_logp[o + t] = d.logp;
                if (a.o_ent) a.o_ent[o + t] = d.ent;
                if (a.o_logits)
                    for (int q = 0; q < NZ; ++q) if (q < A) a.o_logits[(o + t) * A + q] = d.l[q];
            }
            if (ok) {
                const float rho = expf(d.logp - __ldg(a.blogp + o + t));
                rc = fminf(fmaxf(rho, 0.1f), a.rho_bar);
                cc = fminf(rho, a.c_bar);
                disc1 = a.gamma * (1.f - __ldg(a.is_fir + o + t + 1));
                rt = __ldg(a.rew + o + t);
                v1 = __ldg(a.head_out + (o + t + 1) * NH);
                delta = rc * (rt + disc1 * v1 - v);
            }
            float acc = affine_suffix_scan(ok ? cc * disc1 : 1.f, delta, carry);
            if (t == n) acc = 0.f;
            float acc_next = __shfl_down_sync(kFull, acc, 1);
            if (lane == 31) acc_next = carry;
            carry = __shfl_sync(kFull, acc, 0);
            const float vs = v + acc;
            float dz[MAXA], dv = 0.f;
#pragma unroll
            for (int q = 0; q < NZ; ++q) dz[q] = 0.f;
            if (ok) {
                const float vs1 = v1 + ((t + 1 == n) ? 0.f : acc_next);
                const float adv = rc * (rt + disc1 * vs1 - v);
                const float dlogp = -a.c_pol * adv * invN;
                const float dd = v - vs;
                dv = a.c_val * invN * smooth_l1_grad(dd);
                dist_backward<false, AT>(d, A, actv, dlogp, -a.c_ent * invN, dz);
                s_pol += (double)(-d.logp * adv);
                s_val += (double)smooth_l1(dd);
                s_ent += (double)d.ent;
                s_rat += (double)rc;
                rmin = fminf(rmin, rc); rmax = fmaxf(rmax, rc);
                if (a.o_a) a.o_a[on + t] = rc;
                if (a.o_b) a.o_b[on + t] = adv;
            }
            if (in) {
                if (a.o_c) a.o_c[o + t] = vs;
                float* dh = a.dhead + (o + t) * NH;
                dh[0] = dv;
#pragma unroll
                for (int q = 0; q < NZ; ++q) if (q < NH - 1) dh[1 + q] = dz[q];
            }
        }
    }
    flush_scalars(s_pol, s_val, s_ent, s_rat, rmin, rmax, a.scalars);
}

static inline int head_grid(int rows) {
    long long ctas = ((long long)rows + kHeadThreads / 32 - 1) / (kHeadThreads / 32);
    long long cap = 148LL * 4;
    return (int)(ctas < 1 ? 1 : (ctas > cap ? cap : ctas));
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_ppo_head(const float* head_out, const float* act, const float* rew,
                               const float* is_fir, const float* behav_logp,
                               float* dhead, double* scalars,
                               float* logits_norm, float* logp, float* entropy,
                               float* td_target, float* gae, float* ratio,
                               int B, int T, int A, int continuous,
                               float gamma, float lambda_, float eps_clip,
                               float c_policy, float c_value, float c_entropy, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2, "ppo_head: need B >= 1, T >= 2 (got B=%d T=%d)", B, T);
    B200RL_REQUIRE(A >= 1 && A <= MAXA, "ppo_head: A=%d out of range [1,%d]", A, MAXA);
    B200RL_REQUIRE(!continuous || A == 1, "ppo_head: continuous PPO supports act_dim 1 (the reference stores 1-wide actions)");
    HeadArgs a{head_out, act, rew, is_fir, behav_logp, dhead, scalars, logits_norm, logp, entropy, td_target, gae, ratio,
               B, T, A, continuous ? 1 + 2 * A : 1 + A, gamma, lambda_, eps_clip, 0.f, 0.f, c_policy, c_value, c_entropy};
    auto kern = continuous ? ppo_head_kernel<true, 0>
              : A == 2 ? ppo_head_kernel<false, 2> : A == 3 ? ppo_head_kernel<false, 3> : A == 4 ? ppo_head_kernel<false, 4>
              : ppo_head_kernel<false, 0>;
    B200RL_LAUNCH(kern, head_grid(B), kHeadThreads, 0, stream, a);
    return check_launch("ppo_head");
}

extern "C" int b200rl_impala_head(const float* head_out, const float* act, const float* rew,
                                  const float* is_fir, const float* behav_logp,
                                  float* dhead, double* scalars,
                                  float* logits_norm, float* logp, float* entropy,
                                  float* rho, float* adv, float* vs,
                                  int B, int T, int A,
                                  float gamma, float rho_bar, float c_bar,
                                  float c_policy, float c_value, float c_entropy, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2, "impala_head: need B >= 1, T >= 2 (got B=%d T=%d)", B, T);
    B200RL_REQUIRE(A >= 1 && A <= MAXA, "impala_head: A=%d out of range [1,%d]", A, MAXA);
    HeadArgs a{head_out, act, rew, is_fir, behav_logp, dhead, scalars, logits_norm, logp, entropy, rho, adv, vs,
               B, T, A, 1 + A, gamma, 0.f, 0.f, rho_bar, c_bar, c_policy, c_value, c_entropy};
    auto kern = A == 2 ? impala_head_kernel<2> : A == 3 ? impala_head_kernel<3> : A == 4 ? impala_head_kernel<4> : impala_head_kernel<0>;
    B200RL_LAUNCH(kern, head_grid(B), kHeadThreads, 0, stream, a);
    return check_launch("impala_head");
}
