// Distribution math shared by the loss-head kernels (PPO / IMPALA / V-MPO):
// Categorical(probs=softmax(z)) exactly as torch builds it (networks/models.py:58-61):
// q = p / sum(p), logits = log(clamp(q, eps, 1-eps)), entropy = -sum q*logits; and
// Normal(tanh(z_mu), softplus(z_std)) (networks/models.py:114-118); with the analytic
// backward to the head pre-activations.
#pragma once
#include "common.cuh"

namespace b200rl {

constexpr int kHeadThreads = 256;
constexpr int MAXA = B200RL_MAX_ACT;
constexpr float kEps = 1.1920929e-07f;          // torch.finfo(float32).eps
constexpr float kLogSqrt2Pi = 0.91893853320467274f;

__device__ __forceinline__ void atomic_min_double(double* p, double v) {
#ifdef B200RL_EMUL
    std::lock_guard<std::mutex> g(emul::atomic_mu());
    if (v < *p) *p = v;
#else
    unsigned long long* a = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
#endif
}
__device__ __forceinline__ void atomic_max_double(double* p, double v) {
#ifdef B200RL_EMUL
    std::lock_guard<std::mutex> g(emul::atomic_mu());
    if (v > *p) *p = v;
#else
    unsigned long long* a = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
#endif
}

// ---- distribution forward/backward on one (b,t) -----------------------------------------
struct Dist {
    float p[MAXA];      // discrete: softmax probs ; continuous: p[0] = mu, p[1] = sd, p[2] = z_sd
    float l[MAXA];      // discrete: normalised, clamped log-probs ("logits" the reference returns)
    float S;            // discrete: sum of softmax probs (renormaliser)
    float logp, ent;
};

// AT > 0: the number of actions is the compile-time constant AT (== A): the loops below unroll to exactly AT iterations instead
// of MAXA predicated ones (with A = 2 the generic form executes four times the exp / log / divide sequences it needs:
// ncu r2g, 775 instructions per element in ppo_head)
template <bool CONT, int AT = 0>
__device__ __forceinline__ void dist_forward(const float* z, int A_, float act, Dist& d) {
    constexpr int NA = AT > 0 ? AT : MAXA;
    const int A = AT > 0 ? AT : A_;
    if (!CONT) {
        float mx = -INFINITY;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) mx = fmaxf(mx, z[a]);
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) { d.p[a] = expf(z[a] - mx); s += d.p[a]; }
        float S = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) { d.p[a] = d.p[a] / s; S += d.p[a]; }
        d.S = S;
        const int ai = (int)act;
        float ent = 0.f, lp = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) {
            float q = d.p[a] / S;
            float l = logf(fminf(fmaxf(q, kEps), 1.f - kEps));
            d.l[a] = l;
            ent -= q * l;
            if (a == ai) lp = l;
        }
        d.logp = lp; d.ent = ent;
    } else {
        const float mu = tanhf(z[0]);
        const float zs = z[1];
        const float sd = zs > 20.f ? zs : log1pf(expf(zs));      // F.softplus(beta=1, threshold=20)
        d.p[0] = mu; d.p[1] = sd; d.p[2] = zs;
        const float df = act - mu;
        d.logp = -(df * df) / (2.f * sd * sd) - logf(sd) - kLogSqrt2Pi;
        d.ent = 0.5f + kLogSqrt2Pi + logf(sd);
    }
}

// dz[] <- d(loss)/dz given d(loss)/dlogp and d(loss)/dentropy
template <bool CONT, int AT = 0>
__device__ __forceinline__ void dist_backward(const Dist& d, int A_, float act, float dlogp, float dent, float* dz) {
    constexpr int NA = AT > 0 ? AT : MAXA;
    const int A = AT > 0 ? AT : A_;
    if (!CONT) {
        const int ai = (int)act;
        float gq[MAXA];
        float dot = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) {
            const float q = d.p[a] / d.S;
            const float inr = (q >= kEps && q <= 1.f - kEps) ? 1.f : 0.f;
            float g = dent * (-(d.l[a] + inr));
            if (a == ai) g += dlogp * inr / q;
            gq[a] = g;
            dot += g * q;
        }
        float dp[MAXA];
        float dot2 = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) { dp[a] = (gq[a] - dot) / d.S; dot2 += dp[a] * d.p[a]; }
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a < A) dz[a] = d.p[a] * (dp[a] - dot2);
    } else {
        const float mu = d.p[0], sd = d.p[1], zs = d.p[2];
        const float df = act - mu;
        const float dmu = dlogp * df / (sd * sd);
        const float dsd = dlogp * (df * df / (sd * sd * sd) - 1.f / sd) + dent / sd;
        dz[0] = dmu * (1.f - mu * mu);
        dz[1] = dsd * (zs > 20.f ? 1.f : 1.f / (1.f + expf(-zs)));
    }
}

}  // namespace b200rl
