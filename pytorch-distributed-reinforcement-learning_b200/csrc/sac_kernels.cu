// SAC / SAC-Continuous loss heads (sm_100a): element-wise kernels over (B, T); the losses have
// no recurrence, the TD target only looks one step ahead.
//
// Discrete SAC            agents/learner_module/sac/learning.py:36-143
//   sac_actor_head   L_pi = mean_{b,t<T-1} sum_a p (alpha*logp - min(q1,q2)),  logp = log(p + 1e-8[p==0])
//                    -> d/dz of the actor, sum of p*logp for the temperature loss
//   sac_critic_head  y = r + (1-f[t+1])*gamma * sum_a p'(min(q1t,q2t) - alpha*logp')[t+1]  (updated actor p')
//                    L_Q = mean smooth_l1(q1[a]-y) + mean smooth_l1(q2[a]-y) -> dq1, dq2 (one-hot in a)
// Continuous SAC          agents/learner_module/sac_continuous/learning.py:36-131, networks/models.py:199-231
//   sac_cont_sample      x = mu + sigma*eps, a = tanh(x), logp = N(x;mu,sigma).log_prob - log(1-a^2+1e-7)
//   sac_cont_minq_grad   d(-min(q1,q2))/dq_i (torch.min tie -> 1/2 each) for the policy step
//   sac_cont_actor_grad  d/d(mu, raw log-std) from alpha*logp and from dQ/da (critic backward to its input)
//   sac_cont_critic_head y = r + (1-f[t+1])*gamma * sum_d (min(q1t,q2t) - alpha*logp'_d)[t+1]  (the reference's
//                        `.sum(-1)` counts min q once per action dimension -- reproduced), smooth-L1 twin loss
//   sac_alpha_step       temperature loss, its gradient and the fp64 Adam step of log_alpha (learner.py:363-367)
// scalars (8 doubles, accumulated): [0] sum policy term, [1] sum value term, [2] sum entropy term.
#include "common.cuh"
#include "dist.cuh"

namespace b200rl {

__device__ __forceinline__ float alpha32_of(const double* log_alpha) { return (float)exp(*log_alpha); }

__device__ __forceinline__ void block_add3(double a, double b, double c, double* scalars, int ia, int ib, int ic) {
    __shared__ double red[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0, y = 0, z = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { x += red[0][w]; y += red[1][w]; z += red[2][w]; }
        if (ia >= 0) atomicAdd(scalars + ia, x);
        if (ib >= 0) atomicAdd(scalars + ib, y);
        if (ic >= 0) atomicAdd(scalars + ic, z);
    }
}

__device__ __forceinline__ void softmax_sac(const float* z, int A, float* p, float* lp) {
    float mx = -INFINITY, s = 0.f;
#pragma unroll
    for (int a = 0; a < MAXA; ++a) if (a < A) mx = fmaxf(mx, z[a]);
#pragma unroll
    for (int a = 0; a < MAXA; ++a) if (a < A) { p[a] = expf(z[a] - mx); s += p[a]; }
#pragma unroll
    for (int a = 0; a < MAXA; ++a) if (a < A) { p[a] = p[a] / s; lp[a] = logf(p[a] + (p[a] == 0.f ? 1e-8f : 0.f)); }
}

// ------------------------------------------------------------------------- discrete SAC
__global__ void __launch_bounds__(256)
sac_actor_head_kernel(const float* __restrict__ z, const float* __restrict__ q1, const float* __restrict__ q2,
                      const double* __restrict__ log_alpha, float* __restrict__ dz, double* __restrict__ scalars,
                      long long rows, int T, int A, float invN) {
    const float alpha = alpha32_of(log_alpha);
    double s_pol = 0, s_ent = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        float zz[MAXA], p[MAXA], lp[MAXA], g[MAXA];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) if (a < A) zz[a] = __ldg(z + i * A + a);
        softmax_sac(zz, A, p, lp);
        float dot = 0.f, el = 0.f, en = 0.f;
#pragma unroll
        for (int a = 0; a < MAXA; ++a) if (a < A) {
            const float m = fminf(__ldg(q1 + i * A + a), __ldg(q2 + i * A + a));
            el += p[a] * (alpha * lp[a] - m);
            en += p[a] * lp[a];
            g[a] = alpha * lp[a] - m + (p[a] > 0.f ? alpha : 0.f);
            dot += p[a] * g[a];
        }
        const bool ok = t < T - 1;
        if (ok) { s_pol += (double)el; s_ent += (double)en; }
#pragma unroll
        for (int a = 0; a < MAXA; ++a) if (a < A) dz[i * A + a] = ok ? invN * p[a] * (g[a] - dot) : 0.f;
    }
    block_add3(s_pol, s_ent, 0.0, scalars, 0, 2, -1);
}

__global__ void __launch_bounds__(256)
sac_critic_head_kernel(const float* __restrict__ z, const float* __restrict__ q1, const float* __restrict__ q2,
                       const float* __restrict__ tq1, const float* __restrict__ tq2,
                       const float* __restrict__ act, const float* __restrict__ rew, const float* __restrict__ is_fir,
                       const double* __restrict__ log_alpha, float* __restrict__ dq1, float* __restrict__ dq2,
                       float* __restrict__ td_out, double* __restrict__ scalars,
                       long long rows, int T, int A, float gamma, float invN) {
    const float alpha = alpha32_of(log_alpha);
    double s_val = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
#pragma unroll
        for (int a = 0; a < MAXA; ++a) if (a < A) { dq1[i * A + a] = 0.f; dq2[i * A + a] = 0.f; }
        if (t >= T - 1) continue;
        const long long j = i + 1;                                  // (b, t+1)
        float zz[MAXA], p[MAXA], lp[MAXA];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) if (a < A) zz[a] = __ldg(z + j * A + a);
        softmax_sac(zz, A, p, lp);
        float soft = 0.f;
#pragma unroll
        for (int a = 0; a < MAXA; ++a) if (a < A)
            soft += p[a] * (fminf(__ldg(tq1 + j * A + a), __ldg(tq2 + j * A + a)) - alpha * lp[a]);
        const float y = __ldg(rew + i) + (1.f - __ldg(is_fir + j)) * gamma * soft;
        const int ai = (int)__ldg(act + i);
        const float d1 = __ldg(q1 + i * A + ai) - y, d2 = __ldg(q2 + i * A + ai) - y;
        dq1[i * A + ai] = invN * smooth_l1_grad(d1);
        dq2[i * A + ai] = invN * smooth_l1_grad(d2);
        s_val += (double)smooth_l1(d1) + (double)smooth_l1(d2);
        if (td_out) td_out[(i / T) * (T - 1) + t] = y;
    }
    block_add3(s_val, 0.0, 0.0, scalars, 1, -1, -1);
}

// ----------------------------------------------------------------------- continuous SAC
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so element i of draw n is a pure function of (seed, n, i) -- no
// generator state per thread, reproducible for any grid, and safe to replay from a CUDA graph (the draw number lives in
// device memory).  Replaces the two `torch.randn` of `dist.rsample()` (networks/models.py:213, sac_continuous/learning.py:47,68).
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// one N(0,1) per element (Box-Muller on two of the four words)
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long draw, unsigned long long i) {
    unsigned r[4];
    philox4x32_10((unsigned)i, (unsigned)(i >> 32), (unsigned)draw, (unsigned)(draw >> 32), (unsigned)seed, (unsigned)(seed >> 32), r);
    const float u0 = ((float)(r[0] >> 8) + 0.5f) * (1.f / 16777216.f);          // (0, 1)
    const float u1 = ((float)(r[1] >> 8) + 0.5f) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u0)) * cosf(6.283185307179586f * u1);
}

// a = tanh(mu + sd * eps), log pi(a) ; `eps_in` given: the reference's arithmetic on injected noise (parity tests); null: noise of
// draw number rng[1] under seed rng[0], written to `eps_out` for the actor gradient, and the LAST CTA to finish advances rng[1]
// (rng[2] is its ticket counter).  `a_out` has row stride `a_ld` floats: the action lands directly in the critic's [obs | act] input.
__global__ void __launch_bounds__(256)
sac_cont_sample_kernel(const float* __restrict__ ho, const float* __restrict__ eps_in, float* __restrict__ eps_out,
                       unsigned long long* __restrict__ rng, float* __restrict__ a_out, int a_ld,
                       float* __restrict__ logp_out, long long rows, int A) {
    unsigned long long seed = 0, draw = 0;
    if (eps_in == nullptr) { seed = rng[0]; draw = rng[1]; }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * A; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / A; const int d = (int)(i % A);
        const float mu = __ldg(ho + r * 2 * A + d);
        const float ls = fminf(fmaxf(__ldg(ho + r * 2 * A + A + d), -20.f), 2.f);
        const float sd = expf(ls);
        float e;
        if (eps_in != nullptr) e = __ldg(eps_in + i);
        else { e = philox_normal(seed, draw, (unsigned long long)i); if (eps_out != nullptr) eps_out[i] = e; }
        const float x = mu + sd * e;
        const float a = tanhf(x);
        const float df = x - mu;
        float lp = -(df * df) / (2.f * sd * sd) - logf(sd) - kLogSqrt2Pi;
        lp -= logf(1.f - a * a + 1e-7f);
        a_out[r * a_ld + d] = a; logp_out[i] = lp;
    }
    if (eps_in == nullptr) {
        __syncthreads();                                      // every thread of the CTA has read rng[1]
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned* ticket = reinterpret_cast<unsigned*>(rng + 2);
            if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0u; rng[1] = draw + 1; __threadfence(); }
        }
    }
}

// the three [obs | action] inputs of the critics in one pass: u1 = [obs | .], u2 = [obs | .] (actions written by the two samples), u3 = [obs | act]
__global__ void __launch_bounds__(256)
sac_cont_pack_kernel(const float* __restrict__ obs, const float* __restrict__ act, float* __restrict__ u1, float* __restrict__ u2,
                     float* __restrict__ u3, long long rows, int f, int A) {
    const int Fin = f + A;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * Fin; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / Fin; const int c = (int)(i % Fin);
        if (c < f) { const float v = __ldg(obs + r * f + c); u1[i] = v; u2[i] = v; u3[i] = v; }
        else u3[i] = __ldg(act + r * A + (c - f));
    }
}

__global__ void __launch_bounds__(256)
sac_cont_minq_grad_kernel(const float* __restrict__ q1, const float* __restrict__ q2, float* __restrict__ dq1,
                          float* __restrict__ dq2, double* __restrict__ scalars, long long rows, int T, float invN, float A_f) {
    double s_minq = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const float a = __ldg(q1 + i), b = __ldg(q2 + i);
        const bool ok = t < T - 1;
        const float w1 = a < b ? 1.f : (a == b ? 0.5f : 0.f);
        dq1[i] = ok ? -invN * w1 : 0.f;
        dq2[i] = ok ? -invN * (1.f - w1) : 0.f;
        if (ok) s_minq += (double)fminf(a, b);
    }
    block_add3(-s_minq * (double)A_f, 0.0, 0.0, scalars, 0, -1, -1);     // the -min q part of sum_{b,t,d} (alpha*logp - min q)
}

__global__ void __launch_bounds__(256)
sac_cont_actor_grad_kernel(const float* __restrict__ ho, const float* __restrict__ eps, const float* __restrict__ a_in, int a_ld,
                           const float* __restrict__ logp, const float* __restrict__ du1, const float* __restrict__ du2,
                           const double* __restrict__ log_alpha, float* __restrict__ dho, double* __restrict__ scalars,
                           long long rows, int T, int A, int f, float invNA) {
    const float alpha = alpha32_of(log_alpha);
    const int Fin = f + A;
    double s_pol = 0, s_lp = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * A; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / A; const int d = (int)(i % A);
        const int t = (int)(r % T);
        float dmu = 0.f, draw = 0.f;
        if (t < T - 1) {
            const float raw = __ldg(ho + r * 2 * A + A + d);
            const float ls = fminf(fmaxf(raw, -20.f), 2.f);
            const float sd = expf(ls);
            const float a = __ldg(a_in + r * a_ld + d), lp = __ldg(logp + i);
            const float da = __ldg(du1 + r * Fin + f + d) + __ldg(du2 + r * Fin + f + d);     // dL/da through the twin critics
            const float cl = alpha * invNA;                                                    // dL/dlogp_d
            const float om = 1.f - a * a;
            const float dx = cl * (2.f * a * om / (om + 1e-7f)) + da * om;
            dmu = dx;
            const float dls = dx * sd * __ldg(eps + i) - cl;
            draw = (raw >= -20.f && raw <= 2.f) ? dls : 0.f;
            s_pol += (double)(alpha * lp);
            s_lp += (double)lp;
        }
        dho[r * 2 * A + d] = dmu;
        dho[r * 2 * A + A + d] = draw;
    }
    block_add3(s_pol, s_lp, 0.0, scalars, 0, 2, -1);
}

__global__ void __launch_bounds__(256)
sac_cont_critic_head_kernel(const float* __restrict__ tq1, const float* __restrict__ tq2, const float* __restrict__ logp,
                            const float* __restrict__ q1, const float* __restrict__ q2,
                            const float* __restrict__ rew, const float* __restrict__ is_fir,
                            const double* __restrict__ log_alpha, float* __restrict__ dq1, float* __restrict__ dq2,
                            float* __restrict__ td_out, double* __restrict__ scalars,
                            long long rows, int T, int A, float gamma, float invN) {
    const float alpha = alpha32_of(log_alpha);
    double s_val = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        float g1 = 0.f, g2 = 0.f;
        if (t < T - 1) {
            const long long j = i + 1;
            const float mq = fminf(__ldg(tq1 + j), __ldg(tq2 + j));
            float soft = 0.f;
            for (int d = 0; d < A; ++d) soft += mq - alpha * __ldg(logp + j * A + d);
            const float y = __ldg(rew + i) + (1.f - __ldg(is_fir + j)) * gamma * soft;
            const float d1 = __ldg(q1 + i) - y, d2 = __ldg(q2 + i) - y;
            g1 = invN * smooth_l1_grad(d1); g2 = invN * smooth_l1_grad(d2);
            s_val += (double)smooth_l1(d1) + (double)smooth_l1(d2);
            if (td_out) td_out[(i / T) * (T - 1) + t] = y;
        }
        dq1[i] = g1; dq2[i] = g2;
    }
    block_add3(s_val, 0.0, 0.0, scalars, 1, -1, -1);
}

// state = {log_alpha, m, v}; loss_alpha = exp(log_alpha) * (scalars[2]/den + target_entropy); Adam (fp64) on log_alpha
__global__ void sac_alpha_step_kernel(double* state, const double* scalars, double* out, double den, double target_entropy,
                                      double lr, double b1, double b2, double eps, int step) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double alpha = exp(state[0]);
    const double term = (double)((float)(scalars[2] / den) + (float)target_entropy);
    const double g = alpha * term;
    out[0] = alpha * term;                      // loss_alpha
    state[1] = b1 * state[1] + (1 - b1) * g;
    state[2] = b2 * state[2] + (1 - b2) * g * g;
    const double bc1 = 1 - pow(b1, (double)step), bc2 = 1 - pow(b2, (double)step);
    state[0] -= (lr / bc1) * state[1] / (sqrt(state[2]) / sqrt(bc2) + eps);
    out[1] = exp(state[0]);                     // alpha after the step
}

static inline int ew_grid(long long n) { long long g = (n + 255) / 256; return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g)); }

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_sac_actor_head(const float* z, const float* q1, const float* q2, const double* log_alpha,
                                     float* dz, double* scalars, int B, int T, int A, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2 && A >= 1 && A <= MAXA, "sac_actor_head: bad shape B=%d T=%d A=%d", B, T, A);
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_actor_head_kernel, ew_grid(rows), 256, 0, stream, z, q1, q2, log_alpha, dz, scalars, rows, T, A,
                  (float)(1.0 / ((double)B * (T - 1))));
    return check_launch("sac_actor_head");
}

extern "C" int b200rl_sac_critic_head(const float* z, const float* q1, const float* q2, const float* tq1, const float* tq2,
                                      const float* act, const float* rew, const float* is_fir, const double* log_alpha,
                                      float* dq1, float* dq2, float* td_out, double* scalars,
                                      int B, int T, int A, float gamma, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2 && A >= 1 && A <= MAXA, "sac_critic_head: bad shape B=%d T=%d A=%d", B, T, A);
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_critic_head_kernel, ew_grid(rows), 256, 0, stream, z, q1, q2, tq1, tq2, act, rew, is_fir, log_alpha, dq1, dq2,
                  td_out, scalars, rows, T, A, gamma, (float)(1.0 / ((double)B * (T - 1))));
    return check_launch("sac_critic_head");
}

extern "C" int b200rl_sac_cont_sample(const float* head_out, const float* eps_in, float* eps_out, unsigned long long* rng_state,
                                      float* act_out, int act_ld, float* logp_out, int B, int T, int A, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1 && A >= 1 && A <= MAXA, "sac_cont_sample: bad shape");
    B200RL_REQUIRE(eps_in != nullptr || rng_state != nullptr, "sac_cont_sample: neither noise nor a generator state given");
    B200RL_REQUIRE(act_ld >= A, "sac_cont_sample: act_ld=%d < A=%d", act_ld, A);
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_cont_sample_kernel, ew_grid(rows * A), 256, 0, stream, head_out, eps_in, eps_out, rng_state, act_out, act_ld, logp_out, rows, A);
    return check_launch("sac_cont_sample");
}

extern "C" int b200rl_sac_cont_pack(const float* obs, const float* act, float* u1, float* u2, float* u3,
                                    int B, int T, int f, int A, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 1 && f >= 1 && A >= 1, "sac_cont_pack: bad shape");
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_cont_pack_kernel, ew_grid(rows * (f + A)), 256, 0, stream, obs, act, u1, u2, u3, rows, f, A);
    return check_launch("sac_cont_pack");
}

extern "C" int b200rl_sac_cont_minq_grad(const float* q1, const float* q2, float* dq1, float* dq2, double* scalars,
                                         int B, int T, int A, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2, "sac_cont_minq_grad: bad shape");
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_cont_minq_grad_kernel, ew_grid(rows), 256, 0, stream, q1, q2, dq1, dq2, scalars, rows, T,
                  (float)(1.0 / ((double)B * (T - 1))), (float)A);
    return check_launch("sac_cont_minq_grad");
}

extern "C" int b200rl_sac_cont_actor_grad(const float* head_out, const float* eps, const float* act_in, int act_ld, const float* logp,
                                          const float* du1, const float* du2, const double* log_alpha,
                                          float* dhead, double* scalars, int B, int T, int A, int f, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2 && A >= 1 && A <= MAXA, "sac_cont_actor_grad: bad shape");
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_cont_actor_grad_kernel, ew_grid(rows * A), 256, 0, stream, head_out, eps, act_in, act_ld, logp, du1, du2, log_alpha,
                  dhead, scalars, rows, T, A, f, (float)(1.0 / ((double)B * (T - 1) * A)));
    return check_launch("sac_cont_actor_grad");
}

extern "C" int b200rl_sac_cont_critic_head(const float* tq1, const float* tq2, const float* logp, const float* q1, const float* q2,
                                           const float* rew, const float* is_fir, const double* log_alpha,
                                           float* dq1, float* dq2, float* td_out, double* scalars,
                                           int B, int T, int A, float gamma, void* stream) {
    B200RL_REQUIRE(B >= 1 && T >= 2 && A >= 1 && A <= MAXA, "sac_cont_critic_head: bad shape");
    const long long rows = (long long)B * T;
    B200RL_LAUNCH(sac_cont_critic_head_kernel, ew_grid(rows), 256, 0, stream, tq1, tq2, logp, q1, q2, rew, is_fir, log_alpha, dq1, dq2,
                  td_out, scalars, rows, T, A, gamma, (float)(1.0 / ((double)B * (T - 1))));
    return check_launch("sac_cont_critic_head");
}

extern "C" int b200rl_sac_alpha_step(double* state, const double* scalars, double* out, double denominator,
                                     double target_entropy, double lr, int step, void* stream) {
    B200RL_REQUIRE(step >= 1, "sac_alpha_step: step=%d", step);
    B200RL_LAUNCH(sac_alpha_step_kernel, 1, 32, 0, stream, state, scalars, out, denominator, target_entropy, lr, 0.9, 0.999, 1e-8, step);
    return check_launch("sac_alpha_step");
}
