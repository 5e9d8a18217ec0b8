// Fused optimizer step on the flat parameter vector (sm_100a).
//
//   clip_grad_norm_(params, max_norm) + RMSprop.step()
//        agents/learner_module/ppo/learning.py:93-106 ; optimizer built at agents/learner.py:70
//   clip_grad_norm_ + Adam.step()       agents/learner.py:360-367 (SAC actor / twin critic)
//
// The model is tiny (33 795 fp32 parameters for the CartPole actor-critic), so the whole
// step -- optional 1/world scaling after a gradient all-reduce, global L2 norm, clip
// coefficient min(1, max_norm/(norm+1e-6)), moment update, parameter update -- is ONE
// single-CTA launch: grid-wide sync for the norm would cost more than the 135 KB pass.
#include "common.cuh"

namespace b200rl {

constexpr int kOptThreads = 1024;

__device__ __forceinline__ float grad_norm_and_coef(const float* grad, long long n, float grad_scale,
                                                    float max_norm, float* norm_out, const float* mask = nullptr) {
    __shared__ double scratch[32];
    __shared__ float coef_s;
    double s = 0.0;
    // 8 loads in flight per thread: the CTA is alone on the device, so latency, not bandwidth, is what there is to hide
    for (long long i0 = threadIdx.x; i0 < n; i0 += 8LL * kOptThreads) {
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long i = i0 + (long long)e * kOptThreads;
            g[e] = i < n ? grad[i] * grad_scale * (mask ? mask[i] : 1.0f) : 0.0f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += (double)g[e] * (double)g[e];
    }
    const double tot = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(tot);
        float coef = 1.0f;
        if (max_norm > 0.0f) coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
        coef_s = coef;
        if (norm_out) *norm_out = norm;
    }
    __syncthreads();
    return coef_s;
}

__global__ void __launch_bounds__(kOptThreads)
clip_rmsprop_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ sq,
                    long long n, float grad_scale, float max_norm, float lr, float alpha, float eps,
                    float* norm_out) {
    const float coef = grad_norm_and_coef(grad, n, grad_scale, max_norm, norm_out);
    for (long long i0 = threadIdx.x; i0 < n; i0 += 4LL * kOptThreads) {
        float g[4], q[4], w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long i = i0 + (long long)e * kOptThreads;
            if (i < n) { g[e] = grad[i]; q[e] = sq[i]; w[e] = param[i]; }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long i = i0 + (long long)e * kOptThreads;
            if (i < n) {
                const float gg = g[e] * grad_scale * coef;
                grad[i] = gg;                                   // .grad is left clipped, as in torch
                const float v = q[e] * alpha + (1.0f - alpha) * gg * gg;
                sq[i] = v;
                param[i] = w[e] - lr * (gg / (sqrtf(v) + eps));
            }
        }
    }
}

__global__ void __launch_bounds__(kOptThreads)
clip_adam_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, long long n, float grad_scale, float max_norm, float lr,
                 float beta1, float beta2, float eps, int step, float* norm_out, const float* __restrict__ mask) {
    const float coef = grad_norm_and_coef(grad, n, grad_scale, max_norm, norm_out, mask);
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float rsq_bc2 = (float)(1.0 / sqrt(bc2));
    for (long long i0 = threadIdx.x; i0 < n; i0 += 4LL * kOptThreads) {
        float g[4], m0[4], v0[4], w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long i = i0 + (long long)e * kOptThreads;
            if (i < n) { g[e] = grad[i] * (mask ? mask[i] : 1.0f); m0[e] = m[i]; v0[e] = v[i]; w[e] = param[i]; }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long i = i0 + (long long)e * kOptThreads;
            if (i < n) {
                const float gg = g[e] * grad_scale * coef;
                grad[i] = gg;
                const float mi = m0[e] * beta1 + (1.0f - beta1) * gg;
                const float vi = v0[e] * beta2 + (1.0f - beta2) * gg * gg;
                m[i] = mi; v[i] = vi;
                param[i] = w[e] - step_size * (mi / (sqrtf(vi) * rsq_bc2 + eps));
            }
        }
    }
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_clip_rmsprop(float* param, float* grad, float* sq_avg, long long n,
                                   float grad_scale, float max_norm, float lr, float alpha, float eps,
                                   float* norm_out, void* stream) {
    B200RL_REQUIRE(n >= 1, "clip_rmsprop: n=%lld", n);
    B200RL_LAUNCH(clip_rmsprop_kernel, 1, kOptThreads, 0, stream, param, grad, sq_avg, n, grad_scale, max_norm, lr, alpha, eps, norm_out);
    return check_launch("clip_rmsprop");
}

extern "C" int b200rl_clip_adam(float* param, float* grad, float* m, float* v, long long n,
                                float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                float eps, int step, const float* mask, float* norm_out, void* stream) {
    B200RL_REQUIRE(n >= 1 && step >= 1, "clip_adam: n=%lld step=%d", n, step);
    B200RL_LAUNCH(clip_adam_kernel, 1, kOptThreads, 0, stream, param, grad, m, v, n, grad_scale, max_norm, lr, beta1, beta2, eps, step, norm_out, mask);
    return check_launch("clip_adam");
}
