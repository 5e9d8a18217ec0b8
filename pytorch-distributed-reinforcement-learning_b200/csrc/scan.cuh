// Warp-level suffix scan of affine maps -- the primitive behind every per-timestep
// recurrence on the learner path (GAE, V-trace).
#pragma once
#include "common.cuh"

namespace b200rl {

// Lane l enters with the map f_l(y) = a*y + b for timestep t0 + l.
// Returns y_l = f_l(f_{l+1}(... f_31(carry))), i.e. the recurrence value at t0 + l
// given y[t0 + 32] = carry.  Identity lanes pass (a, b) = (1, 0).
__device__ __forceinline__ float affine_suffix_scan(float a, float b, float carry) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float a2 = __shfl_down_sync(kFull, a, d);
        float b2 = __shfl_down_sync(kFull, b, d);
        if (lane + d < 32) { b = fmaf(a, b2, b); a = a * a2; }
    }
    return fmaf(a, carry, b);
}

}  // namespace b200rl
