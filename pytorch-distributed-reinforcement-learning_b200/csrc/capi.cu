// Error plumbing + version for the C-ABI (include/b200rl.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace b200rl {
static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e));
    return 0;
}
}  // namespace b200rl

extern "C" const char* b200rl_last_error(void) { return b200rl::g_err; }
extern "C" int b200rl_version(void) { return 100; }

// Strided host rows -> contiguous device rows in ONE async DMA (cudaMemcpy2DAsync): the learner ships hx / cx = [:, 0] of the
// page-locked (B, T, H) shared arrays without gathering 4096 rows scattered over 134 MB on the host first.
extern "C" int b200rl_h2d_rows(void* dst, const void* src, long long width_bytes, long long rows, long long src_pitch_bytes, void* stream) {
#ifdef B200RL_EMUL
    for (long long r = 0; r < rows; ++r) memcpy((char*)dst + r * width_bytes, (const char*)src + r * src_pitch_bytes, (size_t)width_bytes);
    return 0;
#else
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)width_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes, (size_t)rows,
                                      cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return b200rl::set_error("h2d_rows: %s", cudaGetErrorString(e));
    return 0;
#endif
}
