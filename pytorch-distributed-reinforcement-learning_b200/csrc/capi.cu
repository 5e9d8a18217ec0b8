// Error plumbing + version for the C-ABI (include/b200rl.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

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
