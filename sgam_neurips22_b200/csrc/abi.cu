// Error plumbing and device queries of the C ABI.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_sgam_launches = 0;

void sgam_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *sgam_last_error(void) { return g_err; }
extern "C" int sgam_version(void) { return 100; }
extern "C" int sgam_sm_count(int device) {
    int n = 0;
    SGAM_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    return n;
}
extern "C" unsigned long long sgam_launch_count(void) { return g_sgam_launches; }
