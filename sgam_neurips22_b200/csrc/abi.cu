// Error plumbing and device queries of the C ABI.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

// OFF by default: see the note in common.cuh (measured +5 % single-trajectory with the GEMM families only, and an
// unexplained hang with softmax_split_kernel<8> as a dependent at 512x512).
#define SGAM_PDL_DEFAULT_MASK 0

static thread_local char g_err[512] = "";
unsigned long long g_sgam_launches = 0;

void sgam_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool sgam_pdl_enabled(int family) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("SGAM_PDL"); v = e ? atoi(e) : SGAM_PDL_DEFAULT_MASK; }
    return (v & family) != 0;
}

extern "C" const char *sgam_last_error(void) { return g_err; }
extern "C" int sgam_version(void) { return 100; }
extern "C" int sgam_sm_count(int device) {
    int n = 0;
    SGAM_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    return n;
}
extern "C" unsigned long long sgam_launch_count(void) { return g_sgam_launches; }
