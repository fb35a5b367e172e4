// Shared helpers for libsgam_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sgam_b200.h"

void sgam_set_error(const char *fmt, ...);

#define SGAM_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            sgam_set_error(__VA_ARGS__);        \
            return SGAM_ERR_INVALID;            \
        }                                       \
    } while (0)

#define SGAM_CUDA_OK(call)                                                            \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            sgam_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return SGAM_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

// every kernel launch of the library goes through this macro: it also feeds sgam_launch_count()
extern unsigned long long g_sgam_launches;
#define SGAM_LAUNCH_OK()                  \
    do {                                  \
        ++g_sgam_launches;                \
        SGAM_CUDA_OK(cudaGetLastError()); \
    } while (0)

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Monotone map float -> uint32 (a < b  <=>  ord(a) < ord(b), -0 < +0).
__device__ __forceinline__ uint32_t float_orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float orderable_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// MKL's K=3 sgemm micro-kernel order (probed against torch.bmm on CPU; oracle/csrc/oracle.c dot3):
// one rounded product followed by two fused multiply-adds.
__device__ __forceinline__ float dot3(const float *a, float x0, float x1, float x2) {
    float t = __fmul_rn(a[0], x0);
    t = __fmaf_rn(a[1], x1, t);
    t = __fmaf_rn(a[2], x2, t);
    return t;
}
