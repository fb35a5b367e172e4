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

#ifdef __CUDACC__
// GroupNorm mean / rstd of batch element b from the S fp64 partial sums of gn_stats / splitk_reduce_stats
// ([B][S][32 groups][sum, sumsq]) into shared memory; called by all 256 threads of a CTA, followed by __syncthreads().
// Eight lanes per group stride over the splits and are combined by a fixed butterfly: with up to 128 splits a serial
// loop by one thread per group was a chain of 128 dependent global loads in the prologue of EVERY CTA of the apply
// kernels -- 17-25 us per launch on the small layers (ncu pass E), several times the kernel's own streaming time.
__device__ __forceinline__ void gn_mean_rstd_from_partials(const double *__restrict__ partial, int b, int S, double n,
                                                          float *mean_s, float *rstd_s) {
    const int tid = threadIdx.x, g = tid >> 3, part = tid & 7;
    double a = 0.0, q = 0.0;
    for (int s0 = 0; s0 < S; s0 += 128) {               // S <= 128 in practice: one round, all 16 loads in flight at once
        double2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int s = s0 + part + 8 * i;
            v[i] = s < S ? __ldg(reinterpret_cast<const double2 *>(partial + (((size_t)b * S + s) * 32 + g) * 2)) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { a += v[i].x; q += v[i].y; }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (part == 0) {
        const double mean = a / n;
        double var = q / n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        mean_s[g] = (float)mean;
        rstd_s[g] = (float)(1.0 / sqrt(var + 1e-6));
    }
}
// The same from the fp32 per-pixel-block sums the GEMM epilogues emit ([B][tiles][32 groups][sum, sumsq]); for tensors of at
// most 64 blocks per image the apply kernel finalises them itself instead of waiting for a separate 6 us launch.
__device__ __forceinline__ void gn_mean_rstd_from_tiles(const float *__restrict__ tile_partial, int b, int tiles, double n,
                                                       float *mean_s, float *rstd_s) {
    const int tid = threadIdx.x, g = tid >> 3, part = tid & 7;
    double a = 0.0, q = 0.0;
    for (int t0 = 0; t0 < tiles; t0 += 64) {
        float2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + part + 8 * i;
            v[i] = t < tiles ? __ldg(reinterpret_cast<const float2 *>(tile_partial + (((size_t)b * tiles + t) * 32 + g) * 2)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a += (double)v[i].x; q += (double)v[i].y; }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (part == 0) {
        const double mean = a / n;
        double var = q / n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        mean_s[g] = (float)mean;
        rstd_s[g] = (float)(1.0 / sqrt(var + 1e-6));
    }
}
#endif

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A frame is ~300 dependent kernels, many of them a few microseconds long, so
// the launch / dependency-resolution latency between consecutive kernels is a visible share of a single-trajectory
// step.  Kernels launched through sgam_launch_pdl carry cudaLaunchAttributeProgrammaticStreamSerialization: the
// front end may start scheduling their CTAs as soon as every CTA of the preceding kernel has executed
// griddepcontrol.launch_dependents (or exited).  Every such kernel begins with SGAM_PDL_PROLOGUE() --
// launch_dependents, then griddepcontrol.wait, which blocks until the preceding grid has COMPLETED and its memory is
// visible -- before it touches global memory, so the data dependences are exactly those of plain stream order (also
// when captured into a CUDA graph, where the edge becomes a programmatic dependency).  SGAM_PDL=0 launches them as
// ordinary kernels (the two instructions are then no-ops); SGAM_PDL=<mask> enables it per kernel family.
// STATUS (round 1): EXPERIMENTAL, default mask 0.  Measured on B200 (profiles/r1_pdl_experiment.txt): GEMM families
// only (mask 3) +5 % single-trajectory frames/s, +1 % at 8 trajectories, all parity tests green; elementwise families
// only (mask 28) no gain; GEMMs + softmax (mask 11) HANGS at 512x512 (softmax_split_kernel<8>, 16384 CTAs of 87
// registers, as the dependent of the 2-CTA QK^T GEMM) although every kernel waits before touching memory -- not
// understood yet, so nothing is enabled by default.
#ifdef __CUDACC__
#define SGAM_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define SGAM_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define SGAM_PDL_PROLOGUE() \
    do {                    \
        SGAM_PDL_TRIGGER(); \
        SGAM_PDL_WAIT();    \
    } while (0)

bool sgam_pdl_enabled(int family);
enum { SGAM_PDL_GEMM1 = 1, SGAM_PDL_GEMM2 = 2, SGAM_PDL_NORM = 4, SGAM_PDL_SOFTMAX = 8, SGAM_PDL_MISC = 16 };

template <typename... KArgs, typename... Args>
static inline cudaError_t sgam_launch_pdl(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = sgam_pdl_enabled(family) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// kernel<<<grid, block, smem, stream>>>(args...) + SGAM_LAUNCH_OK(), as a PDL launch; parenthesise template-ids with commas
#define SGAM_PDL_LAUNCH(family, kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                                 \
        ++g_sgam_launches;                                                                               \
        SGAM_CUDA_OK(sgam_launch_pdl(family, kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)); \
    } while (0)
#endif
