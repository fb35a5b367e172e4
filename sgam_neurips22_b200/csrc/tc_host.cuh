// Host-side helpers shared by the tcgen05 GEMM translation units: tensor-map encoding through the runtime-resolved
// driver entry point, SM count.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {   // resolved through the runtime so that the library does not link libcuda (it must load on CPU-only hosts)
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

// bf16 tensor [d3][d2][d1][d0] (d0 innermost, contiguous), box {b0, b1, b2, 1}, swizzle span = box[0] * 2 bytes, zero OOB fill;
// pitch0: elements between consecutive d1 rows when the rows are slices of a wider tensor (0 = dims[0], dense)
inline int make_map(CUtensorMap *m, const void *base, int rank, const long long *dims, const int *box, const int *estride = nullptr,
                    long long pitch0 = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { sgam_set_error("cuTensorMapEncodeTiled is unavailable"); return SGAM_ERR_CUDA; }
    cuuint64_t gd[4], gs[3];
    cuuint32_t bd[4], es[4];
    unsigned long long stride = 2;
    for (int i = 0; i < rank; ++i) {
        gd[i] = (cuuint64_t)dims[i]; bd[i] = (cuuint32_t)box[i]; es[i] = estride ? (cuuint32_t)estride[i] : 1;
        stride *= (unsigned long long)((i == 0 && pitch0) ? pitch0 : dims[i]);
        if (i < rank - 1) gs[i] = stride;
    }
    const CUtensorMapSwizzle sw = (box[0] == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bd, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sgam_set_error("cuTensorMapEncodeTiled failed with %d (rank %d dims %lld %lld %lld)", (int)r, rank, dims[0], dims[1], rank > 2 ? dims[2] : 0); return SGAM_ERR_CUDA; }
    return SGAM_OK;
}

// fp32 tensor [d3][d2][d1][d0] (d0 innermost, contiguous), box {b0, b1, 1, 1}, no swizzle (dense rows of b0 floats in shared memory)
inline int make_map_f32(CUtensorMap *m, const void *base, int rank, const long long *dims, const int *box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { sgam_set_error("cuTensorMapEncodeTiled is unavailable"); return SGAM_ERR_CUDA; }
    cuuint64_t gd[4], gs[3];
    cuuint32_t bd[4], es[4];
    unsigned long long stride = 4;
    for (int i = 0; i < rank; ++i) {
        gd[i] = (cuuint64_t)dims[i]; bd[i] = (cuuint32_t)box[i]; es[i] = 1;
        stride *= (unsigned long long)dims[i];
        if (i < rank - 1) gs[i] = stride;
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bd, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { sgam_set_error("cuTensorMapEncodeTiled (fp32) failed with %d", (int)r); return SGAM_ERR_CUDA; }
    return SGAM_OK;
}

inline int sm_count_cached() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// 2-CTA kernel (net_tc2.cu)
struct TcParams;
bool tc2_applicable(int B, int Ho, int Wo, int N, int stride, int out_nchw, int n_valid, int *BN_out);
int launch_tc2(int BN, const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &b_hi, const CUtensorMap &b_lo, TcParams p,
               int B, int Ho, int Wo, int N, cudaStream_t s);
// swapped-operand kernel for 128-channel convolutions (net_tc3.cu)
bool swap_applicable(int B, int Ho, int Wo, int Cin, int Cout, int stride, bool fp32_nhwc_only);
int launch_conv_swap(const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo, TcParams p, int B, int H, int W, int Cin,
                     int Cout, int ksize, cudaStream_t s);
// GroupNorm + swish applied inside the conv's operand path (net_tc3.cu): x fp32 NHWC, mean / rstd [B][32][2], gamma / beta [Cin]
bool gnconv_applicable(int B, int H, int W, int Cin, int Cout);
int launch_gnconv(const float *x, const float *meanrstd, const float *gamma, const float *beta, const void *w_hi, const void *w_lo, TcParams p,
                  int B, int H, int W, int Cin, int Cout, cudaStream_t s);
}  // namespace tc
