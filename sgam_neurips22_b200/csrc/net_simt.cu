// Stage (iii), fp32 CUDA-core path: implicit-GEMM conv (3x3 / 1x1, stride, asymmetric pad, fused nearest
// up-sampling, bias + residual epilogue), batched A.B^T, GroupNorm(+swish), row softmax, and the two
// odd-shaped layers (5->4 stem, 3x3 -> 4 channel NCHW head).  NHWC activations, K-major weights.
//
// This is the exact-fp32 path: every product is accumulated in fp32 in k order.  The tensor-core path
// (conv_tc.cu) replaces sgam_conv2d / sgam_gemm_nt for the large layers; this file stays as the
// reference implementation on the device and handles the shapes the tensor-core kernels do not.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// implicit GEMM:  C[m, n] = alpha * sum_k A(m, k) * Wt[n, k]  (+ bias_n[n]) (+ bias_m[m]) (+ R[m, n])
//   A(m, k): conv mode   -> x[b, iy, ix, ci] with m = (b, oy, ox), k = (kh*ks + kw)*Cin + ci
//            plain mode  -> A[m*lda + k]
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    const float *A;        // activations (NHWC) or plain matrix
    const float *Wt;       // [N, K] K-major
    const float *bias_n;   // [N] or null
    const float *bias_m;   // [M] or null
    const float *R;        // residual [M, N] or null
    float *C;
    long long sA, sB, sC;  // batch strides (plain mode)
    int M, N, K;
    int lda;               // plain mode row stride
    float alpha;
    // conv geometry (conv mode when ks > 0)
    int ks, stride, pad, up, H, W, Cin, Ho, Wo;
};

constexpr int BM = 128, BK = 16, NTHR = 256;

__device__ __forceinline__ const float *conv_src(const GemmParams &p, int m, int tap, bool &valid) {
    // m -> (b, oy, ox); tap -> (kh, kw); logical input is (H<<up) x (W<<up)
    const int ox = m % p.Wo, t = m / p.Wo, oy = t % p.Ho, b = t / p.Ho;
    const int kh = tap / p.ks, kw = tap - kh * p.ks;
    const int iy = oy * p.stride + kh - p.pad, ix = ox * p.stride + kw - p.pad;
    valid = (iy >= 0) && (iy < (p.H << p.up)) && (ix >= 0) && (ix < (p.W << p.up));
    return p.A + (((size_t)b * p.H + (iy >> p.up)) * p.W + (ix >> p.up)) * p.Cin;
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(NTHR)
gemm_simt_kernel(GemmParams p) {
    constexpr int TN = BN / 16;                  // columns per thread (8 or 4)
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const float *Ab = p.A + (size_t)blockIdx.z * p.sA;
    const float *Bb = p.Wt + (size_t)blockIdx.z * p.sB;
    float *Cb = p.C + (size_t)blockIdx.z * p.sC;
    const bool conv = p.ks > 0;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    // global -> register staging: A tile 128x16 (2 float4 per thread), B tile BNx16 (BN/64 float4 per thread)
    constexpr int A_LD = BM * BK / 4 / NTHR, B_LD = (BN * BK / 4 + NTHR - 1) / NTHR;
    float4 ra[A_LD], rb[B_LD];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int r = 0; r < A_LD; ++r) {
            const int f = tid + r * NTHR, row = f >> 2, kq = f & 3, m = m0 + row, k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (VEC) {
                if (m < p.M) {
                    if (conv) {
                        bool ok;
                        const int tap = k / p.Cin;
                        const float *src = conv_src(p, m, tap, ok);
                        if (ok) v = __ldg(reinterpret_cast<const float4 *>(src + (k - tap * p.Cin)));
                    } else {
                        v = __ldg(reinterpret_cast<const float4 *>(Ab + (size_t)m * p.lda + k));
                    }
                }
            } else {
                float e[4] = {0.f, 0.f, 0.f, 0.f};
                if (m < p.M) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int kk = k + c;
                        if (kk < p.K) {
                            if (conv) {
                                bool ok;
                                const int tap = kk / p.Cin;
                                const float *src = conv_src(p, m, tap, ok);
                                if (ok) e[c] = __ldg(src + (kk - tap * p.Cin));
                            } else {
                                e[c] = __ldg(Ab + (size_t)m * p.lda + kk);
                            }
                        }
                    }
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            ra[r] = v;
        }
#pragma unroll
        for (int r = 0; r < B_LD; ++r) {
            const int f = tid + r * NTHR, row = f >> 2, kq = f & 3, n = n0 + row, k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < BN && n < p.N) {
                if (VEC) {
                    v = __ldg(reinterpret_cast<const float4 *>(Bb + (size_t)n * p.K + k));
                } else {
                    float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (k + c < p.K) e[c] = __ldg(Bb + (size_t)n * p.K + k + c);
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            rb[r] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int r = 0; r < A_LD; ++r) {
            const int f = tid + r * NTHR, row = f >> 2, kq = f & 3;
            As[buf][kq * 4 + 0][row] = ra[r].x; As[buf][kq * 4 + 1][row] = ra[r].y;
            As[buf][kq * 4 + 2][row] = ra[r].z; As[buf][kq * 4 + 3][row] = ra[r].w;
        }
#pragma unroll
        for (int r = 0; r < B_LD; ++r) {
            const int f = tid + r * NTHR, row = f >> 2, kq = f & 3;
            if (row < BN) {
                Bs[buf][kq * 4 + 0][row] = rb[r].x; Bs[buf][kq * 4 + 1][row] = rb[r].y;
                Bs[buf][kq * 4 + 2][row] = rb[r].z; Bs[buf][kq * 4 + 3][row] = rb[r].w;
            }
        }
    };

    const int nk = (p.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
            {
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
                bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                if (TN == 8) {
                    const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][BN / 2 + tx * 4]);
                    bv[4 % TN] = b1.x; bv[5 % TN] = b1.y; bv[6 % TN] = b1.z; bv[7 % TN] = b1.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (m >= p.M) continue;
        const float bm = p.bias_m ? __ldg(p.bias_m + m) : 0.0f;
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int n = n0 + h * (BN / 2) + tx * 4;
            if (n + 3 < p.N) {
                float4 o;
                o.x = p.alpha * acc[i][h * 4 + 0] + bm; o.y = p.alpha * acc[i][h * 4 + 1] + bm;
                o.z = p.alpha * acc[i][h * 4 + 2] + bm; o.w = p.alpha * acc[i][h * 4 + 3] + bm;
                if (p.bias_n) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias_n + n));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                if (p.R) {
                    const float4 r = __ldg(reinterpret_cast<const float4 *>(p.R + (size_t)m * p.N + n));
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4 *>(Cb + (size_t)m * p.N + n) = o;
            } else {
                for (int c = 0; c < 4; ++c) {
                    if (n + c < p.N) {
                        float o = p.alpha * acc[i][h * 4 + c] + bm;
                        if (p.bias_n) o += __ldg(p.bias_n + n + c);
                        if (p.R) o += __ldg(p.R + (size_t)m * p.N + n + c);
                        Cb[(size_t)m * p.N + n + c] = o;
                    }
                }
            }
        }
    }
}

int launch_gemm(const GemmParams &p, int batch, cudaStream_t s) {
    const bool vec = (p.K % BK == 0) && ((p.ks > 0) ? (p.Cin % BK == 0) : (p.lda % 4 == 0)) && (p.N % 4 == 0) &&
                     ((uintptr_t)p.A % 16 == 0) && ((uintptr_t)p.Wt % 16 == 0) && ((uintptr_t)p.C % 16 == 0) &&
                     (p.sA % 4 == 0) && (p.sB % 4 == 0) && (p.sC % 4 == 0) &&
                     (!p.R || (uintptr_t)p.R % 16 == 0) && (!p.bias_n || (uintptr_t)p.bias_n % 16 == 0);
    // wide tile when there is enough work to fill the machine with it
    const bool wide = (p.N >= 128) && ((long long)cdiv(p.M, BM) * cdiv(p.N, 128) * batch >= 148);
    if (wide) {
        dim3 grid(cdiv(p.M, BM), cdiv(p.N, 128), batch);
        if (vec) gemm_simt_kernel<128, true><<<grid, NTHR, 0, s>>>(p);
        else gemm_simt_kernel<128, false><<<grid, NTHR, 0, s>>>(p);
    } else {
        dim3 grid(cdiv(p.M, BM), cdiv(p.N, 64), batch);
        if (vec) gemm_simt_kernel<64, true><<<grid, NTHR, 0, s>>>(p);
        else gemm_simt_kernel<64, false><<<grid, NTHR, 0, s>>>(p);
    }
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

// ------------------------------------------------------------------------------------------------
// stem: cat(x, mask) -> 1x1 conv 5 -> 4, NCHW in, NHWC out (model.py:106-113)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_kernel(const float *__restrict__ x, const uint8_t *__restrict__ mask, const float *__restrict__ w,
            const float *__restrict__ bias, int HW, float *__restrict__ y) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= HW) return;
    float in[5];
#pragma unroll
    for (int c = 0; c < 4; ++c) in[c] = x[((size_t)b * 4 + c) * HW + p];
    in[4] = mask ? (mask[(size_t)b * HW + p] ? 1.0f : 0.0f) : 0.0f;
    float o[4];
#pragma unroll
    for (int co = 0; co < 4; ++co) {
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < 5; ++c) a = fmaf(in[c], __ldg(w + co * 5 + c), a);
        o[co] = a + __ldg(bias + co);
    }
    *reinterpret_cast<float4 *>(y + ((size_t)b * HW + p) * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------
// head: 3x3 conv Cin -> Cout<=4, NHWC in, NCHW or NHWC out.  One warp per output pixel; lanes split Cin.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv3x3_head_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                    float *__restrict__ y, int B, int H, int W, int Cin, int Cout, int out_nchw) {
    extern __shared__ float ws[];                     // [Cout][9*Cin]
    for (int e = threadIdx.x; e < Cout * 9 * Cin; e += blockDim.x) ws[e] = w[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (pix >= (long long)B * H * W) return;
    const int ox = (int)(pix % W), oy = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < 9; ++tap) {
        const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const float *src = x + (((size_t)b * H + iy) * W + ix) * Cin;
        for (int c = lane * 4; c < Cin; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + c));
            for (int co = 0; co < Cout; ++co) {
                const float *wp = ws + (size_t)co * 9 * Cin + tap * Cin + c;
                acc[co] = fmaf(v.x, wp[0], acc[co]); acc[co] = fmaf(v.y, wp[1], acc[co]);
                acc[co] = fmaf(v.z, wp[2], acc[co]); acc[co] = fmaf(v.w, wp[3], acc[co]);
            }
        }
    }
#pragma unroll
    for (int co = 0; co < 4; ++co)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
    if (lane < Cout) {
        const float v = acc[lane] + __ldg(bias + lane);
        if (out_nchw) y[(((size_t)b * Cout + lane) * H + oy) * W + ox] = v;
        else y[(size_t)pix * Cout + lane] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm(32 groups, eps 1e-6) + optional swish.  x [B, HW, C], C % 128 == 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float *__restrict__ x, double *__restrict__ partial, long long HW, int C, int S) {
    SGAM_PDL_PROLOGUE();
    __shared__ double red[256][2];
    const int b = blockIdx.y, s = blockIdx.x, tid = threadIdx.x;
    const int CQ = C / 4, PL = 256 / CQ, cq = tid % CQ, pl = tid / CQ;
    const long long chunk = (HW + S - 1) / S, pbeg = s * chunk, pend = min(HW, pbeg + chunk);
    float sum = 0.f, sq = 0.f;
    if (pl < PL) {
        const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)b * HW * C) + cq;
        for (long long p = pbeg + pl; p < pend; p += PL) {
            const float4 v = __ldg(src + p * CQ);
            sum += (v.x + v.y) + (v.z + v.w);
            sq = fmaf(v.x, v.x, sq); sq = fmaf(v.y, v.y, sq); sq = fmaf(v.z, v.z, sq); sq = fmaf(v.w, v.w, sq);
        }
    }
    red[tid][0] = (double)sum; red[tid][1] = (double)sq;
    __syncthreads();
    if (tid < 32) {                                        // fixed-order (deterministic) group reduction
        const int nq = (C / 32) / 4;                       // float4 lanes per group
        double a = 0.0, q = 0.0;
        for (int l = 0; l < PL; ++l)
            for (int k = 0; k < nq; ++k) {
                const int t = l * CQ + tid * nq + k;
                a += red[t][0]; q += red[t][1];
            }
        double *dst = partial + (((size_t)b * S + s) * 32 + tid) * 2;
        dst[0] = a; dst[1] = q;
    }
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const float *__restrict__ x, const double *__restrict__ partial, const float *__restrict__ gamma,
                const float *__restrict__ beta, float *__restrict__ y, long long HW, int C, int S, int swish) {
    __shared__ float mean_s[32], rstd_s[32];
    const int b = blockIdx.y, tid = threadIdx.x;
    gn_mean_rstd_from_partials(partial, b, S, (double)HW * (C / 32), mean_s, rstd_s);
    __syncthreads();
    const int CQ = C / 4, cpg = C / 32;
    const long long total = HW * CQ;
    const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)b * HW * C);
    float4 *dst = reinterpret_cast<float4 *>(y + (size_t)b * HW * C);
    for (long long e = (long long)blockIdx.x * blockDim.x + tid; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(e % CQ), c = cq * 4, g = c / cpg;
        const float mu = mean_s[g], rs = rstd_s[g];
        const float4 v = __ldg(src + e), ga = __ldg(reinterpret_cast<const float4 *>(gamma + c)),
                     be = __ldg(reinterpret_cast<const float4 *>(beta + c));
        float o[4] = {(v.x - mu) * rs * ga.x + be.x, (v.y - mu) * rs * ga.y + be.y,
                      (v.z - mu) * rs * ga.z + be.z, (v.w - mu) * rs * ga.w + be.w};
        if (swish) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] / (1.0f + expf(-o[k]));     // x * sigmoid(x), model.py:29-31
        }
        dst[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// row softmax in place (model.py:181)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, bool is_max, float *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    v = (lane < (blockDim.x >> 5)) ? sh[lane] : (is_max ? -INFINITY : 0.0f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(256)
softmax_rows_kernel(float *__restrict__ x, int cols) {
    __shared__ float sh[32];
    float *row = x + (size_t)blockIdx.x * cols;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, row[c]);
    mx = block_reduce(mx, true, sh);
    float sum = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        const float e = expf(row[c] - mx);
        row[c] = e;
        sum += e;
    }
    sum = block_reduce(sum, false, sh);
    const float inv = 1.0f / sum;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) row[c] *= inv;
}

}  // namespace

// shared with net_tc.cu (GroupNorm statistics are the same kernel on both paths)
int sgam_gn_stats_launch(const float *x, double *partial, int B, long long HW, int C, cudaStream_t s) {
    const int S = sgam_gn_splits(HW);
    SGAM_PDL_LAUNCH(SGAM_PDL_NORM, gn_stats_kernel, dim3(S, B), 256, 0, s, x, partial, HW, C, S);
    return SGAM_OK;
}

extern "C" int sgam_stem_conv(const float *x, const uint8_t *mask, const float *w, const float *bias, int B, int H,
                              int W, float *y, void *stream) {
    SGAM_REQUIRE(x && w && bias && y && B > 0 && H > 0 && W > 0, "stem_conv: bad arguments");
    dim3 grid(cdiv((long long)H * W, 256), B);
    stem_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, mask, w, bias, H * W, y);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_conv2d(const float *x, const float *w, const float *bias, const float *residual, float *y, int B,
                           int H, int W, int Cin, int Cout, int ksize, int stride, int pad_mode, int upsample,
                           int out_nchw, void *stream) {
    SGAM_REQUIRE(x && w && bias && y, "conv2d: null pointer");
    SGAM_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv2d: bad shape");
    SGAM_REQUIRE(ksize == 1 || ksize == 3, "conv2d: ksize %d (1 or 3)", ksize);
    SGAM_REQUIRE(stride == 1 || stride == 2, "conv2d: stride %d (1 or 2)", stride);
    SGAM_REQUIRE(pad_mode == 0 || (pad_mode == 1 && ksize == 3 && stride == 2), "conv2d: pad_mode 1 is the 3x3 stride-2 Downsample");
    SGAM_REQUIRE(upsample == 0 || (upsample == 1 && stride == 1 && pad_mode == 0), "conv2d: upsample needs stride 1, symmetric pad");
    cudaStream_t s = (cudaStream_t)stream;
    const int Hl = H << upsample, Wl = W << upsample;
    const int pad = (pad_mode == 0) ? ksize / 2 : 0;
    const int Ho = (pad_mode == 0) ? (Hl + 2 * pad - ksize) / stride + 1 : (Hl + 1 - ksize) / stride + 1;
    const int Wo = (pad_mode == 0) ? (Wl + 2 * pad - ksize) / stride + 1 : (Wl + 1 - ksize) / stride + 1;
    if (Cout <= 4) {
        SGAM_REQUIRE(ksize == 3 && stride == 1 && pad_mode == 0 && !upsample && !residual && Cin % 4 == 0,
                     "conv2d: the small-Cout head supports plain 3x3 only");
        const size_t smem = (size_t)Cout * 9 * Cin * sizeof(float);
        SGAM_REQUIRE(smem <= 96 * 1024, "conv2d: head weights do not fit shared memory");
        if (smem > 48 * 1024)
            SGAM_CUDA_OK(cudaFuncSetAttribute(conv3x3_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long pixels = (long long)B * H * W;
        conv3x3_head_kernel<<<cdiv(pixels, 8), 256, smem, s>>>(x, w, bias, y, B, H, W, Cin, Cout, out_nchw);
        SGAM_LAUNCH_OK();
        return SGAM_OK;
    }
    SGAM_REQUIRE(!out_nchw, "conv2d: NCHW output only for Cout <= 4");
    GemmParams p{};
    p.A = x; p.Wt = w; p.bias_n = bias; p.bias_m = nullptr; p.R = residual; p.C = y;
    p.sA = p.sB = p.sC = 0;
    p.M = B * Ho * Wo; p.N = Cout; p.K = ksize * ksize * Cin; p.lda = 0; p.alpha = 1.0f;
    p.ks = ksize; p.stride = stride; p.pad = pad; p.up = upsample; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo;
    return launch_gemm(p, 1, s);
}

extern "C" int sgam_gemm_nt(const float *A, const float *Bm, float *C, const float *bias_m, int batch, int M, int N,
                            int K, long long sA, long long sB, long long sC, float alpha, void *stream) {
    SGAM_REQUIRE(A && Bm && C && batch > 0 && M > 0 && N > 0 && K > 0, "gemm_nt: bad arguments");
    GemmParams p{};
    p.A = A; p.Wt = Bm; p.C = C; p.bias_m = bias_m; p.bias_n = nullptr; p.R = nullptr;
    p.sA = sA; p.sB = sB; p.sC = sC; p.M = M; p.N = N; p.K = K; p.lda = K; p.alpha = alpha; p.ks = 0;
    return launch_gemm(p, batch, (cudaStream_t)stream);
}

extern "C" int sgam_gn_splits(long long HW) {
    long long s = HW / 2;               // >= 2 pixels per block, at most 128 blocks per batch element: the statistics / split-K
                                        // reduce kernels of the 16 x 16 layers are latency chains, so more, shorter blocks win
    return (int)(s < 1 ? 1 : (s > 128 ? 128 : s));
}

extern "C" int sgam_groupnorm(const float *x, const float *gamma, const float *beta, float *y, double *partial, int B,
                              long long HW, int C, int swish, void *stream) {
    SGAM_REQUIRE(x && gamma && beta && y && partial, "groupnorm: null pointer");
    SGAM_REQUIRE(B > 0 && HW > 0 && C % 128 == 0 && C <= 1024, "groupnorm: C=%d must be a multiple of 128 (<= 1024)", C);
    cudaStream_t s = (cudaStream_t)stream;
    const int S = sgam_gn_splits(HW);
    int rc = sgam_gn_stats_launch(x, partial, B, HW, C, s);
    if (rc) return rc;
    const long long total = HW * (C / 4);
    const unsigned blocks = (unsigned)min((long long)148 * 8, (total + 255) / 256);
    gn_apply_kernel<<<dim3(blocks, B), 256, 0, s>>>(x, partial, gamma, beta, y, HW, C, S, swish);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_softmax_rows(float *x, long long rows, int cols, void *stream) {
    SGAM_REQUIRE(x && rows > 0 && cols > 0, "softmax_rows: bad arguments");
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, cols);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}
