// Stage (i): forward splat, hole fill, extrapolation mask, inverse-depth coding; inverse warp; output packing.
//
// Bit-exact with the reference's CPU path (oracle/csrc/oracle.c): every fp32 operation uses an explicit
// round-to-nearest intrinsic in the reference's order, and this file is compiled with --fmad=false, so a
// fused multiply-add appears only where __fmaf_rn is written.
//
// Data flow (HBM):   src_depth --scatter--> winner[B,H,W] u64 (atomicMax, L2 resident: 512 KB @256^2)
//                    winner + gathered src_rgb/src_depth of the winners only --resolve--> x, mask
// The colours of points that lose the collision are never read, so DRAM traffic is below the algorithmic
// N*H*W*16 + H*W*17 bytes the roofline is quoted on.
#include "common.cuh"

namespace {

constexpr int kMaxSrc = 16;

struct Mats {          // per (b, n): inverse source intrinsics and rigid transform rows
    float Ki[9];
    float T[12];
};

__device__ __forceinline__ float transform_z_and_project(const float *Ki, const float *T, const float *Kt,
                                                         int i, int j, float d, float &px, float &py) {
    const float fj = (float)j, fi = (float)i;
    float X0 = __fmul_rn(dot3(Ki + 0, fj, fi, 1.0f), d);       // warp.py:36-40
    float X1 = __fmul_rn(dot3(Ki + 3, fj, fi, 1.0f), d);
    float X2 = __fmul_rn(dot3(Ki + 6, fj, fi, 1.0f), d);
    float Y0 = __fadd_rn(dot3(T + 0, X0, X1, X2), T[3]);       // warp.py:215
    float Y1 = __fadd_rn(dot3(T + 4, X0, X1, X2), T[7]);
    float Y2 = __fadd_rn(dot3(T + 8, X0, X1, X2), T[11]);
    float pz = dot3(Kt + 6, Y0, Y1, Y2);                        // warp.py:222-224
    px = __fdiv_rn(dot3(Kt + 0, Y0, Y1, Y2), pz);
    py = __fdiv_rn(dot3(Kt + 3, Y0, Y1, Y2), pz);
    return Y2;
}

__device__ __forceinline__ float transform_z(const float *Ki, const float *T, int i, int j, float d) {
    const float fj = (float)j, fi = (float)i;
    float X0 = __fmul_rn(dot3(Ki + 0, fj, fi, 1.0f), d);
    float X1 = __fmul_rn(dot3(Ki + 3, fj, fi, 1.0f), d);
    float X2 = __fmul_rn(dot3(Ki + 6, fj, fi, 1.0f), d);
    return __fadd_rn(dot3(T + 8, X0, X1, X2), T[11]);
}

// torch .long() on x86 (cvttss2si): truncation; NaN / out of range never pass the bounds test.
__device__ __forceinline__ bool pixel_index(float v, int limit, int &out) {
    float r = __fadd_rn(v, 0.5f);                               // warp.py:225
    if (!(r > -1.0f && r < 2147483000.0f)) return false;       // r in (-1, 0) truncates to 0 like the reference
    int t = (int)r;
    out = t;
    return t < limit;
}

template <int VEC>
__global__ void __launch_bounds__(256)
splat_scatter_kernel(const float *__restrict__ depth, const float *__restrict__ K_tgt,
                     const float *__restrict__ Kinv, const float *__restrict__ T,
                     unsigned long long *__restrict__ winner, uint8_t *__restrict__ inbounds,
                     int N, int H, int W, int policy) {
    const int bn = blockIdx.y, b = bn / N, n = bn - b * N;
    const int HW = H * W;
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p0 >= HW) return;
    float Ki[9], Tm[12], Kt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Ki[k] = __ldg(Kinv + 9 * bn + k); Kt[k] = __ldg(K_tgt + 9 * b + k); }
#pragma unroll
    for (int k = 0; k < 12; ++k) Tm[k] = __ldg(T + 16 * bn + k);
    float d[VEC];
    if (VEC == 4) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(depth + (size_t)bn * HW + p0));
        d[0] = v.x; d[1 % VEC] = v.y; d[2 % VEC] = v.z; d[3 % VEC] = v.w;
    } else {
        d[0] = __ldg(depth + (size_t)bn * HW + p0);
    }
    const int i = p0 / W, j0 = p0 - i * W;                      // VEC == 4 requires W % 4 == 0: same row
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int p = p0 + v;
        float px, py;
        float z = transform_z_and_project(Ki, Tm, Kt, i, j0 + v, d[v], px, py);
        int x, y;
        bool ok = pixel_index(px, W, x);
        ok = pixel_index(py, H, y) && ok;                       // warp.py:232
        if (inbounds) inbounds[(size_t)b * HW * N + (size_t)p * N + n] = ok ? 1 : 0;
        if (ok) {
            unsigned long long key = (unsigned long long)((unsigned)(p * N + n) + 1u);
            if (policy == SGAM_SPLAT_ZMIN)
                key |= (unsigned long long)(~float_orderable(__fadd_rn(z, 0.0f))) << 32;
            atomicMax(winner + (size_t)b * HW + (size_t)y * W + x, key);
        }
    }
}

// exact lower median of 9 with torch.median's NaN rule (rare path: holes and non-finite neighbourhoods)
__device__ float median9_exact(float *v) {
    for (int a = 0; a < 9; ++a)
        if (v[a] != v[a]) return v[a];
    for (int a = 1; a < 9; ++a) {
        float x = v[a];
        int b = a - 1;
        while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; --b; }
        v[b + 1] = x;
    }
    return v[4];
}

__device__ __forceinline__ float depth_code(float depth, float m, int dataset) {
    // model.py:210-229; python doubles are folded to fp32 scalars by torch's TensorIterator
    float w;
    if (dataset == SGAM_DATASET_CLEVR) {
        const float c16 = (float)(1.0 / 16), s = (float)(1.0 / 7 - 1.0 / 16);
        float d = (depth < 1e-7f) ? 1e-7f : depth;
        w = __fdiv_rn(1.0f, d);
        w = __fdiv_rn(__fsub_rn(w, c16), s);
    } else {
        const float c = (float)(1.0 / 14.765625), s = (float)(1.0 / 10.099975586 - 1.0 / 14.765625);
        w = __fdiv_rn(1.0f, __fadd_rn(depth, 10.0f));
        w = __fdiv_rn(__fsub_rn(w, c), s);
    }
    w = __fsub_rn(__fmul_rn(2.0f, w), 1.0f);
    return __fadd_rn(__fmul_rn(w, 1.0f - m), __fmul_rn(-2.0f, m));
}

constexpr int TX = 32, TY = 8;

__global__ void __launch_bounds__(TX * TY)
splat_resolve_kernel(const unsigned long long *__restrict__ winner, const float *__restrict__ src_rgb,
                     long long cs, long long ps, const float *__restrict__ depth,
                     const float *__restrict__ Kinv, const float *__restrict__ T, int N, int H, int W,
                     int dataset, float *__restrict__ x, uint8_t *__restrict__ mask,
                     float *__restrict__ merge_depth, float *__restrict__ proj) {
    __shared__ float tile[4][TY + 2][TX + 2];
    __shared__ Mats mats[kMaxSrc];
    const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int HW = H * W, tid = threadIdx.x;
    for (int e = tid; e < N * 21; e += TX * TY) {
        int n = e / 21, k = e - n * 21;
        float v = (k < 9) ? Kinv[9 * (b * N + n) + k] : T[16 * (b * N + n) + (k - 9)];
        if (k < 9) mats[n].Ki[k] = v; else mats[n].T[k - 9] = v;
    }
    __syncthreads();
    for (int e = tid; e < (TY + 2) * (TX + 2); e += TX * TY) {
        const int ly = e / (TX + 2), lx = e - ly * (TX + 2);
        const int gy = y0 + ly - 1, gx = x0 + lx - 1;
        float r = 0.f, g = 0.f, bl = 0.f, z = 0.f;                       // zero padding (warp.py:334-338)
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const unsigned lo = (unsigned)(winner[(size_t)b * HW + (size_t)gy * W + gx] & 0xffffffffull);
            if (lo) {
                const unsigned order = lo - 1u;
                const int p = order / N, n = order - p * N;
                const size_t bn = (size_t)b * N + n;
                const float *c = src_rgb + bn * 3 * HW + (size_t)p * ps;
                r = __ldg(c); g = __ldg(c + cs); bl = __ldg(c + 2 * cs);
                const int i = p / W, j = p - i * W;
                z = transform_z(mats[n].Ki, mats[n].T, i, j, __ldg(depth + bn * HW + p));
            }
        }
        tile[0][ly][lx] = r; tile[1][ly][lx] = g; tile[2][ly][lx] = bl; tile[3][ly][lx] = z;
    }
    __syncthreads();
    const int tx = tid % TX, ty = tid / TX, gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    const size_t q = (size_t)gy * W + gx;
    float outv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float v = tile[c][ty + 1][tx + 1];
        bool slow = (v == 0.0f);
        float nb[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            nb[k] = tile[c][ty + k / 3][tx + k % 3];
            slow = slow || !(fabsf(nb[k]) <= 3.0e38f);
        }
        if (slow) {                                                     // warp.py:271-279
            const float med = median9_exact(nb);
            const float m = (v == 0.0f) ? 1.0f : 0.0f;
            outv[c] = __fadd_rn(__fmul_rn(m, med), __fmul_rn(1.0f - m, v));
        } else {
            outv[c] = v;
        }
        if (proj) proj[((size_t)b * 4 + c) * HW + q] = v;
    }
    const float md = outv[3];
    const bool hole = md <= 0.0f;                                       // warp.py:285
    x[((size_t)b * 4 + 0) * HW + q] = outv[0];
    x[((size_t)b * 4 + 1) * HW + q] = outv[1];
    x[((size_t)b * 4 + 2) * HW + q] = outv[2];
    x[((size_t)b * 4 + 3) * HW + q] = depth_code(md, hole ? 1.0f : 0.0f, dataset);
    mask[(size_t)b * HW + q] = hole ? 1 : 0;
    if (merge_depth) merge_depth[(size_t)b * HW + q] = md;
}

__global__ void __launch_bounds__(256)
median_blur3_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const float *src = in + (size_t)blockIdx.y * H * W;
    const int i = p / W, j = p - i * W;
    float nb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int y = i + k / 3 - 1, xx = j + k % 3 - 1;
        nb[k] = (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(src + (size_t)y * W + xx) : 0.0f;
    }
    out[(size_t)blockIdx.y * H * W + p] = median9_exact(nb);
}

__global__ void __launch_bounds__(256)
depth_code_kernel(const float *__restrict__ rgb, const float *__restrict__ depth, int HW, int dataset,
                  float *__restrict__ x, uint8_t *__restrict__ mask) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= HW) return;
    const float d = depth[(size_t)b * HW + p];
    const bool hole = d <= 0.0f;                                        // model.py:199
#pragma unroll
    for (int c = 0; c < 3; ++c) x[((size_t)b * 4 + c) * HW + p] = rgb[((size_t)b * 3 + c) * HW + p];
    x[((size_t)b * 4 + 3) * HW + p] = depth_code(d, hole ? 1.0f : 0.0f, dataset);
    mask[(size_t)b * HW + p] = hole ? 1 : 0;
}

__global__ void __launch_bounds__(256)
inverse_warp_kernel(const float *__restrict__ src_rgb, long long cs, long long ps,
                    const float *__restrict__ src_depth, const float *__restrict__ tgt_depth,
                    const float *__restrict__ Kinv_tgt, const float *__restrict__ proj, int N, int H, int W,
                    float *__restrict__ out, int32_t *__restrict__ best_src) {
    __shared__ float P[kMaxSrc][12];
    const int b = blockIdx.y, HW = H * W;
    for (int e = threadIdx.x; e < N * 12; e += blockDim.x) P[e / 12][e % 12] = proj[(size_t)b * N * 12 + e];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int i = p / W, j = p - i * W;
    const float fj = (float)j, fi = (float)i, d = __ldg(tgt_depth + (size_t)b * HW + p);
    float Ki[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Ki[k] = __ldg(Kinv_tgt + 9 * b + k);
    const float X0 = __fmul_rn(dot3(Ki + 0, fj, fi, 1.0f), d);          // inference_pipeline.py:619-632
    const float X1 = __fmul_rn(dot3(Ki + 3, fj, fi, 1.0f), d);
    const float X2 = __fmul_rn(dot3(Ki + 6, fj, fi, 1.0f), d);
    float zbuf = 99999.0f, r0 = 0.f, r1 = 0.f, r2 = 0.f;                // :721-722
    int best = -1;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1), hw = (float)W / 2.0f, hh = (float)H / 2.0f;
    for (int n = 0; n < N; ++n) {
        const size_t bn = (size_t)b * N + n;
        const float pc0 = __fadd_rn(dot3(&P[n][0], X0, X1, X2), P[n][3]);   // :640-646
        const float pc1 = __fadd_rn(dot3(&P[n][4], X0, X1, X2), P[n][7]);
        const float Z = __fadd_rn(dot3(&P[n][8], X0, X1, X2), P[n][11]);
        const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(pc0, Z)), wm1), 1.0f);   // :655-657
        const float yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(pc1, Z)), hm1), 1.0f);
        // grid_sample(nearest, align_corners=False, zeros) (:707): ix = (x+1)*(W/2) - 0.5, half-to-even
        const float rx = rintf(__fsub_rn(__fmul_rn(__fadd_rn(xn, 1.0f), hw), 0.5f));
        const float ry = rintf(__fsub_rn(__fmul_rn(__fadd_rn(yn, 1.0f), hh), 0.5f));
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (rx >= 0.0f && rx < (float)W && ry >= 0.0f && ry < (float)H) {
            const size_t q = (size_t)ry * W + (size_t)rx;
            const float *c = src_rgb + bn * 3 * HW + q * ps;
            s0 = __fadd_rn(__ldg(c), 2.0f); s1 = __fadd_rn(__ldg(c + cs), 2.0f); s2 = __fadd_rn(__ldg(c + 2 * cs), 2.0f);
        }
        const float diff = fabsf(__fsub_rn(Z, __ldg(src_depth + bn * HW + p)));   // :698 (unsampled source depth)
        const float sum = __fadd_rn(__fadd_rn(s0, s1), s2);
        if (diff < zbuf && Z >= 0.0f && sum > 0.0f) {                            // :731-737
            zbuf = diff; best = n;
            r0 = __fsub_rn(s0, 2.0f); r1 = __fsub_rn(s1, 2.0f); r2 = __fsub_rn(s2, 2.0f);
        }
    }
    out[((size_t)b * 3 + 0) * HW + p] = r0;
    out[((size_t)b * 3 + 1) * HW + p] = r1;
    out[((size_t)b * 3 + 2) * HW + p] = r2;
    if (best_src) best_src[(size_t)b * HW + p] = best;
}

__global__ void __launch_bounds__(256)
frame_outputs_kernel(const float *__restrict__ dec, int HW, int dataset, uint8_t *__restrict__ rgb,
                     float *__restrict__ depth, float *__restrict__ src_rgb) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= HW) return;
    const float *d = dec + (size_t)b * 4 * HW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                       // inference_pipeline.py:898-901
        float v = __fmul_rn(__fdiv_rn(__fadd_rn(d[(size_t)c * HW + p], 1.0f), 2.0f), 255.0f);
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        const uint8_t u = (uint8_t)v;
        rgb[((size_t)b * HW + p) * 3 + c] = u;
        if (src_rgb) src_rgb[((size_t)b * HW + p) * 3 + c] = (float)((double)u / 127.5 - 1.0);   // :534
    }
    const float v = __fdiv_rn(__fadd_rn(d[(size_t)3 * HW + p], 1.0f), 2.0f);  // :906-911
    float o;
    if (dataset == SGAM_DATASET_CLEVR) {
        const float c16 = (float)(1.0 / 16), s = (float)(1.0 / 7 - 1.0 / 16);
        o = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(v, s), c16));
    } else {
        const float c = (float)(1.0 / 14.765625), s = (float)(1.0 / 10.099975586 - 1.0 / 14.765625);
        o = __fsub_rn(__fdiv_rn(1.0f, __fadd_rn(__fmul_rn(v, s), c)), 10.0f);
    }
    depth[(size_t)b * HW + p] = o;
}

}  // namespace

extern "C" size_t sgam_splat_workspace_bytes(int B, int H, int W) { return (size_t)B * H * W * sizeof(unsigned long long); }

extern "C" int sgam_splat_forward(const float *src_rgb, long long rgb_cs, long long rgb_ps, const float *src_depth,
                                  const float *K_tgt, const float *Kinv_src, const float *T_src2tgt, int B, int N,
                                  int H, int W, int policy, int dataset, void *winner, float *x, uint8_t *mask,
                                  float *merge_depth, float *proj, uint8_t *inbounds, void *stream) {
    SGAM_REQUIRE(src_rgb && src_depth && K_tgt && Kinv_src && T_src2tgt && winner && x && mask, "splat: null pointer");
    SGAM_REQUIRE(B > 0 && N > 0 && N <= kMaxSrc && H > 0 && W > 0, "splat: bad shape B=%d N=%d H=%d W=%d (N <= %d)", B, N, H, W, kMaxSrc);
    SGAM_REQUIRE((long long)H * W * N < 0xffffffffll, "splat: H*W*N must fit 32 bits");
    SGAM_REQUIRE(policy == SGAM_SPLAT_LAST_WRITER || policy == SGAM_SPLAT_ZMIN, "splat: bad policy %d", policy);
    SGAM_REQUIRE(dataset == SGAM_DATASET_CLEVR || dataset == SGAM_DATASET_GOOGLE_EARTH, "splat: bad dataset %d", dataset);
    cudaStream_t s = (cudaStream_t)stream;
    const int HW = H * W;
    SGAM_CUDA_OK(cudaMemsetAsync(winner, 0, sgam_splat_workspace_bytes(B, H, W), s));
    unsigned long long *win = (unsigned long long *)winner;
    if (W % 4 == 0 && ((uintptr_t)src_depth % 16) == 0) {
        dim3 grid(cdiv(HW / 4, 256), B * N);
        splat_scatter_kernel<4><<<grid, 256, 0, s>>>(src_depth, K_tgt, Kinv_src, T_src2tgt, win, inbounds, N, H, W, policy);
    } else {
        dim3 grid(cdiv(HW, 256), B * N);
        splat_scatter_kernel<1><<<grid, 256, 0, s>>>(src_depth, K_tgt, Kinv_src, T_src2tgt, win, inbounds, N, H, W, policy);
    }
    SGAM_LAUNCH_OK();
    dim3 grid2(cdiv(W, TX), cdiv(H, TY), B);
    splat_resolve_kernel<<<grid2, TX * TY, 0, s>>>(win, src_rgb, rgb_cs, rgb_ps, src_depth, Kinv_src, T_src2tgt, N, H, W,
                                                   dataset, x, mask, merge_depth, proj);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_median_blur3(const float *in, float *out, int planes, int H, int W, void *stream) {
    SGAM_REQUIRE(in && out && planes > 0 && H > 0 && W > 0, "median_blur3: bad arguments");
    dim3 grid(cdiv((long long)H * W, 256), planes);
    median_blur3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, H, W);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_depth_code(const float *rgb, const float *depth, int B, int H, int W, int dataset, float *x,
                               uint8_t *mask, void *stream) {
    SGAM_REQUIRE(rgb && depth && x && mask && B > 0 && H > 0 && W > 0, "depth_code: bad arguments");
    SGAM_REQUIRE(dataset == SGAM_DATASET_CLEVR || dataset == SGAM_DATASET_GOOGLE_EARTH, "depth_code: bad dataset %d", dataset);
    dim3 grid(cdiv((long long)H * W, 256), B);
    depth_code_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, depth, H * W, dataset, x, mask);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_inverse_warp(const float *src_rgb, long long rgb_cs, long long rgb_ps, const float *src_depth,
                                 const float *tgt_depth, const float *Kinv_tgt, const float *proj, int B, int N, int H,
                                 int W, float *out, int32_t *best_src, void *stream) {
    SGAM_REQUIRE(src_rgb && src_depth && tgt_depth && Kinv_tgt && proj && out, "inverse_warp: null pointer");
    SGAM_REQUIRE(B > 0 && N > 0 && N <= kMaxSrc && H > 0 && W > 0, "inverse_warp: bad shape");
    dim3 grid(cdiv((long long)H * W, 256), B);
    inverse_warp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_rgb, rgb_cs, rgb_ps, src_depth, tgt_depth, Kinv_tgt,
                                                                 proj, N, H, W, out, best_src);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_frame_outputs(const float *dec, int B, int H, int W, int dataset, uint8_t *rgb_u8, float *depth,
                                  float *src_rgb, void *stream) {
    SGAM_REQUIRE(dec && rgb_u8 && depth && B > 0 && H > 0 && W > 0, "frame_outputs: bad arguments");
    SGAM_REQUIRE(dataset == SGAM_DATASET_CLEVR || dataset == SGAM_DATASET_GOOGLE_EARTH, "frame_outputs: bad dataset %d", dataset);
    dim3 grid(cdiv((long long)H * W, 256), B);
    frame_outputs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dec, H * W, dataset, rgb_u8, depth, src_rgb);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}
