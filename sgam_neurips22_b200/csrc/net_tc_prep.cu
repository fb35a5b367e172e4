// Producers of the tensor-core path's split-bf16 operands (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)):
// fp32 -> split (+ fused nearest x2 up-sampling), GroupNorm apply (+ swish) -> split, row softmax -> split, and the
// finaliser of the GroupNorm statistics that tc_gemm_kernel's epilogue emits.  All HBM-bound, 8 B per element.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_split.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ producers of split bf16
// x fp32 [B, H, W, C] -> hi / lo bf16 [B, H<<up, W<<up, C] (nearest x2 up-sampling fused when up = 1).
// 8 channels per thread: two 16-byte loads, one 16-byte store per plane.
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo,
                  long long total_o, int H, int W, int CO, int up) {
    SGAM_PDL_PROLOGUE();
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total_o; e += (long long)gridDim.x * blockDim.x) {
        long long src = e;
        if (up) {
            const int co = (int)(e % CO);
            long long pix = e / CO;
            const int Wo = W * 2, Ho = H * 2;
            const int ox = (int)(pix % Wo); pix /= Wo;
            const int oy = (int)(pix % Ho); const long long b = pix / Ho;
            src = (((b * H + (oy >> 1)) * W + (ox >> 1)) * CO) + co;
        }
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(x) + 2 * src), v1 = __ldg(reinterpret_cast<const float4 *>(x) + 2 * src + 1);
        uint32_t h[4], l[4];
        split2(v0.x, v0.y, h[0], l[0]); split2(v0.z, v0.w, h[1], l[1]);
        split2(v1.x, v1.y, h[2], l[2]); split2(v1.z, v1.w, h[3], l[3]);
        reinterpret_cast<uint4 *>(hi)[e] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4 *>(lo)[e] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// GroupNorm apply (+ swish) with split-bf16 output; statistics come from gn_stats (net_simt.cu) partials.
// U independent 8-channel groups per thread and iteration; the first iteration's loads are issued BEFORE the statistics
// prologue so that the two global-memory round trips overlap (most launches of a step have exactly one iteration).
template <int U, bool FIXED_C>
__global__ void __launch_bounds__(256, U == 4 ? 3 : (U == 2 ? 4 : 5))
gn_apply_split_kernel(const float *__restrict__ x, const double *__restrict__ partial, const float *__restrict__ meanrstd,
                      const float *__restrict__ gamma, const float *__restrict__ beta, __nv_bfloat16 *__restrict__ hi,
                      __nv_bfloat16 *__restrict__ lo, long long HW, int C, int S, int swish) {
    SGAM_PDL_PROLOGUE();
    __shared__ float mean_s[32], rstd_s[32];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int CO = C / 8, cpg = C / 32;                 // 8 channels per thread (cpg >= 4: at most two groups)
    const long long total = HW * CO;
    const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)b * HW * C);
    uint4 *dh = reinterpret_cast<uint4 *>(hi + (size_t)b * HW * C), *dl = reinterpret_cast<uint4 *>(lo + (size_t)b * HW * C);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long e0 = (long long)blockIdx.x * blockDim.x + tid;
    float4 v0[U], v1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long e = e0 + u * stride;
        if (e < total) { v0[u] = __ldg(src + 2 * e); v1[u] = __ldg(src + 2 * e + 1); }
    }
    if (meanrstd && S > 0) {            // per-pixel-block sums from the producing conv's epilogue, S = blocks per image (<= 64)
        gn_mean_rstd_from_tiles(meanrstd, b, S, (double)HW * (C / 32), mean_s, rstd_s);
    } else if (meanrstd) {              // statistics already finalised by gn_finalize_kernel
        if (tid < 32) { mean_s[tid] = meanrstd[(b * 32 + tid) * 2]; rstd_s[tid] = meanrstd[(b * 32 + tid) * 2 + 1]; }
    } else {
        gn_mean_rstd_from_partials(partial, b, S, (double)HW * (C / 32), mean_s, rstd_s);
    }
    __syncthreads();
    // FIXED_C: the grid stride is a multiple of the channel-octet count (C = 128 .. 1024, powers of two): a thread keeps ONE
    // channel octet for its whole life and its affine parameters stay in registers
    constexpr bool fixed_c = FIXED_C;
    int c = (int)(e0 % CO) * 8;
    float mu0 = mean_s[c / cpg], rs0 = rstd_s[c / cpg], mu1 = mean_s[(c + 4) / cpg], rs1 = rstd_s[(c + 4) / cpg];
    float4 ga0 = __ldg(reinterpret_cast<const float4 *>(gamma + c)), ga1 = __ldg(reinterpret_cast<const float4 *>(gamma + c + 4));
    float4 be0 = __ldg(reinterpret_cast<const float4 *>(beta + c)), be1 = __ldg(reinterpret_cast<const float4 *>(beta + c + 4));
    while (e0 < total) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = e0 + u * stride;
            if (e >= total) break;
            if (!fixed_c) {
                c = (int)(e % CO) * 8;
                mu0 = mean_s[c / cpg]; rs0 = rstd_s[c / cpg]; mu1 = mean_s[(c + 4) / cpg]; rs1 = rstd_s[(c + 4) / cpg];
                ga0 = __ldg(reinterpret_cast<const float4 *>(gamma + c)); ga1 = __ldg(reinterpret_cast<const float4 *>(gamma + c + 4));
                be0 = __ldg(reinterpret_cast<const float4 *>(beta + c)); be1 = __ldg(reinterpret_cast<const float4 *>(beta + c + 4));
            }
            float o[8] = {(v0[u].x - mu0) * rs0 * ga0.x + be0.x, (v0[u].y - mu0) * rs0 * ga0.y + be0.y,
                          (v0[u].z - mu0) * rs0 * ga0.z + be0.z, (v0[u].w - mu0) * rs0 * ga0.w + be0.w,
                          (v1[u].x - mu1) * rs1 * ga1.x + be1.x, (v1[u].y - mu1) * rs1 * ga1.y + be1.y,
                          (v1[u].z - mu1) * rs1 * ga1.z + be1.z, (v1[u].w - mu1) * rs1 * ga1.w + be1.w};
            if (swish) {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = __fdividef(o[k], 1.0f + __expf(-o[k]));   // <= 3 ulp; keeps the kernel HBM-bound
            }
            uint32_t h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split2(o[2 * k], o[2 * k + 1], h[k], l[k]);
            dh[e] = make_uint4(h[0], h[1], h[2], h[3]);
            dl[e] = make_uint4(l[0], l[1], l[2], l[3]);
        }
        e0 += U * stride;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long e = e0 + u * stride;
            if (e < total) { v0[u] = __ldg(src + 2 * e); v1[u] = __ldg(src + 2 * e + 1); }
        }
    }
}

// row softmax of fp32 scores -> split-bf16 probabilities (model.py:181); cols % 8 == 0.
// VPT > 0: the row (cols <= 2048*VPT) is read ONCE into registers, 8 consecutive elements per thread per chunk
// (two 16-byte loads, one 16-byte store per plane); VPT = 0: three-pass fallback for very long rows.
template <int VPT>
__global__ void __launch_bounds__(256)
softmax_split_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int cols) {
    SGAM_PDL_PROLOGUE();
    __shared__ float sh[8];
    const float *row = x + (size_t)blockIdx.x * cols;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, oct = cols / 8;
    uint4 *dh = reinterpret_cast<uint4 *>(hi + (size_t)blockIdx.x * cols), *dl = reinterpret_cast<uint4 *>(lo + (size_t)blockIdx.x * cols);
    float v[VPT > 0 ? VPT : 1][8];
    float mx = -INFINITY;
    if (VPT > 0) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int c = threadIdx.x + i * 256;
            if (c < oct) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(row) + 2 * c), b4 = __ldg(reinterpret_cast<const float4 *>(row) + 2 * c + 1);
                v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b4.x; v[i][5] = b4.y; v[i][6] = b4.z; v[i][7] = b4.w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[i][k] = -INFINITY;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) mx = fmaxf(mx, v[i][k]);
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += 256) mx = fmaxf(mx, row[c]);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sh[warp] = mx;
    __syncthreads();
    mx = sh[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sh[w]);
    __syncthreads();
    float sum = 0.f;
    if (VPT > 0) {
#pragma unroll
        for (int i = 0; i < VPT; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[i][k] = expf(v[i][k] - mx); sum += v[i][k]; }
    } else {
        for (int c = threadIdx.x; c < cols; c += 256) sum += expf(row[c] - mx);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) sh[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += sh[w];
    const float inv = 1.0f / sum;
    if (VPT > 0) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int c = threadIdx.x + i * 256;
            if (c < oct) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) split2(v[i][2 * k] * inv, v[i][2 * k + 1] * inv, h[k], l[k]);
                dh[c] = make_uint4(h[0], h[1], h[2], h[3]);
                dl[c] = make_uint4(l[0], l[1], l[2], l[3]);
            }
        }
    } else {
        for (int c = threadIdx.x; c < oct; c += 256) {
            const float4 a = *reinterpret_cast<const float4 *>(row + 8 * c), b4 = *reinterpret_cast<const float4 *>(row + 8 * c + 4);
            uint32_t h[4], l[4];
            split2(expf(a.x - mx) * inv, expf(a.y - mx) * inv, h[0], l[0]); split2(expf(a.z - mx) * inv, expf(a.w - mx) * inv, h[1], l[1]);
            split2(expf(b4.x - mx) * inv, expf(b4.y - mx) * inv, h[2], l[2]); split2(expf(b4.z - mx) * inv, expf(b4.w - mx) * inv, h[3], l[3]);
            dh[c] = make_uint4(h[0], h[1], h[2], h[3]);
            dl[c] = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

// VQModel.encode stem for the tensor-core path: cat(x, mask) -> 1x1 conv 5 -> 4 (model.py:106-113), written as split
// bf16 NHWC with the 4 channels zero-padded to CP (64) so that encoder.conv_in (3x3, Cin = 4) runs as a K = 9*64 GEMM.
__global__ void __launch_bounds__(256)
stem_split_kernel(const float *__restrict__ x, const uint8_t *__restrict__ mask, const float *__restrict__ w,
                  const float *__restrict__ bias, int HW, int CP, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
    SGAM_PDL_PROLOGUE();
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
    uint32_t h[2], l[2];
    split2(o[0], o[1], h[0], l[0]);
    split2(o[2], o[3], h[1], l[1]);
    uint4 *dh = reinterpret_cast<uint4 *>(hi + ((size_t)b * HW + p) * CP), *dl = reinterpret_cast<uint4 *>(lo + ((size_t)b * HW + p) * CP);
    dh[0] = make_uint4(h[0], h[1], 0u, 0u);
    dl[0] = make_uint4(l[0], l[1], 0u, 0u);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int k = 1; k < CP / 8; ++k) { dh[k] = z; dl[k] = z; }
}

// reduce the per-pixel-block partial sums written by the GEMM epilogue: [B][tiles][32][2] fp32 -> mean, rstd.
// One WARP per (batch element, group): lane l adds tiles l, l + 32, ... in fp64, then a fixed butterfly combines the
// lanes (deterministic).  256 warps for 8 trajectories instead of 8 CTAs walking 64 tiles per thread: the kernel is a
// pure latency chain, so the shorter chain is what matters (7 us -> ~3 us, 48 launches per step).
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float *__restrict__ partial, float *__restrict__ meanrstd, int tiles, double count, int pairs) {
    SGAM_PDL_PROLOGUE();
    const int lane = threadIdx.x & 31, pair = blockIdx.x * 8 + (threadIdx.x >> 5);      // pair = b * 32 + g
    if (pair >= pairs) return;
    const int b = pair >> 5, g = pair & 31;
    double a = 0.0, q = 0.0;
    for (int t = lane; t < tiles; t += 32) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(partial + (((size_t)b * tiles + t) * 32 + g) * 2));
        a += (double)v.x; q += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) {
        const double mean = a / count;
        double var = q / count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        meanrstd[pair * 2] = (float)mean;
        meanrstd[pair * 2 + 1] = (float)(1.0 / sqrt(var + 1e-6));
    }
}


// ------------------------------------------------------------------------------------------------ decoder head
// decoder.norm_out + swish + decoder.conv_out (GroupNorm(32,128) -> x*sigmoid(x) -> 3x3 conv 128 -> 4, NCHW result;
// diffusionmodules/model.py:534-538) as ONE exact-fp32 kernel.  As a tensor-core GEMM this layer is hopeless: 4 output
// channels fill 3 % of an N = 128 tile, and the nine filter taps re-read the 128-channel operand nine times through L2
// (round 1: 87 us of GroupNorm-apply + 341 us of GEMM at 14 TFLOP/s, L2-bound).  Here a CTA stages the normalised,
// activated (8+2) x (16+2) x 128 halo tile in shared memory ONCE (zero padding applied after the activation, as the conv
// sees it) and reduces over the channels on the FP32 pipes: lane l owns channels 4l..4l+3 (one GroupNorm group); the
// eight warps form four PAIRS, a pair owns two rows of the tile and its two warps split the four output channels, so
// each lane keeps 9 x 4 x 2 filter taps in registers (72, not 144: twice the resident warps) and no cross-warp sum is
// needed.  A warp walks 2-row x 4-pixel segments: the 4 x 6 window is read once (24 conflict-free 16-byte loads for 576
// fused multiply-adds) and the sixteen per-lane partial sums are combined over the 32 lanes by recursive halving (16
// shuffles instead of 80).  4608 FMAs per output pixel: FP32-FMA bound.
constexpr int HD_TH = 8, HD_TW = 16, HD_C = 128;
constexpr int HD_PIX = (HD_TH + 2) * (HD_TW + 2);
constexpr int HD_SMEM = HD_PIX * HD_C * 4;

__global__ void __launch_bounds__(256, 2)
gn_head_conv_kernel(const float *__restrict__ x, const float *__restrict__ meanrstd, const float *__restrict__ gamma,
                    const float *__restrict__ beta, const float *__restrict__ w /* [9][128][4] */, const float *__restrict__ bias,
                    float *__restrict__ y /* [B][4][H][W] */, int H, int W) {
    SGAM_PDL_PROLOGUE();
    extern __shared__ float4 hs[];                                  // [(TH+2) * (TW+2)][32] float4 = 4 channels each
    const int b = blockIdx.z, y0 = blockIdx.y * HD_TH, x0 = blockIdx.x * HD_TW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- stage: normalise + swish every element of the halo tile once
    {
        const float mean = meanrstd[(b * 32 + lane) * 2], rstd = meanrstd[(b * 32 + lane) * 2 + 1];   // group = channel quad (128 / 32 = 4)
        const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma) + lane), be = __ldg(reinterpret_cast<const float4 *>(beta) + lane);
        const float4 *src = reinterpret_cast<const float4 *>(x) + (size_t)b * H * W * 32;
        // 180 halo pixels over 8 warps, in batches of 8 whose loads are all in flight before the first is used (a loop of
        // load -> transform -> store exposed one full DRAM latency per pixel: 460 us for the layer, ncu pass E)
        constexpr int BATCH = 8;
#pragma unroll 1
        for (int b0 = 0; b0 < (HD_PIX + 7) / 8; b0 += BATCH) {
            float4 v[BATCH];
            bool ok[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const int pix = warp + 8 * (b0 + i);
                const int gy = y0 - 1 + pix / (HD_TW + 2), gx = x0 - 1 + pix % (HD_TW + 2);
                ok[i] = pix < HD_PIX && gy >= 0 && gy < H && gx >= 0 && gx < W;
                v[i] = ok[i] ? __ldg(src + ((size_t)gy * W + gx) * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
                const int pix = warp + 8 * (b0 + i);
                float4 t = v[i];
                if (ok[i]) {
                    t.x = (t.x - mean) * rstd * ga.x + be.x; t.y = (t.y - mean) * rstd * ga.y + be.y;
                    t.z = (t.z - mean) * rstd * ga.z + be.z; t.w = (t.w - mean) * rstd * ga.w + be.w;
                    t.x = __fdividef(t.x, 1.0f + __expf(-t.x)); t.y = __fdividef(t.y, 1.0f + __expf(-t.y));
                    t.z = __fdividef(t.z, 1.0f + __expf(-t.z)); t.w = __fdividef(t.w, 1.0f + __expf(-t.w));
                }
                if (pix < HD_PIX) hs[pix * 32 + lane] = t;
            }
        }
    }
    // ---- this lane's filter taps: wr[tap][channel of the quad] = output channels 2 op, 2 op + 1 (op = which warp of the pair)
    const int pair = warp >> 1, op = warp & 1;
    float2 wr[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) wr[t][c] = __ldg(reinterpret_cast<const float2 *>(w) + ((t * HD_C + lane * 4 + c) * 2 + op));
    __syncthreads();
    const float b0v = __ldg(bias + 2 * op), b1v = __ldg(bias + 2 * op + 1);
    const size_t plane = (size_t)H * W;
    // pair -> tile rows 2 pair, 2 pair + 1; segments of 2 rows x 4 pixels
#pragma unroll 1
    for (int seg = 0; seg < HD_TW / 4; ++seg) {
        const int ly = 2 * pair, lx = seg * 4;
        float acc[16];                                              // [row 0..1][pixel 0..3][out 0..1]
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {                              // window row r feeds output row 0 as tap row r and output row 1 as tap row r - 1
            float4 xv[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) xv[j] = hs[((ly + r) * (HD_TW + 2) + lx + j) * 32 + lane];
#pragma unroll
            for (int orow = 0; orow < 2; ++orow) {
                const int kh = r - orow;
                if (kh < 0 || kh > 2) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int t = kh * 3 + kw;
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float4 v = xv[px + kw];
                        float *a = acc + (orow * 4 + px) * 2;
                        a[0] = fmaf(v.x, wr[t][0].x, a[0]); a[1] = fmaf(v.x, wr[t][0].y, a[1]);
                        a[0] = fmaf(v.y, wr[t][1].x, a[0]); a[1] = fmaf(v.y, wr[t][1].y, a[1]);
                        a[0] = fmaf(v.z, wr[t][2].x, a[0]); a[1] = fmaf(v.z, wr[t][2].y, a[1]);
                        a[0] = fmaf(v.w, wr[t][3].x, a[0]); a[1] = fmaf(v.w, wr[t][3].y, a[1]);
                    }
                }
            }
        }
        // recursive halving over the lanes: after the step with distance d a lane keeps the half of its values selected by
        // its own bit d; value index v = (bit4, bit3, bit2, bit1) ends up complete in the lanes carrying those bits
        float r8[8], r4[4], r2[2], r1;
        {
            const bool up = lane & 16;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float send = up ? acc[i] : acc[i + 8], keep = up ? acc[i + 8] : acc[i]; r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
        }
        {
            const bool up = lane & 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float send = up ? r8[i] : r8[i + 4], keep = up ? r8[i + 4] : r8[i]; r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
        }
        {
            const bool up = lane & 4;
#pragma unroll
            for (int i = 0; i < 2; ++i) { const float send = up ? r4[i] : r4[i + 2], keep = up ? r4[i + 2] : r4[i]; r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
        }
        {
            const bool up = lane & 2;
            const float send = up ? r2[0] : r2[1], keep = up ? r2[1] : r2[0];
            r1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
        if ((lane & 1) == 0) {
            const int v = lane >> 1, orow = v >> 3, px = (v >> 1) & 3, o = v & 1;      // v = 8 bit4 + 4 bit3 + 2 bit2 + bit1
            const int gy = y0 + ly + orow, gx = x0 + lx + px;
            if (gy < H && gx < W) y[((size_t)b * 4 + 2 * op + o) * plane + (size_t)gy * W + gx] = r1 + (o ? b1v : b0v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ encoder stem
// VQModel.encode's stem (cat(x, mask) -> 1x1 conv 5 -> 4, model.py:106-113) + encoder.conv_in (3x3, 4 -> 128,
// diffusionmodules/model.py:370) + the GroupNorm partial sums of the result, as ONE exact-fp32 kernel.  Round 1 ran
// conv_in on the tensor cores by zero-padding its 4 input channels to 64 (K = 576 for 36 real taps: 94 % of the MMA work
// and of the 537 MB padded operand were zeros; 59 us + 169 us at 8 trajectories).  The layer is 4608 FMAs per pixel --
// FP32-pipe work: a CTA owns one 128-pixel block of the statistics layout (BW x BH pixels), stages the 4-channel stem
// values of its halo in shared memory (zero outside the image: the conv pads the STEM OUTPUT), thread c owns output channel
// c with its 36 taps in registers and walks the block in segments of four pixels (the 3 x 6 window is read once with
// broadcast 16-byte loads), writes 512-byte rows, and the per-channel sums are folded to the 32 groups with two shuffles.
// 147 us at 8 trajectories = 2.4 G FMAs at 64 lanes per clock and SM: three-register FFMAs issue every other cycle on this
// part, and the packed FFMA2 form measured the same (profiles/r2_rejected_experiments.txt) -- this is the FP32-pipe roofline.
__global__ void __launch_bounds__(128)
stem_conv_in_kernel(const float *__restrict__ x, const uint8_t *__restrict__ mask, const float *__restrict__ w1, const float *__restrict__ b1,
                    const float *__restrict__ w3 /* [128][9*4] */, const float *__restrict__ b3, float *__restrict__ y,
                    float *__restrict__ stats, int H, int W, int BW, int BH, int tiles_x, int tiles) {
    SGAM_PDL_PROLOGUE();
    extern __shared__ float4 st[];                                  // [(BH+2)][(BW+2)] stem values (4 channels)
    const int b = blockIdx.y, tile = blockIdx.x, ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * BH, x0 = tx * BW, c = threadIdx.x, SW = BW + 2;
    const size_t HW = (size_t)H * W;
    for (int p = threadIdx.x; p < (BH + 2) * SW; p += 128) {
        const int gy = y0 - 1 + p / SW, gx = x0 - 1 + p % SW;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t pix = (size_t)gy * W + gx;
            float in[5];
#pragma unroll
            for (int k = 0; k < 4; ++k) in[k] = __ldg(x + ((size_t)b * 4 + k) * HW + pix);
            in[4] = mask ? (mask[(size_t)b * HW + pix] ? 1.0f : 0.0f) : 0.0f;
            float r[4];
#pragma unroll
            for (int co = 0; co < 4; ++co) {
                float a = 0.0f;
#pragma unroll
                for (int k = 0; k < 5; ++k) a = fmaf(in[k], __ldg(w1 + co * 5 + k), a);
                r[co] = a + __ldg(b1 + co);
            }
            o = make_float4(r[0], r[1], r[2], r[3]);
        }
        st[p] = o;
    }
    float4 wr[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t] = __ldg(reinterpret_cast<const float4 *>(w3) + c * 9 + t);
    const float bias = __ldg(b3 + c);
    __syncthreads();
    float s_acc = 0.0f, q_acc = 0.0f;
    for (int ly = 0; ly < BH; ++ly) {
        const int gy = y0 + ly;
        if (gy >= H) break;
#pragma unroll 1
        for (int lx = 0; lx < BW; lx += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                float4 v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) v[j] = st[(ly + kh) * SW + lx + j];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 wt = wr[kh * 3 + kw];
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        acc[px] = fmaf(v[px + kw].x, wt.x, acc[px]); acc[px] = fmaf(v[px + kw].y, wt.y, acc[px]);
                        acc[px] = fmaf(v[px + kw].z, wt.z, acc[px]); acc[px] = fmaf(v[px + kw].w, wt.w, acc[px]);
                    }
                }
            }
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int gx = x0 + lx + px;
                if (gx < W) {
                    const float o = acc[px] + bias;
                    y[(((size_t)b * H + gy) * W + gx) * 128 + c] = o;
                    s_acc += o; q_acc = fmaf(o, o, q_acc);
                }
            }
        }
    }
    if (stats) {                                                    // channels 4g .. 4g+3 = group g: lanes 4g .. 4g+3
        s_acc += __shfl_xor_sync(0xffffffffu, s_acc, 1); q_acc += __shfl_xor_sync(0xffffffffu, q_acc, 1);
        s_acc += __shfl_xor_sync(0xffffffffu, s_acc, 2); q_acc += __shfl_xor_sync(0xffffffffu, q_acc, 2);
        if ((c & 3) == 0) {
            float *dst = stats + (((size_t)b * tiles + tile) * 32 + (c >> 2)) * 2;
            dst[0] = s_acc; dst[1] = q_acc;
        }
    }
}

}  // namespace

// CTAs per image of gn_apply_split_kernel<U>: each thread should own U 8-channel groups (one unrolled iteration) and the
// whole launch should not exceed ~8 CTAs per SM; small tensors get fewer CTAs rather than one group per thread
static unsigned gn_apply_blocks(long long total, int B, int U) {
    long long want = (total + 256 * U - 1) / (256 * U);
    const long long cap = ((long long)148 * 8 + B - 1) / B;
    if (want > cap) want = cap;
    return (unsigned)(want < 1 ? 1 : want);
}

// groups per thread; experiments: SGAM_GN_APPLY_U=1|2|4
static int gn_apply_unroll(long long total, int B) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("SGAM_GN_APPLY_U");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    (void)total; (void)B;
    return 2;                           // profiles/r2_gn_apply_sweep.txt: 2 is best or tied at every shape of the step
}

static cudaError_t gn_apply_launch(cudaStream_t s, const float *x, const double *partial, const float *meanrstd, const float *gamma,
                                   const float *beta, __nv_bfloat16 *hi, __nv_bfloat16 *lo, int B, long long HW, int C, int S, int swish) {
    const long long total = HW * (C / 8);
    const int U = gn_apply_unroll(total, B);
    const dim3 grid(gn_apply_blocks(total, B, U), B);
    const bool fixed = ((long long)grid.x * 256) % (C / 8) == 0;
#define SGAM_GN_APPLY(UU, FF) sgam_launch_pdl(SGAM_PDL_NORM, gn_apply_split_kernel<UU, FF>, grid, dim3(256), 0, s, x, partial, meanrstd, gamma, beta, hi, lo, HW, C, S, swish)
    if (U == 1) return fixed ? SGAM_GN_APPLY(1, true) : SGAM_GN_APPLY(1, false);
    if (U == 2) return fixed ? SGAM_GN_APPLY(2, true) : SGAM_GN_APPLY(2, false);
    return fixed ? SGAM_GN_APPLY(4, true) : SGAM_GN_APPLY(4, false);
#undef SGAM_GN_APPLY
}

// gn_stats launcher lives in net_simt.cu
int sgam_gn_stats_launch(const float *x, double *partial, int B, long long HW, int C, cudaStream_t s);

extern "C" int sgam_stem_conv_split(const float *x, const uint8_t *mask, const float *w, const float *bias, int B, int H, int W,
                                    int Cpad, void *hi, void *lo, void *stream) {
    SGAM_REQUIRE(x && w && bias && hi && lo && B > 0 && H > 0 && W > 0 && Cpad >= 8 && Cpad % 8 == 0, "stem_conv_split: bad arguments");
    dim3 grid(cdiv((long long)H * W, 256), B);
    SGAM_CUDA_OK(sgam_launch_pdl(SGAM_PDL_MISC, stem_split_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, mask, w, bias, H * W, Cpad, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo));
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_split_bf16(const float *x, void *hi, void *lo, int B, int H, int W, int C, int upsample, void *stream) {
    SGAM_REQUIRE(x && hi && lo && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "split_bf16: C must be a multiple of 8");
    const long long total_o = (long long)B * (H << upsample) * (W << upsample) * (C / 8);
    const unsigned blocks = (unsigned)min((long long)148 * 16, (total_o + 255) / 256);
    SGAM_PDL_LAUNCH(SGAM_PDL_MISC, split_bf16_kernel, blocks, 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo, total_o, H, W, C / 8, upsample);
    return SGAM_OK;
}

extern "C" int sgam_groupnorm_split(const float *x, const float *gamma, const float *beta, void *hi, void *lo, double *partial,
                                    int B, long long HW, int C, int swish, void *stream) {
    SGAM_REQUIRE(x && gamma && beta && hi && lo && partial, "groupnorm_split: null pointer");
    SGAM_REQUIRE(B > 0 && HW > 0 && C % 128 == 0 && C <= 1024, "groupnorm_split: C=%d must be a multiple of 128 (<= 1024)", C);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = sgam_gn_stats_launch(x, partial, B, HW, C, s);
    if (rc) return rc;
    ++g_sgam_launches;
    SGAM_CUDA_OK(gn_apply_launch(s, x, partial, nullptr, gamma, beta, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo, B, HW, C, sgam_gn_splits(HW), swish));
    return SGAM_OK;
}

extern "C" int sgam_groupnorm_split_apply(const float *x, const float *gamma, const float *beta, void *hi, void *lo,
                                          const double *partial, int B, long long HW, int C, int swish, void *stream) {
    SGAM_REQUIRE(x && gamma && beta && hi && lo && partial, "groupnorm_split_apply: null pointer");
    SGAM_REQUIRE(B > 0 && HW > 0 && C % 128 == 0 && C <= 1024, "groupnorm_split_apply: C=%d must be a multiple of 128 (<= 1024)", C);
    ++g_sgam_launches;
    SGAM_CUDA_OK(gn_apply_launch((cudaStream_t)stream, x, partial, nullptr, gamma, beta, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo, B, HW, C, sgam_gn_splits(HW), swish));
    return SGAM_OK;
}

extern "C" long long sgam_tc_gn_partial_floats(int B, int Ho, int Wo) {
    const int BW = Wo >= 128 ? 128 : Wo, BH = 128 / BW;
    const long long tiles = (long long)cdiv(Wo, BW) * cdiv(Ho, BH);
    return (long long)B * tiles * 64 + (long long)B * 64;           // partial sums, then [B][32][mean, rstd]
}

extern "C" int sgam_groupnorm_split_fused(const float *x, const float *gamma, const float *beta, void *hi, void *lo,
                                          float *gn_partial, int B, int Ho, int Wo, int C, int swish, void *stream) {
    SGAM_REQUIRE(x && gamma && beta && hi && lo && gn_partial, "groupnorm_split_fused: null pointer");
    SGAM_REQUIRE(B > 0 && Ho > 0 && Wo > 0 && C % 128 == 0 && C <= 512, "groupnorm_split_fused: C=%d must be 128, 256, 384 or 512", C);
    cudaStream_t s = (cudaStream_t)stream;
    const int BW = Wo >= 128 ? 128 : Wo, BH = 128 / BW;
    const int tiles = cdiv(Wo, BW) * cdiv(Ho, BH);
    const long long HW = (long long)Ho * Wo;
    if (tiles <= 64) {                  // few blocks per image: every CTA of the apply kernel finalises the sums itself (one launch less)
        ++g_sgam_launches;
        SGAM_CUDA_OK(gn_apply_launch(s, x, nullptr, gn_partial, gamma, beta, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo, B, HW, C, tiles, swish));
        return SGAM_OK;
    }
    float *meanrstd = gn_partial + (long long)B * tiles * 64;
    SGAM_PDL_LAUNCH(SGAM_PDL_NORM, gn_finalize_kernel, cdiv(B * 32, 8), 256, 0, s, gn_partial, meanrstd, tiles, (double)HW * (C / 32), B * 32);
    ++g_sgam_launches;
    SGAM_CUDA_OK(gn_apply_launch(s, x, nullptr, meanrstd, gamma, beta, (__nv_bfloat16 *)hi, (__nv_bfloat16 *)lo, B, HW, C, 0, swish));
    return SGAM_OK;
}

// mean / rstd of every (image, group) from the per-pixel-block sums a conv epilogue left in gn_partial (layout of
// sgam_tc_gn_partial_floats: the [B][32][2] result goes behind the partial sums); used by the fused GroupNorm + conv kernel (net_tc3.cu)
int sgam_gn_finalize_launch(float *gn_partial, int B, int H, int W, int C, cudaStream_t s, float **meanrstd_out) {
    const int BW = W >= 128 ? 128 : W, BH = 128 / BW;
    const int tiles = cdiv(W, BW) * cdiv(H, BH);
    float *meanrstd = gn_partial + (long long)B * tiles * 64;
    SGAM_PDL_LAUNCH(SGAM_PDL_NORM, gn_finalize_kernel, cdiv(B * 32, 8), 256, 0, s, gn_partial, meanrstd, tiles, (double)H * W * (C / 32), B * 32);
    *meanrstd_out = meanrstd;
    return SGAM_OK;
}

extern "C" int sgam_softmax_split(const float *x, void *hi, void *lo, long long rows, int cols, void *stream) {
    SGAM_REQUIRE(x && hi && lo && rows > 0 && cols > 0 && cols % 8 == 0, "softmax_split: cols must be a multiple of 8");
    cudaStream_t s = (cudaStream_t)stream;
    __nv_bfloat16 *h = (__nv_bfloat16 *)hi, *l = (__nv_bfloat16 *)lo;
    if (cols <= 2048) SGAM_PDL_LAUNCH(SGAM_PDL_SOFTMAX, softmax_split_kernel<1>, (unsigned)rows, 256, 0, s, x, h, l, cols);
    else if (cols <= 4096) SGAM_PDL_LAUNCH(SGAM_PDL_SOFTMAX, softmax_split_kernel<2>, (unsigned)rows, 256, 0, s, x, h, l, cols);
    else if (cols <= 16384) SGAM_PDL_LAUNCH(SGAM_PDL_SOFTMAX, softmax_split_kernel<8>, (unsigned)rows, 256, 0, s, x, h, l, cols);
    else SGAM_PDL_LAUNCH(SGAM_PDL_SOFTMAX, softmax_split_kernel<0>, (unsigned)rows, 256, 0, s, x, h, l, cols);
    return SGAM_OK;
}


// decoder.norm_out + swish + decoder.conv_out in one fp32 kernel (see gn_head_conv_kernel).  x fp32 NHWC [B,H,W,128] with
// the GroupNorm partial sums its producing conv emitted (gn_partial, sgam_tc_gn_partial_floats(B,H,W) floats);
// w_t [9][128][4] fp32 (tap, input channel, output channel); y [B,4,H,W].
extern "C" int sgam_gn_head_conv(const float *x, const float *gamma, const float *beta, float *gn_partial, const float *w_t,
                                 const float *bias, float *y, int B, int H, int W, int C, int Cout, void *stream) {
    SGAM_REQUIRE(x && gamma && beta && gn_partial && w_t && bias && y, "gn_head_conv: null pointer");
    SGAM_REQUIRE(B > 0 && H > 0 && W > 0 && C == HD_C && Cout == 4, "gn_head_conv: needs 128 input and 4 output channels (C=%d Cout=%d)", C, Cout);
    cudaStream_t s = (cudaStream_t)stream;
    const int BW = W >= 128 ? 128 : W, BH = 128 / BW;
    const int tiles = cdiv(W, BW) * cdiv(H, BH);
    float *meanrstd = gn_partial + (long long)B * tiles * 64;
    SGAM_PDL_LAUNCH(SGAM_PDL_NORM, gn_finalize_kernel, cdiv(B * 32, 8), 256, 0, s, gn_partial, meanrstd, tiles, (double)H * W * (C / 32), B * 32);
    static bool configured = false;
    if (!configured) {
        SGAM_CUDA_OK(cudaFuncSetAttribute(gn_head_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_SMEM));
        configured = true;
    }
    SGAM_PDL_LAUNCH(SGAM_PDL_NORM, gn_head_conv_kernel, dim3(cdiv(W, HD_TW), cdiv(H, HD_TH), B), 256, HD_SMEM, s, x, meanrstd, gamma, beta, w_t, bias, y, H, W);
    return SGAM_OK;
}

// Stem + encoder.conv_in + GroupNorm partial sums in one fp32 kernel (see stem_conv_in_kernel).  x [B,4,H,W] NCHW, mask
// [B,H,W] u8 or NULL, w1 [4,5] / b1 [4] (the 1x1 stem), w3 [128, 9*4] K-major (kh, kw, ci) / b3 [128]; y fp32 NHWC
// [B,H,W,128]; gn_partial (sgam_tc_gn_partial_floats(B,H,W) floats) or NULL.
extern "C" int sgam_stem_conv_in(const float *x, const uint8_t *mask, const float *w1, const float *b1, const float *w3,
                                 const float *b3, float *y, float *gn_partial, int B, int H, int W, int Cout, void *stream) {
    SGAM_REQUIRE(x && w1 && b1 && w3 && b3 && y, "stem_conv_in: null pointer");
    SGAM_REQUIRE(B > 0 && H > 0 && W >= 4 && Cout == 128, "stem_conv_in: needs 128 output channels and W >= 4 (W=%d Cout=%d)", W, Cout);
    SGAM_REQUIRE(W % 128 == 0 || ((W & (W - 1)) == 0 && W <= 128), "stem_conv_in: W must be a power of two <= 128 or a multiple of 128");
    const int BW = W >= 128 ? 128 : W, BH = 128 / BW;
    const int tiles_x = cdiv(W, BW), tiles = tiles_x * cdiv(H, BH);
    const size_t smem = (size_t)(BH + 2) * (BW + 2) * sizeof(float4);
    SGAM_PDL_LAUNCH(SGAM_PDL_MISC, stem_conv_in_kernel, dim3(tiles, B), 128, smem, (cudaStream_t)stream, x, mask, w1, b1, w3, b3, y, gn_partial, H, W,
                    BW, BH, tiles_x, tiles);
    return SGAM_OK;
}
