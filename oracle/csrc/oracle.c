/* CPU oracle for the SGAM hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the byte / integer / index work of the reference
 * (yshen47/SGAM_NeurIPS22 @ 780feff; paths relative to /root/reference).  Every
 * fp32 operation is written out in the order torch-CPU (MKL sgemm with K=3, ATen
 * elementwise kernels) evaluates it, so the outputs are bit-identical to the
 * reference's CPU path; tests/test_oracle_golden.py pins that against vectors
 * produced by the unmodified reference.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp (oracle/native.py).
 * -ffp-contract=off is essential: fused multiply-adds appear only where fmaf()
 * is written explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* MKL's K=3 sgemm micro-kernel: c = a0*x0 (rounded), then two fused multiply-adds
 * (probed against torch.bmm for every column count >= 64). */
static inline float dot3(const float *a, float x0, float x1, float x2) {
    float t = a[0] * x0;
    t = fmaf(a[1], x1, t);
    t = fmaf(a[2], x2, t);
    return t;
}

/* torch `.long()` on x86 = cvttss2si: truncation toward zero; NaN / out-of-range -> INT64_MIN. */
static inline int64_t f2l(float v) {
    if (!(v > -9.2e18f && v < 9.2e18f)) return INT64_MIN;
    return (int64_t)v;
}

/* warp.py:28-40 pixel2cam + :215 rigid transform.  (j = column = u, i = row = v; warp.py:12-18) */
static inline void unproject_transform(const float *Kinv, const float *T, int i, int j, float d, float *Y) {
    float fj = (float)j, fi = (float)i;
    float X[3];
    for (int r = 0; r < 3; ++r) X[r] = dot3(Kinv + 3 * r, fj, fi, 1.0f) * d;
    for (int r = 0; r < 3; ++r) Y[r] = dot3(T + 4 * r, X[0], X[1], X[2]) + T[4 * r + 3];
}

static float median9(float *v) {
    /* torch.median(dim): NaN if any NaN, else the lower median = 5th smallest of 9 (warp.py:341-345) */
    for (int a = 0; a < 9; ++a) if (v[a] != v[a]) return v[a];
    for (int a = 1; a < 9; ++a) {
        float x = v[a]; int b = a - 1;
        while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; --b; }
        v[b + 1] = x;
    }
    return v[4];
}

/* warp.py:289-347 median_blur(input,(3,3)): zero padded 3x3 median per plane. */
void oracle_median_blur3(const float *in, float *out, int planes, int H, int W) {
    for (int p = 0; p < planes; ++p) {
        const float *src = in + (size_t)p * H * W;
        float *dst = out + (size_t)p * H * W;
        for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) {
            float v[9]; int c = 0;
            for (int di = -1; di <= 1; ++di) for (int dj = -1; dj <= 1; ++dj) {
                int y = i + di, x = j + dj;
                v[c++] = (y >= 0 && y < H && x >= 0 && x < W) ? src[(size_t)y * W + x] : 0.0f;
            }
            dst[(size_t)i * W + j] = median9(v);
        }
    }
}

/* warp.py:193-286 render_projection_from_srcs_fast (dynamic_masks=None, depth_range=None),
 * sequential-scatter semantics (parallel=False; equal to parallel=True under
 * torch.use_deterministic_algorithms / one thread).
 *   src_rgb   [B,N,3,H,W]   src_depth [B,N,H,W]   K_tgt [B,3,3]   Kinv_src [B*N,3,3]   T [B*N,4,4]
 *   proj_rgb  [B,3,H,W]  proj_depth [B,H,W]   : scattered, unfilled (warp.py:251,261)
 *   merge_rgb [B,3,H,W]  merge_depth[B,H,W]   : after per-channel hole fill (warp.py:271-279)
 *   mask      [B,H,W] u8                      : merge_depth <= 0 (warp.py:285)
 *   winner    [B,H,W] i32                     : order index (pixel*N+n) of the last writer, -1 if none
 *   inbounds  [B,H*W*N] u8                    : warp.py:232 `mask`
 * zmin != 0 selects the nearest-depth policy instead (north_star's atomic-min z-buffer; a flagged
 * deviation from the reference): winner = smallest X_tgt.z, ties -> larger order index. */
void oracle_splat_forward(const float *src_rgb, const float *src_depth, const float *K_tgt,
                          const float *Kinv_src, const float *T, int B, int N, int H, int W, int zmin,
                          float *proj_rgb, float *proj_depth, float *merge_rgb, float *merge_depth,
                          uint8_t *mask, int32_t *winner, uint8_t *inbounds) {
    const size_t HW = (size_t)H * W;
    memset(proj_rgb, 0, sizeof(float) * B * 3 * HW);
    memset(proj_depth, 0, sizeof(float) * B * HW);
    for (size_t k = 0; k < (size_t)B * HW; ++k) winner[k] = -1;
    for (int b = 0; b < B; ++b) {
        const float *Kt = K_tgt + 9 * b;
        for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) for (int n = 0; n < N; ++n) {
            const int bn = b * N + n;
            const size_t p = (size_t)i * W + j;
            float d = src_depth[(size_t)bn * HW + p];
            float Y[3];
            unproject_transform(Kinv_src + 9 * bn, T + 16 * bn, i, j, d, Y);
            /* warp.py:222-225 */
            float pz = dot3(Kt + 6, Y[0], Y[1], Y[2]);
            float px = dot3(Kt + 0, Y[0], Y[1], Y[2]) / pz;
            float py = dot3(Kt + 3, Y[0], Y[1], Y[2]) / pz;
            int64_t x = f2l(px + 0.5f), y = f2l(py + 0.5f);
            int ok = (x >= 0) && (x < W) && (y >= 0) && (y < H);     /* warp.py:232 */
            if (inbounds) inbounds[(size_t)b * HW * N + p * N + n] = (uint8_t)ok;
            if (!ok) continue;
            size_t q = (size_t)y * W + x;
            int32_t order = (int32_t)(p * N + n);
            if (zmin && winner[b * HW + q] >= 0) {
                float cur = proj_depth[b * HW + q];
                if (!(Y[2] <= cur)) continue;   /* keep nearer; ties -> later order wins */
            }
            winner[b * HW + q] = order;
            for (int c = 0; c < 3; ++c)
                proj_rgb[((size_t)b * 3 + c) * HW + q] = src_rgb[((size_t)bn * 3 + c) * HW + p];
            proj_depth[b * HW + q] = Y[2];
        }
    }
    float *med = (float *)malloc(sizeof(float) * HW);
    for (int b = 0; b < B; ++b) for (int c = 0; c < 4; ++c) {
        const float *src = (c < 3) ? proj_rgb + ((size_t)b * 3 + c) * HW : proj_depth + (size_t)b * HW;
        float *dst = (c < 3) ? merge_rgb + ((size_t)b * 3 + c) * HW : merge_depth + (size_t)b * HW;
        oracle_median_blur3(src, med, 1, H, W);
        for (size_t k = 0; k < HW; ++k) {
            float m = (src[k] == 0.0f) ? 1.0f : 0.0f;          /* warp.py:271-279: bool*float arithmetic */
            dst[k] = m * med[k] + (1.0f - m) * src[k];
        }
    }
    free(med);
    for (size_t k = 0; k < (size_t)B * HW; ++k) mask[k] = (merge_depth[k] <= 0.0f) ? 1 : 0;
}

/* model.py:210-229 inverse-depth coding of the warped depth (dataset 0 = clevr-infinite, 1 = google_earth) */
void oracle_depth_code(const float *depth, const uint8_t *mask, float *out, size_t n, int dataset) {
    const float c16 = (float)(1.0 / 16), sC = (float)(1.0 / 7 - 1.0 / 16);
    const float cG = (float)(1.0 / 14.765625), sG = (float)(1.0 / 10.099975586 - 1.0 / 14.765625);
    for (size_t k = 0; k < n; ++k) {
        float w;
        if (dataset == 0) {
            float d = depth[k];
            d = (d < 1e-7f) ? 1e-7f : d;        /* torch.clip(x, 1e-7): NaN propagates */
            if (depth[k] != depth[k]) d = depth[k];
            w = 1.0f / d;
            w = (w - c16) / sC;
        } else {
            w = 1.0f / (depth[k] + 10.0f);
            w = (w - cG) / sG;
        }
        w = 2.0f * w - 1.0f;
        float m = mask[k] ? 1.0f : 0.0f;
        out[k] = w * (1.0f - m) + (1.0f * -2.0f) * m;
    }
}

/* inference_pipeline.py:906-911 metric depth from the decoded depth code */
void oracle_depth_decode(const float *code, float *out, size_t n, int dataset) {
    const float c16 = (float)(1.0 / 16), sC = (float)(1.0 / 7 - 1.0 / 16);
    const float cG = (float)(1.0 / 14.765625), sG = (float)(1.0 / 10.099975586 - 1.0 / 14.765625);
    for (size_t k = 0; k < n; ++k) {
        float v = (code[k] + 1.0f) / 2.0f;
        if (dataset == 0) out[k] = 1.0f / (v * sC + c16);
        else out[k] = 1.0f / (v * sG + cG) - 10.0f;
    }
}

/* inference_pipeline.py:898-901: clip((x+1)/2*255, 0, 255).astype(uint8) -- truncation. in: [3,H,W] -> out [H,W,3] */
void oracle_pack_u8(const float *rgb, uint8_t *out, int H, int W) {
    const size_t HW = (size_t)H * W;
    for (size_t p = 0; p < HW; ++p) for (int c = 0; c < 3; ++c) {
        float v = (rgb[c * HW + p] + 1.0f) / 2.0f * 255.0f;
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        out[p * 3 + c] = (uint8_t)v;
    }
}

/* inference_pipeline.py:619-743 inverse_warping (B = 1 per call of the reference; batched here).
 *   src_rgb [B,N,3,H,W]  src_depth [B,N,H,W]  tgt_depth [B,H,W]  Kinv_tgt [B,3,3]
 *   proj [B*N,3,4] = src_intrinsics @ T_tgt2srcs[:, :3]  (host-side 3x4 product, :696)
 *   out  [B,3,H,W]   best_src [B,H,W] i32 (-1 = none) */
void oracle_inverse_warp(const float *src_rgb, const float *src_depth, const float *tgt_depth,
                         const float *Kinv_tgt, const float *proj, int B, int N, int H, int W,
                         float *out, int32_t *best_src) {
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) {
        const size_t p = (size_t)i * W + j;
        float fj = (float)j, fi = (float)i, d = tgt_depth[b * HW + p];
        float X[3];
        for (int r = 0; r < 3; ++r) X[r] = dot3(Kinv_tgt + 9 * b + 3 * r, fj, fi, 1.0f) * d;   /* :680 */
        float zbuf = 99999.0f, res[3] = {0.0f, 0.0f, 0.0f};
        int32_t best = -1;
        for (int n = 0; n < N; ++n) {
            const int bn = b * N + n;
            const float *P = proj + 12 * bn;
            float pc[3];
            for (int r = 0; r < 3; ++r) pc[r] = dot3(P + 4 * r, X[0], X[1], X[2]) + P[4 * r + 3];  /* :640-646 */
            float Z = pc[2];
            float xn = 2.0f * (pc[0] / Z) / (float)(W - 1) - 1.0f;                                  /* :655-657 */
            float yn = 2.0f * (pc[1] / Z) / (float)(H - 1) - 1.0f;
            /* F.grid_sample(mode='nearest', align_corners=False, padding zeros) (:707); ATen CPU:
             * ix = (x+1)*(W/2) - 0.5, nearbyint (half to even), zero outside. */
            float ix = (xn + 1.0f) * ((float)W / 2.0f) - 0.5f;
            float iy = (yn + 1.0f) * ((float)H / 2.0f) - 0.5f;
            float rx = nearbyintf(ix), ry = nearbyintf(iy);
            float s[3] = {0.0f, 0.0f, 0.0f};
            if (rx >= 0.0f && rx < (float)W && ry >= 0.0f && ry < (float)H) {
                size_t q = (size_t)ry * W + (size_t)rx;
                for (int c = 0; c < 3; ++c) s[c] = src_rgb[((size_t)bn * 3 + c) * HW + q] + 2.0f;
            }
            float diff = fabsf(Z - src_depth[(size_t)bn * HW + p]);                              /* :698: unsampled */
            float sum = (s[0] + s[1]) + s[2];
            int m = (diff < zbuf) && (Z >= 0.0f) && (sum > 0.0f);                                 /* :731-733 */
            if (m) { zbuf = diff; best = n; for (int c = 0; c < 3; ++c) res[c] = s[c] - 2.0f; }
        }
        for (int c = 0; c < 3; ++c) out[((size_t)b * 3 + c) * HW + p] = res[c];
        if (best_src) best_src[b * HW + p] = best;
    }
}

/* quantize.py:285-289 / :348-352: d = (sum z^2 + sum e^2) - 2 z.e ; argmin with first-index ties.
 * Canonical fp32 order (shared bit-for-bit with the CUDA kernel): every dot product / squared norm
 * is one sequential fmaf chain over k = 0..D-1 starting from +0.
 *   z [T,D] (token-major), E [n_e,D]  ->  idx [T] i64, dmin [T], d2 [T] (second smallest, for gap reports) */
void oracle_vq_nearest(const float *z, const float *E, int T, int n_e, int D,
                       int64_t *idx, float *dmin, float *d2) {
    float *ee = (float *)malloc(sizeof(float) * n_e);
    #pragma omp parallel for
    for (int e = 0; e < n_e; ++e) {
        float s = 0.0f;
        for (int k = 0; k < D; ++k) s = fmaf(E[(size_t)e * D + k], E[(size_t)e * D + k], s);
        ee[e] = s;
    }
    #pragma omp parallel for
    for (int t = 0; t < T; ++t) {
        const float *zt = z + (size_t)t * D;
        float zz = 0.0f;
        for (int k = 0; k < D; ++k) zz = fmaf(zt[k], zt[k], zz);
        float best = INFINITY, second = INFINITY; int64_t bi = 0;
        for (int e = 0; e < n_e; ++e) {
            const float *ev = E + (size_t)e * D;
            float dot = 0.0f;
            for (int k = 0; k < D; ++k) dot = fmaf(zt[k], ev[k], dot);
            float d = (zz + ee[e]) - 2.0f * dot;
            if (d < best) { second = best; best = d; bi = e; }
            else if (d < second) second = d;
        }
        idx[t] = bi; dmin[t] = best; if (d2) d2[t] = second;
    }
    free(ee);
}

/* inference_pipeline.py:1014-1036 prepare_pcd: X_w = inv(Rt) [K^-1 [u v 1]^T d ; 1], float64.
 * numpy's dgemm evaluates the K = 3 / K = 4 dot products as one rounded product followed by fused multiply-adds (same
 * pattern as the fp32 dot3 above; pinned bit-exact by tests/golden/prepare_pcd_vectors.npz). */
void oracle_unproject_world(const float *depth, const double *Kinv, const double *Rt_inv, int H, int W, double *xyz) {
    for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) {
        size_t p = (size_t)i * W + j;
        double d = (double)depth[p], c[3];
        for (int r = 0; r < 3; ++r) {
            double t = Kinv[3 * r] * (double)j;
            t = fma(Kinv[3 * r + 1], (double)i, t);
            t = fma(Kinv[3 * r + 2], 1.0, t);
            c[r] = d * t;
        }
        for (int r = 0; r < 3; ++r) {
            double t = Rt_inv[4 * r] * c[0];
            t = fma(Rt_inv[4 * r + 1], c[1], t);
            t = fma(Rt_inv[4 * r + 2], c[2], t);
            t = fma(Rt_inv[4 * r + 3], 1.0, t);
            xyz[p * 3 + r] = t;
        }
    }
}
