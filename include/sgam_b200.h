/* sgam_b200.h -- C ABI of libsgam_b200.so: the B200 (sm_100a) kernels behind the SGAM per-frame hot path.
 *
 * The reference (yshen47/SGAM_NeurIPS22 @ 780feff) is pure PyTorch and has no FFI of its own; each
 * entry point below replaces the ATen call sequence of one reference function (cited as
 * file:line relative to the reference root) and is bound from Python with ctypes
 * (sgam_neurips22_b200/_lib.py; the reference-side stub is shown in INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory (PyTorch tensors) unless named host_*;
 *     the library never allocates, frees or keeps a pointer past the call;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no synchronisation), so they
 *     can be captured into a CUDA graph;
 *   - return 0 on success, a negative sgam_status on failure; sgam_last_error() gives the message of the
 *     last failure on the calling thread; nothing throws;
 *   - fp32 everywhere unless stated; "NHWC" = [B,H,W,C] contiguous, "NCHW" = [B,C,H,W] contiguous;
 *   - re-entrant per stream: one process per GPU (or several threads with distinct streams and buffers).
 */
#ifndef SGAM_B200_H
#define SGAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SGAM_OK = 0,
    SGAM_ERR_INVALID = -1,     /* bad shape / null pointer / unsupported flag combination */
    SGAM_ERR_CUDA = -2,        /* a CUDA runtime call or kernel launch failed */
    SGAM_ERR_UNSUPPORTED = -3  /* valid request the library has no kernel for */
} sgam_status;

enum { SGAM_DATASET_CLEVR = 0, SGAM_DATASET_GOOGLE_EARTH = 1 };
enum { SGAM_SPLAT_LAST_WRITER = 0,   /* reference order: last point in (pixel-major, source-minor) order wins */
       SGAM_SPLAT_ZMIN = 1 };        /* nearest depth wins (atomic-min z-buffer); flagged deviation */

const char *sgam_last_error(void);
int sgam_version(void);              /* 10000*major + 100*minor + patch */
int sgam_sm_count(int device);       /* multiprocessors of `device` (148 on B200), <0 on error */
unsigned long long sgam_launch_count(void);   /* kernels this library has launched (or captured) so far in the process */

/* ------------------------------------------------------------------------------------------------
 * Stage (i): forward splat.  Replaces sgam/point_rendering/warp.py:193-286
 * render_projection_from_srcs_fast (pixel2cam :28-40, rigid transform :215, projection :222-225,
 * index_put_ scatter :251,261, per-channel 3x3 median hole fill :271-279 via median_blur :289-347,
 * hole mask :285) fused with VQModel.get_x's inverse-depth coding (sgam/generative_sensing_module/
 * model.py:210-229,237).
 *
 *   src_rgb    [B,N,*]  source colours; element (b,n,c,pixel p) at ((b*N+n)*3*H*W + c*rgb_cs + p*rgb_ps):
 *              rgb_cs=H*W, rgb_ps=1 for [B,N,3,H,W]; rgb_cs=1, rgb_ps=3 for the batch's native [B,N,H,W,3]
 *   src_depth  [B,N,H,W]
 *   K_tgt      [B,3,3]   target intrinsics            Kinv_src [B*N,3,3] inverse source intrinsics
 *   T_src2tgt  [B*N,4,4] row-major rigid transforms (model.py:188-195)
 *   winner     [B,H,W] u64 workspace (sgam_splat_workspace_bytes); on return holds the per-pixel winner
 *              key: 0 = no point; low 32 bits = 1 + (p*N + n) of the winning source point
 *   x          [B,4,H,W] out: filled rgb (3) + inverse-depth code (holes = -2)      (model.py:237)
 *   mask       [B,H,W] u8 out: extrapolation mask (merged depth <= 0)                 (warp.py:285)
 *   merge_depth[B,H,W] out or NULL: filled metric depth                              (warp.py:279)
 *   proj       [B,4,H,W] out or NULL: scattered, unfilled rgb + depth                 (warp.py:251,261)
 *   inbounds   [B,H*W*N] u8 out or NULL: warp.py:232 bounds mask in (pixel, source) order
 */
size_t sgam_splat_workspace_bytes(int B, int H, int W);
int sgam_splat_forward(const float *src_rgb, long long rgb_cs, long long rgb_ps, const float *src_depth,
                       const float *K_tgt, const float *Kinv_src, const float *T_src2tgt,
                       int B, int N, int H, int W, int policy, int dataset,
                       void *winner, float *x, uint8_t *mask, float *merge_depth, float *proj,
                       uint8_t *inbounds, void *stream);

/* 3x3 zero-padded lower median per plane.  Replaces warp.py:289-347 median_blur(input,(3,3)). */
int sgam_median_blur3(const float *in, float *out, int planes, int H, int W, void *stream);

/* get_x for a pre-warped input (use_rgbd_integration=True): model.py:196-199,210-229,237.
 *   rgb [B,3,H,W], depth [B,H,W] -> x [B,4,H,W], mask [B,H,W] u8 (depth <= 0) */
int sgam_depth_code(const float *rgb, const float *depth, int B, int H, int W, int dataset,
                    float *x, uint8_t *mask, void *stream);

/* Backward warp + per-pixel source selection.  Replaces sgam/inference_pipeline.py:662-743
 * InfiniteSceneGeneration.inverse_warping (pixel2cam :619-632, cam2pixel :634-660, nearest
 * grid_sample :707, depth-residual z-test loop :725-737).
 *   src_rgb as in sgam_splat_forward; src_depth [B,N,H,W]; tgt_depth [B,H,W]; Kinv_tgt [B,3,3];
 *   proj [B*N,3,4] = K_src @ T_tgt2src[:3] (host-side 3x4 product, :696)
 *   out [B,3,H,W] (0 where no source is valid); best_src [B,H,W] i32 or NULL (-1 = none) */
int sgam_inverse_warp(const float *src_rgb, long long rgb_cs, long long rgb_ps, const float *src_depth,
                      const float *tgt_depth, const float *Kinv_tgt, const float *proj,
                      int B, int N, int H, int W, float *out, int32_t *best_src, void *stream);

/* Per-frame output conversion.  Replaces inference_pipeline.py:893-911 and, for the device-resident frame
 * store, the PNG re-load of :534.
 *   dec [B,4,H,W] -> rgb_u8 [B,H,W,3] (clip((x+1)/2*255) truncated), depth [B,H,W] metric,
 *   src_rgb [B,H,W,3] or NULL: (float)(u8 / 127.5 - 1.0) evaluated in double, i.e. the fp32 source image a later
 *   step would obtain by re-reading the saved PNG */
int sgam_frame_outputs(const float *dec, int B, int H, int W, int dataset, uint8_t *rgb_u8, float *depth,
                       float *src_rgb, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage (ii): codebook nearest neighbour.  Replaces sgam/generative_sensing_module/modules/vqvae/
 * quantize.py:275-319 VectorQuantizer2.forward and :344-381 get_multiple_codewords(topk=1):
 * d = (|z|^2 + |e|^2) - 2 z.e, arg-min with first-index ties, z_q = E[idx].
 *   z [T,D] token-major (= NHWC latent), codebook [n_e,D], best [T] u64 workspace,
 *   idx [T] i64 out, z_q [T,D] out, dmin [T] out or NULL.  D % 32 == 0, n_e % 128 == 0.
 */
size_t sgam_vq_workspace_bytes(int T);
int sgam_vq_nearest(const float *z, const float *codebook, int T, int n_e, int D, void *best,
                    int64_t *idx, float *z_q, float *dmin, void *stream);

/* quantize.py:344-381 get_multiple_codewords for topk > 1: the topk nearest codes per token
 * (ascending canonical distance), p = softmax(-d_topk), `samples` multinomial draws with replacement from a
 * counter-based generator keyed by (seed, token, sample), tokens whose nearest-down-sampled extrapolation mask is 0
 * pinned to the nearest code.  row0_probs = 1 reproduces the reference, which draws every token from token 0's
 * probabilities (quantize.py:358).  Bit parity with torch's global generator is impossible; parity is distributional.
 *   mask [images, H, W] u8 or NULL (all extrapolated); token (ly,lx) of image i reads mask[i*mask_bstride + (ly*fy)*mask_w + lx*fx]
 *   outputs: topk_idx [T,topk] i64, topk_p [T,topk], idx [T,samples] i64, z_q [T,samples,D] */
size_t sgam_vq_topk_workspace_bytes(int T, int n_e);
int sgam_vq_topk_sample(const float *z, const float *codebook, const uint8_t *mask, long long mask_bstride, int mask_w,
                        int fy, int fx, int lat_w, int tokens_per_image, int T, int n_e, int D, int topk, int samples,
                        unsigned long long seed, int row0_probs, void *workspace, int64_t *topk_idx, float *topk_p,
                        int64_t *idx, float *z_q, void *stream);

/* Tensor-core variant of the same search (bit-identical idx / z_q / dmin): approximate distances on tcgen05 (split
 * bf16, fp32 accumulate) reduced to one minimum per (token, 128-code tile), then the tiles within a proven error
 * slack of the best are re-evaluated in the canonical fp32 order.  z_hi/z_lo, e_hi/e_lo: split-bf16 planes of z and
 * the codebook (sgam_split_bf16); ee [n_e] = canonical squared code norms (sgam_vq_norms, cacheable per codebook),
 * ee_max = max(ee); workspace: sgam_vq_tc_workspace_bytes(T, n_e).  D % 64 == 0, n_e % 128 == 0. */
int sgam_vq_norms(const float *x, float *out, int rows, int D, void *stream);
size_t sgam_vq_tc_workspace_bytes(int T, int n_e);
int sgam_vq_nearest_tc(const float *z, const void *z_hi, const void *z_lo, const float *codebook, const void *e_hi,
                       const void *e_lo, const float *ee, float ee_max, int T, int n_e, int D, void *workspace,
                       int64_t *idx, float *z_q, float *dmin, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage (iii): VQGAN encoder / decoder operators (sgam/generative_sensing_module/modules/
 * diffusionmodules/model.py).  Activations are NHWC inside the network.
 */

/* VQModel.encode stem: cat(x, mask) -> 1x1 conv 5->4 (model.py:106-113, conv_in :54).
 *   x [B,4,H,W] NCHW, mask [B,H,W] u8 or NULL (zeros), w [4,5], bias [4] -> y [B,H,W,4] NHWC */
int sgam_stem_conv(const float *x, const uint8_t *mask, const float *w, const float *bias,
                   int B, int H, int W, float *y, void *stream);

/* Conv2d 3x3 / 1x1 as implicit GEMM (model.py:43,62-72,88-117; Downsample :56-75; Upsample :38-53).
 *   x [B,H,W,Cin]; w [Cout, k*k*Cin] with K index (kh*k+kw)*Cin+ci (OIHW repacked once by the host);
 *   bias [Cout]; residual [B,Ho,Wo,Cout] or NULL (added in the epilogue); y [B,Ho,Wo,Cout]
 *   ksize 1|3; stride 1|2; pad_mode 0 = symmetric k/2, 1 = right/bottom only (Downsample: pad (0,1,0,1),
 *   stride 2); upsample 1 = the input is read through a nearest x2 up-sampling (Upsample fused).
 *   out_nchw 1 = write y as [B,Cout,Ho,Wo] (decoder conv_out). */
int sgam_conv2d(const float *x, const float *w, const float *bias, const float *residual, float *y,
                int B, int H, int W, int Cin, int Cout, int ksize, int stride, int pad_mode,
                int upsample, int out_nchw, void *stream);

/* GroupNorm(32, C, eps=1e-6) (+ swish) (model.py:29-35).  Two launches: statistics, then apply.
 *   x [B,HW,C]; partial [B,S,32,2] f64 workspace (S = sgam_gn_splits(HW)); y [B,HW,C] */
int sgam_gn_splits(long long HW);
int sgam_groupnorm(const float *x, const float *gamma, const float *beta, float *y, double *partial,
                   int B, long long HW, int C, int swish, void *stream);

/* Batched C = alpha * A . B^T (+ bias_m[row]) : A [batch,M,K], B [batch,N,K], C [batch,M,N]; strides in
 * elements; used for the attention score / value products and V^T = W_v . h^T (model.py:168-190). */
int sgam_gemm_nt(const float *A, const float *Bm, float *C, const float *bias_m, int batch, int M, int N,
                 int K, long long sA, long long sB, long long sC, float alpha, void *stream);

/* Row softmax in place: x [rows, cols] (model.py:181). */
int sgam_softmax_rows(float *x, long long rows, int cols, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stage (iii), tensor-core path (tcgen05 + TMA + TMEM).  Operands are split-bf16 planes: x = hi + lo with
 * hi = bf16(x), lo = bf16(x - hi); products are accumulated as hi*hi + hi*lo + lo*hi in fp32 (nsplit = 3) or
 * hi*hi only (nsplit = 1).  Same operators as above (model.py line references as for sgam_conv2d /
 * sgam_groupnorm / sgam_gemm_nt / sgam_softmax_rows).
 */

/* Stem for the tensor-core path: as sgam_stem_conv, but the [B,H,W,4] result is written as split bf16 with the channel
 * dimension zero-padded to Cpad (64), so that encoder.conv_in (Cin = 4) runs as a K = 9*Cpad tensor-core GEMM. */
int sgam_stem_conv_split(const float *x, const uint8_t *mask, const float *w, const float *bias, int B, int H, int W,
                         int Cpad, void *hi, void *lo, void *stream);

/* fp32 NHWC [B,H,W,C] -> hi / lo bf16 [B,H<<up,W<<up,C] (nearest x2 up-sampling fused: Upsample, model.py:50). */
int sgam_split_bf16(const float *x, void *hi, void *lo, int B, int H, int W, int C, int upsample, void *stream);

/* GroupNorm(32, C, eps=1e-6) (+ swish) with split-bf16 output; partial as for sgam_groupnorm. */
int sgam_groupnorm_split(const float *x, const float *gamma, const float *beta, void *hi, void *lo, double *partial,
                         int B, long long HW, int C, int swish, void *stream);

/* Row softmax of fp32 scores x [rows, cols] -> split-bf16 probabilities. */
int sgam_softmax_split(const float *x, void *hi, void *lo, long long rows, int cols, void *stream);

/* 1 if sgam_conv2d_tc has a kernel for a conv whose OUTPUT grid is H x W. */
int sgam_tc_supported_conv(int H, int W, int Cin, int Cout, int ksize, int stride);

/* Conv 3x3 / 1x1 on tensor cores: stride 1 with symmetric padding, or stride 2 = the Downsample (zero pad right /
 * bottom, 3x3, model.py:68-72; strided TMA box).  x_hi/x_lo [B,H,W,Cin] bf16; w_hi/w_lo [ceil32(Cout), k*k*Cin]
 * bf16 K-major (rows beyond Cout zero); bias [Cout] or NULL; residual [B,Ho,Wo,Cout] fp32 or NULL.
 * Output fp32 y (NHWC, or [B,Cout,Ho,Wo] when out_nchw: the decoder head) and/or split-bf16 y_hi/y_lo. */
int sgam_conv2d_tc(const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo, const float *bias,
                   const float *residual, float *y, void *y_hi, void *y_lo, int B, int H, int W, int Cin, int Cout,
                   int ksize, int stride, int out_nchw, int nsplit, float *gn_partial, float *splitk_ws,
                   double *splitk_gn_partial, int *host_stats_written, void *stream);
/* host_stats_written (HOST pointer or NULL) receives which GroupNorm statistics the call produced: 0 none, 1 the fp32
 * per-pixel-block sums in gn_partial, 2 the fp64 split sums in splitk_gn_partial -- the library, not the caller, knows
 * which kernel it chose. */

/* Upsample = nearest x2 + 3x3 conv (diffusionmodules/model.py:49-52) in sub-pixel form: each output parity (py, px) is a
 * 2x2 convolution of the LOW-resolution operand with pre-summed taps.  x_hi/x_lo [B,H,W,Cin] (low resolution);
 * w_hi/w_lo [4][Cout][4*Cin] (parity 2*py+px; K index (dy*2+dx)*Cin+ci; dy/dx = 0 is the row/column i-1+py / j-1+px);
 * y [B,2H,2W,Cout] fp32; gn_partial as for sgam_conv2d_tc, sized for the OUTPUT grid.  16/36 of the MMA work of the 3x3
 * form and no up-sampled operand in HBM. */
int sgam_conv2d_tc_up2_supported(int B, int H, int W, int Cin, int Cout);
int sgam_conv2d_tc_up2(const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo, const float *bias, float *y,
                       int B, int H, int W, int Cin, int Cout, int nsplit, float *gn_partial, void *stream);

/* Split-K for under-filled grids (low-resolution layers, single-trajectory batches): when splitk_ws is given and
 * sgam_conv2d_tc_splitk_floats(...) > 0 floats, the K loop is divided among CTAs, partial tiles go to the workspace
 * and a reduce kernel adds them in split order with bias / residual (deterministic; fp32 NHWC output only, no fused
 * GroupNorm statistics). */
long long sgam_conv2d_tc_splitk_floats(int B, int H, int W, int Cin, int Cout, int ksize, int stride);
/* When the K loop is split and splitk_gn_partial ([B][sgam_gn_splits(Ho*Wo)][32][2] f64) is given (Cout % 128 == 0), the
 * reduce kernel also accumulates the GroupNorm partial sums of the finished tensor -- the layout sgam_groupnorm_split
 * computes with its own statistics pass -- and sgam_groupnorm_split_apply consumes them without re-reading statistics. */
int sgam_groupnorm_split_apply(const float *x, const float *gamma, const float *beta, void *hi, void *lo,
                               const double *partial, int B, long long HW, int C, int swish, void *stream);

/* GroupNorm statistics fused into the conv epilogue: when gn_partial (sgam_tc_gn_partial_floats(B,Ho,Wo) floats) is
 * passed to sgam_conv2d_tc, each pixel block writes the per-group sum / sum of squares of the finished output, and
 * sgam_groupnorm_split_fused finalises them (fp64) and applies GroupNorm(32,C,eps=1e-6) (+ swish) -> split bf16
 * without re-reading the tensor for statistics. */
long long sgam_tc_gn_partial_floats(int B, int Ho, int Wo);
int sgam_groupnorm_split_fused(const float *x, const float *gamma, const float *beta, void *hi, void *lo,
                               float *gn_partial, int B, int Ho, int Wo, int C, int swish, void *stream);

/* Encoder entry: the stem (cat(x, mask) -> 1x1 conv 5 -> 4, model.py:106-113) + encoder.conv_in (3x3, 4 -> 128,
 * diffusionmodules/model.py:370) + the GroupNorm partial sums of the result in one exact-fp32 kernel (36 real taps: FP32-pipe
 * work, not a K = 576 zero-padded GEMM).  x [B,4,H,W] NCHW, mask [B,H,W] u8 or NULL, w1 [4,5], b1 [4], w3 [128, 36]
 * K-major (kh, kw, ci), b3 [128]; y fp32 NHWC [B,H,W,128]; gn_partial: sgam_tc_gn_partial_floats(B,H,W) floats or NULL. */
int sgam_stem_conv_in(const float *x, const uint8_t *mask, const float *w1, const float *b1, const float *w3, const float *b3,
                      float *y, float *gn_partial, int B, int H, int W, int Cout, void *stream);

/* Decoder head: decoder.norm_out (GroupNorm(32,128), eps 1e-6) + swish + decoder.conv_out (3x3, 128 -> 4, NCHW output;
 * diffusionmodules/model.py:534-538) in one exact-fp32 kernel: the normalised halo tile is staged in shared memory once,
 * the channel reduction runs on the FP32 pipes (a 4-column GEMM tile would waste 97 % of the tensor pipe and re-read the
 * operand nine times).  x fp32 NHWC [B,H,W,128]; gn_partial = the partial sums sgam_conv2d_tc emitted for x
 * (sgam_tc_gn_partial_floats(B,H,W) floats); w_t [9][128][4] (tap, input channel, output channel); y [B,4,H,W]. */
int sgam_gn_head_conv(const float *x, const float *gamma, const float *beta, float *gn_partial, const float *w_t,
                      const float *bias, float *y, int B, int H, int W, int C, int Cout, void *stream);

/* Fused single-head self-attention of a 256-channel AttnBlock (diffusionmodules/model.py:168-192: bmm(q,k) * C^-0.5,
 * softmax over keys, bmm(v, w^T)) as ONE flash-style kernel: scores and probabilities live in tensor memory, the
 * [B,T,T] matrix is never written.  q, k: [B,T,C] split bf16; vt = V^T [B,C,T] split bf16 (so that P.V is an A.B^T
 * product with K-major operands); o: [B,T,C] split bf16 (the proj_out conv's operand).  Needs C == 256, T % 256 == 0. */
int sgam_attention_tc_supported(int B, int T, int C);
/* Small batches do not have enough 256-query tiles to fill the SM pairs: the keys of a tile are then split kv_splits ways
 * (flash-decoding style), every (tile, split) item leaves an un-normalised O and its (reference maximum, row sum) in the
 * workspace and a merge kernel combines them.  sgam_attention_tc_splits gives the split count the library would choose
 * (1 = no split, no workspace); the workspace holds sgam_attention_tc_workspace_bytes(B, T, kv_splits) bytes. */
int sgam_attention_tc_splits(int B, int T);
size_t sgam_attention_tc_workspace_bytes(int B, int T, int kv_splits);
/* ld_qk: elements between consecutive token rows of q and of k (0 = C, dense); 2C when q and k are the two column halves of
 * sgam_qkv_tc's [B,T,2C] output. */
int sgam_attention_tc(const void *q_hi, const void *q_lo, const void *k_hi, const void *k_lo, const void *vt_hi,
                      const void *vt_lo, void *o_hi, void *o_lo, int B, int T, int C, float scale, int kv_splits,
                      void *workspace, int ld_qk, void *stream);

/* GroupNorm(32, Cin, eps = 1e-6) + swish + 3x3 conv (the ResnetBlock pattern norm -> nonlinearity -> conv, diffusionmodules/
 * model.py:117-131) as ONE tensor-core kernel for the 128-channel layers on wide images: the fp32 activation x [B,H,W,Cin] is
 * normalised, activated and split into bf16 planes inside the conv's operand path (no separate apply pass, no split operand in
 * HBM).  gn_partial_in: the per-pixel-block sums x's producer left (sgam_tc_gn_partial_floats(B,H,W) floats; the library
 * finalises them into its tail); w [Cout, 9 Cin] split bf16; residual / y / (y_hi, y_lo) / gn_partial_out as sgam_conv2d_tc.
 * Operand bits and results are identical to sgam_groupnorm_split_fused followed by sgam_conv2d_tc. */
int sgam_gn_conv2d_tc_supported(int B, int H, int W, int Cin, int Cout);
int sgam_gn_conv2d_tc(const float *x, float *gn_partial_in, const float *gamma, const float *beta, const void *w_hi,
                      const void *w_lo, const float *bias, const float *residual, float *y, void *y_hi, void *y_lo, int B, int H,
                      int W, int Cin, int Cout, int nsplit, float *gn_partial_out, void *stream);

/* The q, k and v projections of an AttnBlock (three 1x1 convs of the same GroupNorm output, diffusionmodules/model.py:
 * 158-175) as ONE tcgen05 implicit GEMM with 3C output columns.  x [B,H,W,C] split bf16; w [3C, C] split bf16 = the rows of
 * q.weight, k.weight, v.weight; bias [3C] fp32.  qk [B, H*W, 2C] split bf16 (q = columns [0,C), k = [C,2C)); vt [B, C, H*W]
 * split bf16 = V^T, stored transposed by the epilogue (the K-major operand sgam_attention_tc / sgam_gemm_nt_tc expect).
 * Needs C % 128 == 0, H*W % 8 == 0 and a shape sgam_tc_supported_conv accepts. */
int sgam_qkv_tc_supported(int B, int H, int W, int C);
int sgam_qkv_tc(const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo, const float *bias, void *qk_hi,
                void *qk_lo, void *vt_hi, void *vt_lo, int B, int H, int W, int C, int nsplit, void *stream);

/* Batched C = alpha * A . B^T (+ bias_m[row]) on tensor cores.  A [batch|1, M, K], B [batch|1, N, K] split-bf16
 * (a_batched / b_batched say whether the operand has the batch dimension); output fp32 C and/or split-bf16.
 * lda / ldb: elements between consecutive rows of A / B (0 = K, dense) -- e.g. 2C for the q and k halves of sgam_qkv_tc's
 * output; the batch slices are then M * lda (N * ldb) elements apart. */
int sgam_gemm_nt_tc(const void *a_hi, const void *a_lo, const void *b_hi, const void *b_lo, const float *bias_m,
                    float *C, void *c_hi, void *c_lo, int batch, int M, int N, int K, int a_batched, int b_batched,
                    float alpha, int nsplit, int lda, int ldb, void *stream);

/* ------------------------------------------------------------------------------------------------
 * RGB-D integration (use_rgbd_integration=True).  Replaces InfiniteSceneGeneration.rgbd_integration
 * (sgam/inference_pipeline.py:745-838) and volume.extract_point_cloud() (:446-447), i.e. the reference's calls into
 * open3d==0.15.2 (ScalableTSDFVolume :119-131, integrate :777, extract_triangle_mesh :786,
 * OffscreenRenderer.render_to_depth_image :825).  The unit hash becomes a PAGE TABLE over the grid of 16^3-voxel units
 * of the box [o*, o*+n*) (unit indices; unit i covers world [i*16*voxel_length, (i+1)*16*voxel_length)) in front of a
 * pool of unit blocks; the target depth is ray-cast from the volume instead of meshed and rasterised (parity
 * with Open3D is unpinned).
 *   stamp [nx*ny*nz] u32: 0 = never opened, else the `frame` counter of the unit's last integration (zero it once)
 *   page  [nx*ny*nz] i32: 0 = no block, else 1 + the unit's block in the pool (zero it once)
 *   pool_state [3] i32: blocks handed out (may run past the capacity), capacity in blocks (set by the caller), units
 *                 dropped because the pool was full (zero [0] and [2] once)
 *   vol   [capacity][4096][2] fp32 (tsdf, weight), zero-initialised; color [capacity][4096][3] fp32 (0..255) or NULL
 *   unit (ux,uy,uz) -> (uz*ny + uy)*nx + ux ; voxel (lx,ly,lz) -> (lx*16 + ly)*16 + lz
 *   host_* pointers are HOST memory: row-major 3x4 poses and K = {fx, fy, cx, cy}; they are copied into the launch.
 */
size_t sgam_tsdf_volume_bytes(int nx, int ny, int nz, int with_color);      /* bytes of a pool that holds EVERY unit of the box */
size_t sgam_tsdf_block_bytes(int with_color);                               /* bytes of one pool block */
/* one RGB-D frame: depth [H,W] (>= depth_trunc or <= 0 ignored), rgb [H,W,3] fp32 in [-1,1] (uint8 lattice) or NULL;
 * opens the units within sdf_trunc of every `stride`-th depth sample (host_cam2world, fp64) and integrates them
 * (host_world2cam, fp32).  frame must be non-zero and distinct per call; work [1 + nx*ny*nz] i32 scratch (the
 * frame's list of opened units). */
int sgam_tsdf_integrate(const float *depth, const float *rgb, int H, int W, const double *host_cam2world,
                        const float *host_world2cam, const double *host_K, int stride, float depth_trunc,
                        int ox, int oy, int oz, int nx, int ny, int nz, float voxel_length, float sdf_trunc,
                        uint32_t *stamp, uint32_t frame, int *work, int *page, int *pool_state, float *vol, float *color,
                        void *stream);
/* view-space z of the first + -> - crossing along the ray of pixel (u + pixel_center, v + pixel_center); 0 = no surface.
 * Samples every step_vox voxel lengths of z in [z_near, z_far], skipping never-opened units.  out [H,W]. */
int sgam_tsdf_raycast(const int *page, const float *vol, int ox, int oy, int oz, int nx, int ny, int nz,
                      float voxel_length, float sdf_trunc, const float *host_cam2world, const double *host_K,
                      float pixel_center, int H, int W, float z_near, float z_far, float step_vox, float *out,
                      void *stream);
/* zero-crossing point cloud.  Pass 1 (unit_offsets NULL): unit_counts[unit] = crossings in the unit.  Pass 2:
 * unit_offsets = exclusive prefix sum of the counts; xyz [n,3], rgb [n,3] in [0,1] are written in (unit, lx, ly, lz,
 * axis) order -- deterministic. */
int sgam_tsdf_extract(const int *page, const float *vol, const float *color, int ox, int oy, int oz,
                      int nx, int ny, int nz, float voxel_length, float sdf_trunc, long long *unit_counts,
                      const long long *unit_offsets, float *xyz, float *rgb, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Final map.  Replaces InfiniteSceneGeneration.prepare_pcd (sgam/inference_pipeline.py:1014-1036) for a stack of
 * frames: xyz[f,p] = inv(Rt_f) [K^-1 [u v 1]^T depth ; 1] in float64 (numpy dgemm rounding order), colors = u8 / 255.
 *   depth [F,H,W] fp32; rgb_u8 [F,H,W,3] or NULL; host_Kinv: 9 doubles (HOST); Rt_inv [F,12] doubles (device; rows 0-2
 *   of the inverse world->camera matrix); xyz [F*H*W,3] f64; colors [F*H*W,3] f64 or NULL. */
int sgam_unproject_points(const float *depth, const uint8_t *rgb_u8, const double *host_Kinv, const double *Rt_inv,
                          int F, int H, int W, double *xyz, double *colors, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SGAM_B200_H */
