// Final map: per-frame unprojection to world-space points.  Replaces InfiniteSceneGeneration.prepare_pcd
// (sgam/inference_pipeline.py:1014-1036) as called by unproject_to_color_point_cloud (:1038-1062):
//   X_w = inv(Rt) [ K^-1 [u v 1]^T d ; 1 ]   (float64),   colour = uint8 / 255.
// One thread per pixel, frames in grid.y; HBM-bound (4 B depth + 3 B colour in, 48 B out per point).  The dot products
// follow numpy's dgemm order -- one rounded product, then fused multiply-adds (oracle/csrc/oracle.c, pinned to the
// reference by tests/golden/prepare_pcd_vectors.npz) -- so the points are bit-identical to the reference's.
#include "common.cuh"

namespace {

struct Mat3d { double m[9]; };

__global__ void __launch_bounds__(256)
unproject_points_kernel(const float *__restrict__ depth, const uint8_t *__restrict__ rgb, Mat3d Kinv,
                        const double *__restrict__ Rt_inv, int H, int W, double *__restrict__ xyz, double *__restrict__ col) {
    const int f = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int i = p / W, j = p - i * W;
    const size_t gp = (size_t)f * H * W + p;
    const double d = (double)depth[gp];
    const double *R = Rt_inv + (size_t)f * 12;
    double c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double t = __dmul_rn(Kinv.m[3 * r], (double)j);
        t = __fma_rn(Kinv.m[3 * r + 1], (double)i, t);
        t = __fma_rn(Kinv.m[3 * r + 2], 1.0, t);
        c[r] = __dmul_rn(d, t);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double t = __dmul_rn(R[4 * r], c[0]);
        t = __fma_rn(R[4 * r + 1], c[1], t);
        t = __fma_rn(R[4 * r + 2], c[2], t);
        t = __fma_rn(R[4 * r + 3], 1.0, t);
        xyz[gp * 3 + r] = t;
    }
    if (col) {
#pragma unroll
        for (int k = 0; k < 3; ++k) col[gp * 3 + k] = (double)rgb[gp * 3 + k] / 255.0;
    }
}

}  // namespace

extern "C" int sgam_unproject_points(const float *depth, const uint8_t *rgb_u8, const double *host_Kinv, const double *Rt_inv,
                                     int F, int H, int W, double *xyz, double *colors, void *stream) {
    SGAM_REQUIRE(depth && host_Kinv && Rt_inv && xyz, "unproject_points: null pointer");
    SGAM_REQUIRE((rgb_u8 == nullptr) == (colors == nullptr), "unproject_points: rgb_u8 and colors go together");
    SGAM_REQUIRE(F > 0 && H > 0 && W > 0 && F <= 65535, "unproject_points: bad F/H/W");
    Mat3d K;
    for (int i = 0; i < 9; ++i) K.m[i] = host_Kinv[i];
    unproject_points_kernel<<<dim3(cdiv((long long)H * W, 256), F), 256, 0, (cudaStream_t)stream>>>(depth, rgb_u8, K, Rt_inv, H, W, xyz, colors);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}
