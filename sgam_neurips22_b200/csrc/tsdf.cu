// RGB-D integration on the device: dense-grid TSDF fusion + ray-cast target depth + zero-crossing point cloud.
//
// Replaces InfiniteSceneGeneration.rgbd_integration (sgam/inference_pipeline.py:745-838) and the
// volume.extract_point_cloud() of scene_expansion (:446-447), i.e. the reference's calls into open3d==0.15.2
// (ScalableTSDFVolume.integrate / extract_triangle_mesh / OffscreenRenderer.render_to_depth_image).  The arithmetic
// follows the published algorithm of that version (see oracle/csrc/tsdf_oracle.c, to which these kernels are
// bit-exact); the B200 re-design is
//   * a PAGE TABLE over the dense grid of 16^3-voxel units of the reachable box (4 bytes per unit) in front of a POOL of
//     unit blocks in HBM, instead of a host-side hash of units: a unit is "opened" by stamping it (the first stamper of a
//     frame queues it on the frame's work list and, the first time ever, takes the next pool block with one integer
//     atomic), every unit is one contiguous 32 KB (+ 48 KB colour) block swept by one CTA with fully coalesced 8-byte
//     accesses, no host round trip, and results that do not depend on which block a unit got (round 1 kept the whole
//     box dense: 7-15 GB per GoogleEarth trajectory; the pool holds only the units near observed surfaces);
//   * the target depth is ray-cast straight from the volume (one thread per pixel, trilinear samples, empty units
//     skipped) instead of marching cubes + mesh upload + rasterisation every step.
// All three kernels are HBM / L2 latency bound integer-and-fp32 work; compiled with --fmad=false so that every
// operation rounds exactly like the C oracle.
#include "common.cuh"

namespace {

constexpr int RES = 16, UNIT_VOX = RES * RES * RES;

struct Grid {
    int ox, oy, oz, nx, ny, nz;
    float voxel_length, sdf_trunc;
};
struct Pose34f { float m[12]; };
struct Pose34d { double m[12]; };
struct Intr { double fx, fy, cx, cy; };

__device__ __forceinline__ long long unit_index(const Grid &g, int ux, int uy, int uz) {
    ux -= g.ox; uy -= g.oy; uz -= g.oz;
    if (ux < 0 || uy < 0 || uz < 0 || ux >= g.nx || uy >= g.ny || uz >= g.nz) return -1;
    return ((long long)uz * g.ny + uy) * g.nx + ux;
}

// ScalableTSDFVolume::Integrate, first half: one thread per strided depth sample opens the units around its point.
// pool_state: [0] blocks handed out so far (may run past the capacity), [1] capacity, [2] units dropped for lack of blocks
__global__ void __launch_bounds__(256)
tsdf_touch_kernel(const float *__restrict__ depth, int H, int W, Pose34d c2w, Intr K, int stride, float depth_trunc,
                  Grid g, uint32_t *__restrict__ stamp, uint32_t frame, int *__restrict__ work, int *__restrict__ page,
                  int *__restrict__ pool_state) {
    const int sw = (W + stride - 1) / stride, sh = (H + stride - 1) / stride;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sw * sh) return;
    const int i = (s / sw) * stride, j = (s % sw) * stride;
    float d = depth[i * W + j];
    if (d >= depth_trunc) d = 0.0f;
    if (!(d > 0.0f)) return;
    const double unit_len = (double)g.voxel_length * RES, trunc = (double)g.sdf_trunc;
    const double z = (double)d, x = (j - K.cx) * z / K.fx, y = (i - K.cy) * z / K.fy;
    int lo[3], hi[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double p = c2w.m[4 * r] * x + c2w.m[4 * r + 1] * y + c2w.m[4 * r + 2] * z + c2w.m[4 * r + 3];
        lo[r] = (int)floor((p - trunc) / unit_len);
        hi[r] = (int)floor((p + trunc) / unit_len);
    }
    for (int ux = lo[0]; ux <= hi[0]; ++ux)
        for (int uy = lo[1]; uy <= hi[1]; ++uy)
            for (int uz = lo[2]; uz <= hi[2]; ++uz) {
                const long long u = unit_index(g, ux, uy, uz);
                if (u < 0) continue;
                // the first thread to stamp a unit this frame queues it; the list order varies from run to run but the
                // units are disjoint, so the result does not.  Exactly one thread per (unit, frame) gets here, so the
                // page entry needs no atomic; which pool block a unit receives varies, its contents do not.
                if (atomicExch(stamp + u, frame) != frame) {
                    int pg = page[u];
                    if (pg == 0) {
                        const int slot = atomicAdd(pool_state, 1);
                        if (slot < pool_state[1]) page[u] = pg = slot + 1;
                        else atomicAdd(pool_state + 2, 1);                 // pool exhausted: the unit stays closed
                    }
                    if (pg > 0) work[1 + atomicAdd(work, 1)] = (int)u;
                }
            }
}

// UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier over the frame's work list (work[0] units at
// work[1..]); a fixed grid of CTAs strides over the list, so the load is balanced and no empty CTA is launched.
// Per unit the CTA sweeps lx = 0..15; thread (ly, lz) owns one voxel of the 16x16 slab, so the 256 threads read and
// write 2 KB of contiguous (tsdf, weight) pairs.  Open3D walks z incrementally (pc += R[:,2] * voxel_length per
// step); each thread replays its lz additions so that the rounding is the same.
__global__ void __launch_bounds__(256)
tsdf_integrate_kernel(const float *__restrict__ depth, const float *__restrict__ rgb, int H, int W, Pose34f w2c, Intr Kd,
                      float depth_trunc, Grid g, const int *__restrict__ work, const int *__restrict__ page,
                      float2 *__restrict__ vol, float *__restrict__ color) {
    const int n_open = work[0];
    const float fx = (float)Kd.fx, fy = (float)Kd.fy, cx = (float)Kd.cx, cy = (float)Kd.cy;
    const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
    const float vl = g.voxel_length, half = vl * 0.5f, trunc = g.sdf_trunc, trunc_inv = 1.0f / trunc;
    const float safe_w = (float)W - 0.0001f, safe_h = (float)H - 0.0001f, unit_len = vl * RES;
    const int ly = threadIdx.x >> 4, lz = threadIdx.x & 15;
    const float inc[3] = {w2c.m[2] * vl, w2c.m[6] * vl, w2c.m[10] * vl};
    for (int k = blockIdx.x; k < n_open; k += gridDim.x) {
        const long long u = work[1 + k];
        const long long blk = page[u] - 1;                          // pool block of the unit (queued units always have one)
        const int ux = (int)(u % g.nx) + g.ox, uy = (int)((u / g.nx) % g.ny) + g.oy, uz = (int)(u / ((long long)g.nx * g.ny)) + g.oz;
        const float p0y = half + vl * (float)ly + (float)uy * unit_len, p0z = half + (float)uz * unit_len;
        for (int lx = 0; lx < RES; ++lx) {
            const float p0x = half + vl * (float)lx + (float)ux * unit_len;
            float pc[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) pc[r] = w2c.m[4 * r] * p0x + w2c.m[4 * r + 1] * p0y + w2c.m[4 * r + 2] * p0z + w2c.m[4 * r + 3];
            for (int s = 0; s < lz; ++s) {
#pragma unroll
                for (int r = 0; r < 3; ++r) pc[r] += inc[r];
            }
            if (!(pc[2] > 0.0f)) continue;
            const float u_f = pc[0] * fx / pc[2] + cx + 0.5f, v_f = pc[1] * fy / pc[2] + cy + 0.5f;
            if (!(u_f >= 0.0001f && u_f < safe_w && v_f >= 0.0001f && v_f < safe_h)) continue;
            const int pu = (int)u_f, pv = (int)v_f;
            float d = __ldg(depth + pv * W + pu);
            if (d >= depth_trunc) d = 0.0f;
            if (!(d > 0.0f)) continue;
            const float xx = ((float)pu - cx) * inv_fx, yy = ((float)pv - cy) * inv_fy;
            const float mult = sqrtf(xx * xx + yy * yy + 1.0f);
            const float sdf = (d - pc[2]) * mult;
            if (!(sdf > -trunc)) continue;
            const float tsdf = fminf(1.0f, sdf * trunc_inv);
            const size_t v = (size_t)blk * UNIT_VOX + (size_t)lx * (RES * RES) + threadIdx.x;
            float2 fw = vol[v];
            const float w = fw.y;
            fw.x = (fw.x * w + tsdf) / (w + 1.0f);
            if (color && rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float c8 = rintf((__ldg(rgb + (size_t)(pv * W + pu) * 3 + c) + 1.0f) * 127.5f);
                    color[3 * v + c] = (color[3 * v + c] * w + c8) / (w + 1.0f);
                }
            }
            fw.y = w + 1.0f;
            vol[v] = fw;
        }
    }
}

__device__ __forceinline__ float2 fetch(const Grid &g, const int *__restrict__ page, const float2 *__restrict__ vol,
                                        int gx, int gy, int gz) {
    const long long u = unit_index(g, gx >> 4, gy >> 4, gz >> 4);
    const int pg = u < 0 ? 0 : __ldg(page + u);
    if (pg <= 0) return make_float2(0.0f, 0.0f);
    return __ldg(vol + (size_t)(pg - 1) * UNIT_VOX + (((gx & 15) * RES) + (gy & 15)) * RES + (gz & 15));
}

// Target depth by ray casting; one thread per pixel, 8x8 pixel tiles (a warp = 8x4 pixels) so that neighbouring rays
// share voxels in L1/L2 and diverge little.
__global__ void __launch_bounds__(64)
tsdf_raycast_kernel(Grid g, const int *__restrict__ page, const float2 *__restrict__ vol, Pose34f c2w, Intr Kd,
                    float pixel_center, int H, int W, float z_near, float z_far, float step_vox, float *__restrict__ out) {
    const int u = blockIdx.x * 8 + (threadIdx.x & 7), v = blockIdx.y * 8 + (threadIdx.x >> 3);
    if (u >= W || v >= H) return;
    const float fx = (float)Kd.fx, fy = (float)Kd.fy, cx = (float)Kd.cx, cy = (float)Kd.cy;
    const float vl = g.voxel_length, inv_vl = 1.0f / vl, dt = step_vox * vl;
    const float dc[3] = {((float)u + pixel_center - cx) / fx, ((float)v + pixel_center - cy) / fy, 1.0f};
    float dw[3], ow[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        dw[r] = (c2w.m[4 * r] * dc[0] + c2w.m[4 * r + 1] * dc[1] + c2w.m[4 * r + 2] * dc[2]) * inv_vl;
        ow[r] = c2w.m[4 * r + 3] * inv_vl - 0.5f;
    }
    // observed free space (tsdf clamped at +1: every integrating view saw the surface >= sdf_trunc further along its
    // ray) is crossed in coarse steps of 0.8 sdf_trunc of ray length; a sign change found by a coarse step is re-walked
    const float coarse = 0.8f * g.sdf_trunc / sqrtf(dc[0] * dc[0] + dc[1] * dc[1] + 1.0f);
    float t = z_near, t_prev = 0.0f, f_prev = 0.0f, hit = 0.0f, fine_until = -1.0f;
    bool prev_valid = false;
    int guard = 0;
    while (t <= z_far && guard++ < 100000) {
        const float p[3] = {ow[0] + t * dw[0], ow[1] + t * dw[1], ow[2] + t * dw[2]};
        const float fl[3] = {floorf(p[0]), floorf(p[1]), floorf(p[2])};
        const int b[3] = {(int)fl[0], (int)fl[1], (int)fl[2]};
        const long long unit = unit_index(g, b[0] >> 4, b[1] >> 4, b[2] >> 4);
        const int pg = unit < 0 ? 0 : __ldg(page + unit);
        if (pg <= 0) {
            float t_exit = INFINITY;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (dw[r] == 0.0f) continue;
                const float edge = (float)(((b[r] >> 4) + (dw[r] > 0.0f ? 1 : 0)) * RES);
                const float te = (edge - ow[r]) / dw[r];
                if (te < t_exit) t_exit = te;
            }
            const float t_next = t + dt;
            t = (t_exit > t_next && t_exit < INFINITY) ? t_exit : t_next;
            prev_valid = false;
            continue;
        }
        const float a[3] = {p[0] - fl[0], p[1] - fl[1], p[2] - fl[2]};
        // the 8 trilinear corners: all loads are issued before any is used (the kernel is L2-latency bound); when the
        // 2x2x2 cell lies inside the base unit (82 % of the samples) they are 8 fixed offsets from one address
        float2 fw[8];
        const int l0 = b[0] & 15, l1 = b[1] & 15, l2 = b[2] & 15;
        if (l0 < 15 && l1 < 15 && l2 < 15) {
            const float2 *cell = vol + (size_t)(pg - 1) * UNIT_VOX + ((l0 * RES) + l1) * RES + l2;
#pragma unroll
            for (int c = 0; c < 8; ++c) fw[c] = __ldg(cell + (c & 1) * RES * RES + ((c >> 1) & 1) * RES + (c >> 2));
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) fw[c] = fetch(g, page, vol, b[0] + (c & 1), b[1] + ((c >> 1) & 1), b[2] + (c >> 2));
        }
        float f = 0.0f;
        bool valid = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ix = c & 1, iy = (c >> 1) & 1, iz = c >> 2;
            valid = valid && (fw[c].y > 0.0f);
            const float wx = ix ? a[0] : 1.0f - a[0], wy = iy ? a[1] : 1.0f - a[1], wz = iz ? a[2] : 1.0f - a[2];
            f += fw[c].x * (wx * wy * wz);
        }
        if (valid && prev_valid && f_prev > 0.0f && f <= 0.0f) {
            if (t - t_prev > 1.5f * dt) {               // overshot with a coarse step: walk the interval finely
                fine_until = t;
                t = t_prev + dt;
                continue;
            }
            hit = t_prev + (t - t_prev) * (f_prev / (f_prev - f));
            break;
        }
        prev_valid = valid; f_prev = f; t_prev = t;
        t += (valid && f >= 1.0f && t > fine_until && coarse > dt) ? coarse : dt;
    }
    out[v * W + u] = hit;
}

// ScalableTSDFVolume::ExtractPointCloud.  offsets == nullptr: counts[unit] = number of crossings of the unit;
// otherwise write them at offsets[unit] in (lx, ly, lz, axis) order (block-wide exclusive scan).
__global__ void __launch_bounds__(256)
tsdf_extract_kernel(Grid g, const int *__restrict__ page, const float2 *__restrict__ vol, const float *__restrict__ color,
                    long long *__restrict__ counts, const long long *__restrict__ offsets, float *__restrict__ xyz,
                    float *__restrict__ rgb) {
    __shared__ int scan[256];
    const long long u = blockIdx.x;
    const long long blk = page[u] - 1;
    if (blk < 0) {                                    // uniform per CTA
        if (!offsets && threadIdx.x == 0) counts[u] = 0;
        return;
    }
    const float vl = g.voxel_length, half = vl * 0.5f, unit_len = vl * RES;
    const int ux = (int)(u % g.nx) + g.ox, uy = (int)((u / g.nx) % g.ny) + g.oy, uz = (int)(u / ((long long)g.nx * g.ny)) + g.oz;
    const int lx = threadIdx.x >> 4, ly = threadIdx.x & 15;
    const size_t base = (size_t)blk * UNIT_VOX + (size_t)(lx * RES + ly) * RES;
    for (int pass = offsets ? 0 : 1; pass < 2; ++pass) {
        // pass 0 (fill mode only): count, then scan; pass 1: count (count mode) or write (fill mode)
        const bool write = offsets && pass == 1;
        long long at = 0;
        if (write) at = offsets[u] + scan[threadIdx.x];
        int n = 0;
        for (int lz = 0; lz < RES; ++lz) {
            const float2 fw0 = vol[base + lz];
            const float f0 = fw0.x;
            if (!(fw0.y != 0.0f && f0 < 0.98f && f0 >= -0.98f)) continue;
            const int g0[3] = {ux * RES + lx, uy * RES + ly, uz * RES + lz};
            const float p0[3] = {half + vl * (float)lx + (float)ux * unit_len, half + vl * (float)ly + (float)uy * unit_len,
                                 half + vl * (float)lz + (float)uz * unit_len};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const int g1x = g0[0] + (ax == 0), g1y = g0[1] + (ax == 1), g1z = g0[2] + (ax == 2);
                const long long u1 = unit_index(g, g1x >> 4, g1y >> 4, g1z >> 4);
                const long long blk1 = u1 < 0 ? -1 : (long long)page[u1] - 1;
                if (blk1 < 0) continue;
                const size_t v1 = (size_t)blk1 * UNIT_VOX + (((g1x & 15) * RES) + (g1y & 15)) * RES + (g1z & 15);
                const float2 fw1 = vol[v1];
                const float f1 = fw1.x;
                if (!(fw1.y != 0.0f && f1 < 0.98f && f1 >= -0.98f && f0 * f1 < 0.0f)) continue;
                if (write) {
                    const float r0 = fabsf(f0), r1 = fabsf(f1);
                    float p[3] = {p0[0], p0[1], p0[2]};
                    p[ax] = (p0[ax] * r1 + (p0[ax] + vl) * r0) / (r0 + r1);
                    const size_t v0 = base + lz;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        xyz[3 * (at + n) + c] = p[c];
                        rgb[3 * (at + n) + c] = color ? ((color[3 * v0 + c] * r1 + color[3 * v1 + c] * r0) / (r0 + r1)) / 255.0f : 0.0f;
                    }
                }
                ++n;
            }
        }
        if (write) break;
        // block-wide exclusive scan of the per-thread counts (thread order = (lx, ly) order)
        scan[threadIdx.x] = n;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            const int add = threadIdx.x >= off ? scan[threadIdx.x - off] : 0;
            __syncthreads();
            scan[threadIdx.x] += add;
            __syncthreads();
        }
        if (!offsets) {
            if (threadIdx.x == 255) counts[u] = scan[255];
            return;
        }
        const int incl = scan[threadIdx.x];
        __syncthreads();
        scan[threadIdx.x] = incl - n;
        __syncthreads();
    }
}

int check_grid(const char *what, int nx, int ny, int nz, float voxel_length, float sdf_trunc) {
    SGAM_REQUIRE(nx > 0 && ny > 0 && nz > 0 && (long long)nx * ny * nz < (1ll << 31), "%s: bad unit grid %d x %d x %d", what, nx, ny, nz);
    SGAM_REQUIRE(voxel_length > 0.0f && sdf_trunc > 0.0f, "%s: voxel_length and sdf_trunc must be positive", what);
    return SGAM_OK;
}

}  // namespace

extern "C" size_t sgam_tsdf_volume_bytes(int nx, int ny, int nz, int with_color) {
    return (size_t)nx * ny * nz * UNIT_VOX * (with_color ? 5 : 2) * sizeof(float);
}
extern "C" size_t sgam_tsdf_block_bytes(int with_color) { return (size_t)UNIT_VOX * (with_color ? 5 : 2) * sizeof(float); }

extern "C" int sgam_tsdf_integrate(const float *depth, const float *rgb, int H, int W, const double *host_cam2world,
                                   const float *host_world2cam, const double *host_K, int stride, float depth_trunc,
                                   int ox, int oy, int oz, int nx, int ny, int nz, float voxel_length, float sdf_trunc,
                                   uint32_t *stamp, uint32_t frame, int *work, int *page, int *pool_state, float *vol, float *color,
                                   void *stream) {
    SGAM_REQUIRE(depth && host_cam2world && host_world2cam && host_K && stamp && work && page && pool_state && vol, "tsdf_integrate: null pointer");
    SGAM_REQUIRE(H > 0 && W > 0 && stride > 0 && frame != 0, "tsdf_integrate: bad H/W/stride, or frame stamp 0");
    SGAM_REQUIRE((color == nullptr) == (rgb == nullptr), "tsdf_integrate: rgb and color go together");
    if (int rc = check_grid("tsdf_integrate", nx, ny, nz, voxel_length, sdf_trunc)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Grid g{ox, oy, oz, nx, ny, nz, voxel_length, sdf_trunc};
    Pose34d c2w; Pose34f w2c; Intr K{host_K[0], host_K[1], host_K[2], host_K[3]};
    for (int i = 0; i < 12; ++i) { c2w.m[i] = host_cam2world[i]; w2c.m[i] = host_world2cam[i]; }
    const int samples = ((W + stride - 1) / stride) * ((H + stride - 1) / stride);
    SGAM_CUDA_OK(cudaMemsetAsync(work, 0, sizeof(int), s));
    tsdf_touch_kernel<<<cdiv(samples, 256), 256, 0, s>>>(depth, H, W, c2w, K, stride, depth_trunc, g, stamp, frame, work, page, pool_state);
    SGAM_LAUNCH_OK();
    tsdf_integrate_kernel<<<148 * 4, 256, 0, s>>>(depth, rgb, H, W, w2c, K, depth_trunc, g, work, page, reinterpret_cast<float2 *>(vol), color);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_tsdf_raycast(const int *page, const float *vol, int ox, int oy, int oz, int nx, int ny, int nz,
                                 float voxel_length, float sdf_trunc, const float *host_cam2world, const double *host_K,
                                 float pixel_center, int H, int W, float z_near, float z_far, float step_vox, float *out,
                                 void *stream) {
    SGAM_REQUIRE(page && vol && host_cam2world && host_K && out, "tsdf_raycast: null pointer");
    SGAM_REQUIRE(H > 0 && W > 0 && step_vox > 0.0f && z_far >= z_near && z_near >= 0.0f, "tsdf_raycast: bad H/W/step/z range");
    if (int rc = check_grid("tsdf_raycast", nx, ny, nz, voxel_length, sdf_trunc)) return rc;
    Grid g{ox, oy, oz, nx, ny, nz, voxel_length, sdf_trunc};
    Pose34f c2w; Intr K{host_K[0], host_K[1], host_K[2], host_K[3]};
    for (int i = 0; i < 12; ++i) c2w.m[i] = host_cam2world[i];
    tsdf_raycast_kernel<<<dim3(cdiv(W, 8), cdiv(H, 8)), 64, 0, (cudaStream_t)stream>>>(
        g, page, reinterpret_cast<const float2 *>(vol), c2w, K, pixel_center, H, W, z_near, z_far, step_vox, out);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}

extern "C" int sgam_tsdf_extract(const int *page, const float *vol, const float *color, int ox, int oy, int oz,
                                 int nx, int ny, int nz, float voxel_length, float sdf_trunc, long long *unit_counts,
                                 const long long *unit_offsets, float *xyz, float *rgb, void *stream) {
    SGAM_REQUIRE(page && vol, "tsdf_extract: null pointer");
    SGAM_REQUIRE(unit_offsets ? (xyz && rgb) : (unit_counts != nullptr), "tsdf_extract: count pass needs unit_counts, fill pass needs xyz and rgb");
    if (int rc = check_grid("tsdf_extract", nx, ny, nz, voxel_length, sdf_trunc)) return rc;
    Grid g{ox, oy, oz, nx, ny, nz, voxel_length, sdf_trunc};
    tsdf_extract_kernel<<<(unsigned)((long long)nx * ny * nz), 256, 0, (cudaStream_t)stream>>>(
        g, page, reinterpret_cast<const float2 *>(vol), color, unit_counts, unit_offsets, xyz, rgb);
    SGAM_LAUNCH_OK();
    return SGAM_OK;
}
