/* CPU oracle for the RGB-D integration row of the SGAM hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED.  The reference delegates this row to a third-party wheel that is neither in
 * /root/reference nor installable here: open3d==0.15.2 (requirement.txt:8), call sites
 * sgam/inference_pipeline.py:119-131 (ScalableTSDFVolume(voxel_length, sdf_trunc, RGB8)),
 * :762-777 (RGBDImage.create_from_color_and_depth(depth_scale=1, depth_trunc=20), volume.integrate),
 * :786-827 (extract_triangle_mesh + OffscreenRenderer.render_to_depth_image(z_in_view_space=True), inf -> 0)
 * and :446-447 (volume.extract_point_cloud).  No golden vector of it exists, so this file restates the
 * PUBLISHED algorithm of that version:
 *
 *   ScalableTSDFVolume::Integrate        -- depth -> point cloud with depth_sampling_stride 4; every 16^3-voxel
 *                                           volume unit whose index lies in floor((p -/+ sdf_trunc) / unit_length)
 *                                           is opened; every opened unit integrates the frame;
 *   UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier
 *                                        -- voxel centre -> camera, u = (int)(x fx / z + cx + 0.5), sdf =
 *                                           (d - z) * sqrt(1 + xx^2 + yy^2), running average of min(1, sdf/trunc)
 *                                           and of the colour where sdf > -trunc, weight += 1;
 *   ScalableTSDFVolume::ExtractPointCloud-- zero crossings along +x, +y, +z between voxels with weight != 0 and
 *                                           tsdf in [-0.98, 0.98), linear interpolation of position and colour.
 *
 * Two deliberate re-designs, shared with the CUDA path and declared in DESIGN.md:
 *   (1) the unordered hash of volume units becomes a DENSE grid of units over a caller-given box (units outside
 *       it are dropped);
 *   (2) the target depth is ray-cast from the TSDF (trilinear samples, first + -> - crossing) instead of
 *       marching cubes + Filament rasterisation.
 * The GPU kernels are held bit-exact to THIS file (tests/test_gpu_tsdf.py); this file is held to analytic
 * scenes (tests/test_oracle_tsdf.py).
 *
 * Layout: unit (ux,uy,uz) of a grid dims = (nx,ny,nz) has linear index (uz*ny + uy)*nx + ux; voxel (lx,ly,lz) of
 * a unit sits at (lx*16 + ly)*16 + lz (Open3D's IndexOf); vol = [units][4096][2] (tsdf, weight);
 * color = [units][4096][3] (0..255, float); stamp[unit] = 0 (never opened) or the frame counter of its last
 * integration.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RES 16
#define UNIT_VOX (RES * RES * RES)

typedef struct {
    int ox, oy, oz;          /* unit index of the grid corner (world unit i covers [i*U, (i+1)*U)) */
    int nx, ny, nz;          /* grid size in units */
    float voxel_length, sdf_trunc;
} tsdf_grid;

static inline long unit_index(const tsdf_grid *g, int ux, int uy, int uz) {
    ux -= g->ox; uy -= g->oy; uz -= g->oz;
    if (ux < 0 || uy < 0 || uz < 0 || ux >= g->nx || uy >= g->ny || uz >= g->nz) return -1;
    return ((long)uz * g->ny + uy) * g->nx + ux;
}

/* ScalableTSDFVolume::Integrate, first half: open the units around the strided depth samples.
 * cam2world: row-major 3x4 double = inverse extrinsic; K = fx, fy, cx, cy (double). */
void oracle_tsdf_touch(const float *depth, int H, int W, const double *cam2world, const double *K, int stride,
                       float depth_trunc, const tsdf_grid *g, uint32_t *stamp, uint32_t frame) {
    const double unit_len = (double)g->voxel_length * RES, trunc = (double)g->sdf_trunc;
    for (int i = 0; i < H; i += stride) for (int j = 0; j < W; j += stride) {
        float d = depth[i * W + j];
        if (d >= depth_trunc) d = 0.0f;
        if (!(d > 0.0f)) continue;
        const double z = (double)d, x = (j - K[2]) * z / K[0], y = (i - K[3]) * z / K[1];
        double p[3];
        for (int r = 0; r < 3; ++r)
            p[r] = cam2world[4 * r] * x + cam2world[4 * r + 1] * y + cam2world[4 * r + 2] * z + cam2world[4 * r + 3];
        int lo[3], hi[3];
        for (int r = 0; r < 3; ++r) {
            lo[r] = (int)floor((p[r] - trunc) / unit_len);
            hi[r] = (int)floor((p[r] + trunc) / unit_len);
        }
        for (int ux = lo[0]; ux <= hi[0]; ++ux) for (int uy = lo[1]; uy <= hi[1]; ++uy) for (int uz = lo[2]; uz <= hi[2]; ++uz) {
            const long u = unit_index(g, ux, uy, uz);
            if (u >= 0) stamp[u] = frame;
        }
    }
}

/* UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier over every unit opened for `frame`.
 * world2cam: row-major 3x4 float (extrinsic.cast<float>()); rgb [H,W,3] fp32 in [-1,1] on the uint8 lattice
 * (the resident frame store) or NULL. */
void oracle_tsdf_integrate(const float *depth, const float *rgb, int H, int W, const float *world2cam, const double *K,
                           float depth_trunc, const tsdf_grid *g, const uint32_t *stamp, uint32_t frame,
                           float *vol, float *color) {
    const float fx = (float)K[0], fy = (float)K[1], cx = (float)K[2], cy = (float)K[3];
    const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
    const float vl = g->voxel_length, half = vl * 0.5f, trunc = g->sdf_trunc, trunc_inv = 1.0f / trunc;
    const float safe_w = (float)W - 0.0001f, safe_h = (float)H - 0.0001f;
    const float unit_len = vl * RES;
    const long n_units = (long)g->nx * g->ny * g->nz;
    for (long u = 0; u < n_units; ++u) {
        if (stamp[u] != frame) continue;
        const int ux = (int)(u % g->nx) + g->ox, uy = (int)((u / g->nx) % g->ny) + g->oy, uz = (int)(u / ((long)g->nx * g->ny)) + g->oz;
        const float org[3] = {(float)ux * unit_len, (float)uy * unit_len, (float)uz * unit_len};
        for (int lx = 0; lx < RES; ++lx) for (int ly = 0; ly < RES; ++ly) {
            const float p0[3] = {half + vl * (float)lx + org[0], half + vl * (float)ly + org[1], half + org[2]};
            float pc[3];
            for (int r = 0; r < 3; ++r)
                pc[r] = world2cam[4 * r] * p0[0] + world2cam[4 * r + 1] * p0[1] + world2cam[4 * r + 2] * p0[2] + world2cam[4 * r + 3];
            for (int lz = 0; lz < RES; ++lz) {
                if (lz > 0) for (int r = 0; r < 3; ++r) pc[r] += world2cam[4 * r + 2] * vl;
                if (!(pc[2] > 0.0f)) continue;
                const float u_f = pc[0] * fx / pc[2] + cx + 0.5f, v_f = pc[1] * fy / pc[2] + cy + 0.5f;
                if (!(u_f >= 0.0001f && u_f < safe_w && v_f >= 0.0001f && v_f < safe_h)) continue;
                const int pu = (int)u_f, pv = (int)v_f;
                float d = depth[pv * W + pu];
                if (d >= depth_trunc) d = 0.0f;
                if (!(d > 0.0f)) continue;
                const float xx = ((float)pu - cx) * inv_fx, yy = ((float)pv - cy) * inv_fy;
                const float mult = sqrtf(xx * xx + yy * yy + 1.0f);
                const float sdf = (d - pc[2]) * mult;
                if (!(sdf > -trunc)) continue;
                const float tsdf = fminf(1.0f, sdf * trunc_inv);
                const size_t v = (size_t)u * UNIT_VOX + (lx * RES + ly) * RES + lz;
                const float w = vol[2 * v + 1];
                vol[2 * v] = (vol[2 * v] * w + tsdf) / (w + 1.0f);
                if (color && rgb) for (int c = 0; c < 3; ++c) {
                    const float c8 = rintf((rgb[(pv * W + pu) * 3 + c] + 1.0f) * 127.5f);
                    color[3 * v + c] = (color[3 * v + c] * w + c8) / (w + 1.0f);
                }
                vol[2 * v + 1] = w + 1.0f;
            }
        }
    }
}

/* (tsdf, weight) of global voxel (gx,gy,gz); weight 0 when the unit is outside the grid or never opened */
static inline void fetch(const tsdf_grid *g, const uint32_t *stamp, const float *vol, int gx, int gy, int gz, float *f, float *w) {
    const long u = unit_index(g, gx >> 4, gy >> 4, gz >> 4);
    if (u < 0 || stamp[u] == 0) { *f = 0.0f; *w = 0.0f; return; }
    const size_t v = (size_t)u * UNIT_VOX + (((gx & 15) * RES) + (gy & 15)) * RES + (gz & 15);
    *f = vol[2 * v]; *w = vol[2 * v + 1];
}

/* Target depth by ray casting (re-design (2) above).  cam2world: row-major 3x4 float; pixel (u,v) looks along
 * ((u + pc - cx)/fx, (v + pc - cy)/fy, 1) so the ray parameter IS the view-space z the reference asks for
 * (render_to_depth_image(z_in_view_space=True)); no hit -> 0 (:827 inf -> 0).
 * Samples every `step_vox` voxel lengths of z in [z_near, z_far]; inside a never-opened unit the ray jumps to the
 * unit's exit; through observed free space it takes coarse steps.  A sample is valid when all 8 trilinear corners have
 * weight > 0. */
void oracle_tsdf_raycast(const tsdf_grid *g, const uint32_t *stamp, const float *vol, const float *cam2world,
                         const double *K, float pixel_center, int H, int W, float z_near, float z_far, float step_vox,
                         float *out) {
    const float fx = (float)K[0], fy = (float)K[1], cx = (float)K[2], cy = (float)K[3];
    const float vl = g->voxel_length, inv_vl = 1.0f / vl, dt = step_vox * vl;
    for (int v = 0; v < H; ++v) for (int u = 0; u < W; ++u) {
        const float dc[3] = {((float)u + pixel_center - cx) / fx, ((float)v + pixel_center - cy) / fy, 1.0f};
        float dw[3], ow[3];
        for (int r = 0; r < 3; ++r) {
            dw[r] = (cam2world[4 * r] * dc[0] + cam2world[4 * r + 1] * dc[1] + cam2world[4 * r + 2] * dc[2]) * inv_vl;
            ow[r] = cam2world[4 * r + 3] * inv_vl - 0.5f;          /* voxel coordinates: voxel i is centred at i */
        }
        /* observed free space (tsdf clamped at +1) is crossed in coarse steps of 0.8 sdf_trunc of ray length; a sign
         * change found by a coarse step is re-walked finely */
        const float coarse = 0.8f * g->sdf_trunc / sqrtf(dc[0] * dc[0] + dc[1] * dc[1] + 1.0f);
        float t = z_near, t_prev = 0.0f, f_prev = 0.0f, hit = 0.0f, fine_until = -1.0f;
        int prev_valid = 0, guard = 0;
        while (t <= z_far && guard++ < 100000) {
            const float p[3] = {ow[0] + t * dw[0], ow[1] + t * dw[1], ow[2] + t * dw[2]};
            const float fl[3] = {floorf(p[0]), floorf(p[1]), floorf(p[2])};
            const int b[3] = {(int)fl[0], (int)fl[1], (int)fl[2]};
            const long unit = unit_index(g, b[0] >> 4, b[1] >> 4, b[2] >> 4);
            if (unit < 0 || stamp[unit] == 0) {
                /* jump to where the ray leaves this 16-voxel unit (at least one step) */
                float t_exit = INFINITY;
                for (int r = 0; r < 3; ++r) {
                    if (dw[r] == 0.0f) continue;
                    const float edge = (float)(((b[r] >> 4) + (dw[r] > 0.0f ? 1 : 0)) * RES);
                    const float te = (edge - ow[r]) / dw[r];
                    if (te < t_exit) t_exit = te;
                }
                const float t_next = t + dt;
                t = (t_exit > t_next && t_exit < INFINITY) ? t_exit : t_next;
                prev_valid = 0;
                continue;
            }
            const float a[3] = {p[0] - fl[0], p[1] - fl[1], p[2] - fl[2]};
            float f = 0.0f;
            int valid = 1;
            for (int c = 0; c < 8; ++c) {
                const int ix = c & 1, iy = (c >> 1) & 1, iz = c >> 2;
                float fv, wv;
                fetch(g, stamp, vol, b[0] + ix, b[1] + iy, b[2] + iz, &fv, &wv);
                if (!(wv > 0.0f)) valid = 0;
                const float wx = ix ? a[0] : 1.0f - a[0], wy = iy ? a[1] : 1.0f - a[1], wz = iz ? a[2] : 1.0f - a[2];
                f += fv * (wx * wy * wz);
            }
            if (valid && prev_valid && f_prev > 0.0f && f <= 0.0f) {
                if (t - t_prev > 1.5f * dt) { fine_until = t; t = t_prev + dt; continue; }
                hit = t_prev + (t - t_prev) * (f_prev / (f_prev - f));
                break;
            }
            prev_valid = valid; f_prev = f; t_prev = t;
            t += (valid && f >= 1.0f && t > fine_until && coarse > dt) ? coarse : dt;
        }
        out[v * W + u] = hit;
    }
}

/* ScalableTSDFVolume::ExtractPointCloud.  Order: units ascending, then lx, ly, lz, axis.  With xyz == NULL only
 * counts.  Returns the number of points; xyz [n,3] float, rgb [n,3] float in [0,1]. */
long oracle_tsdf_extract(const tsdf_grid *g, const uint32_t *stamp, const float *vol, const float *color,
                         float *xyz, float *rgb) {
    const float vl = g->voxel_length, half = vl * 0.5f, unit_len = vl * RES;
    const long n_units = (long)g->nx * g->ny * g->nz;
    long n = 0;
    for (long u = 0; u < n_units; ++u) {
        if (stamp[u] == 0) continue;
        const int ux = (int)(u % g->nx) + g->ox, uy = (int)((u / g->nx) % g->ny) + g->oy, uz = (int)(u / ((long)g->nx * g->ny)) + g->oz;
        for (int lx = 0; lx < RES; ++lx) for (int ly = 0; ly < RES; ++ly) for (int lz = 0; lz < RES; ++lz) {
            const size_t v0 = (size_t)u * UNIT_VOX + (lx * RES + ly) * RES + lz;
            const float f0 = vol[2 * v0], w0 = vol[2 * v0 + 1];
            if (!(w0 != 0.0f && f0 < 0.98f && f0 >= -0.98f)) continue;
            const int g0[3] = {ux * RES + lx, uy * RES + ly, uz * RES + lz};
            const float p0[3] = {half + vl * (float)lx + (float)ux * unit_len, half + vl * (float)ly + (float)uy * unit_len,
                                 half + vl * (float)lz + (float)uz * unit_len};
            for (int ax = 0; ax < 3; ++ax) {
                int g1[3] = {g0[0], g0[1], g0[2]};
                g1[ax] += 1;
                const long u1 = unit_index(g, g1[0] >> 4, g1[1] >> 4, g1[2] >> 4);
                if (u1 < 0 || stamp[u1] == 0) continue;
                const size_t v1 = (size_t)u1 * UNIT_VOX + (((g1[0] & 15) * RES) + (g1[1] & 15)) * RES + (g1[2] & 15);
                const float f1 = vol[2 * v1], w1 = vol[2 * v1 + 1];
                if (!(w1 != 0.0f && f1 < 0.98f && f1 >= -0.98f && f0 * f1 < 0.0f)) continue;
                if (xyz) {
                    const float r0 = fabsf(f0), r1 = fabsf(f1);
                    float p[3] = {p0[0], p0[1], p0[2]};
                    p[ax] = (p0[ax] * r1 + (p0[ax] + vl) * r0) / (r0 + r1);
                    for (int c = 0; c < 3; ++c) xyz[3 * n + c] = p[c];
                    for (int c = 0; c < 3; ++c)
                        rgb[3 * n + c] = color ? ((color[3 * v0 + c] * r1 + color[3 * v1 + c] * r0) / (r0 + r1)) / 255.0f : 0.0f;
                }
                ++n;
            }
        }
    }
    return n;
}
