"""CPU checks of the TSDF oracle (oracle/csrc/tsdf_oracle.c) on analytic scenes.

The reference delegates RGB-D integration to open3d==0.15.2 (sgam/inference_pipeline.py:119-131, 745-838), which is
absent and has no golden vectors: parity with Open3D is UNPINNED.  What is pinned here is that the restatement of its
published algorithm behaves like a TSDF: surfaces come back where they were put, weights count observations, units
open only around observed surface samples, and the extracted cloud lies on the zero level set."""
import numpy as np
import pytest

from oracle import native

K4 = (248.88887, 248.88887, 64.0, 64.0)
H = W = 128


def plane_depth(z0):
    return np.full((H, W), z0, np.float32)


def make_volume(vox=0.01, trunc=0.03, z0=2.0):
    # box around a fronto-parallel plane at z0 seen from the identity pose
    half = z0 * 64.0 / K4[0] + 0.2
    unit = 16 * vox
    lo = np.floor(np.array([-half, -half, z0 - 0.3]) / unit).astype(int)
    hi = np.floor(np.array([half, half, z0 + 0.3]) / unit).astype(int)
    return native.TsdfVolume(vox, trunc, lo, hi - lo + 1)


def test_plane_is_recovered_from_the_integrating_view_and_from_a_shifted_view():
    z0 = 2.0
    vol = make_volume(z0=z0)
    rgb = np.full((H, W, 3), 0.5, np.float32)
    vol.integrate(plane_depth(z0), rgb, K4, np.eye(4))
    d = vol.render_depth(K4, np.eye(4), H, W, z_far=4.0)
    inner = d[8:-8, 8:-8]
    assert (inner > 0).all()
    assert np.abs(inner - z0).max() < 0.25 * 0.01                    # a quarter of a voxel
    # camera moved 5 cm sideways and 10 cm towards the plane: view-space depth of the same plane is z0 - 0.1
    T = np.eye(4); T[:3, 3] = [-0.05, 0.0, -0.1]
    d2 = vol.render_depth(K4, T, H, W, z_far=4.0)
    seen = d2[16:-16, 16:-16]
    assert (seen > 0).mean() > 0.99
    assert np.abs(seen[seen > 0] - (z0 - 0.1)).max() < 0.5 * 0.01


def test_weights_count_observations_and_units_open_only_near_the_surface():
    z0 = 2.0
    vol = make_volume(z0=z0)
    for _ in range(3):
        vol.integrate(plane_depth(z0), None, K4, np.eye(4))
    w = vol.vol[..., 1]
    assert set(np.unique(w)) <= {0.0, 3.0}
    assert (w == 3.0).sum() > 0
    opened = np.nonzero(vol.stamp)[0]
    assert len(opened) > 0 and (vol.stamp[opened] == 3).all()
    g = vol.grid
    uz = opened // (g.nx * g.ny) + g.oz
    unit = 16 * 0.01
    # opened units straddle z0 +/- trunc only
    assert ((uz + 1) * unit > z0 - 0.03 - 1e-6).all() and (uz * unit < z0 + 0.03 + 1e-6).all()
    # tsdf is the truncated, normalised signed distance along the ray: +1 well in front, negative just behind
    f = vol.vol[..., 0][w > 0]
    assert f.max() <= 1.0 and f.min() > -1.0


def test_depth_trunc_and_invalid_depths_are_ignored():
    vol = make_volume()
    d = plane_depth(2.0)
    d[:, : W // 2] = 25.0                                               # >= depth_trunc (20) -> dropped (:771-772)
    d[: H // 2, W // 2:] = 0.0
    vol.integrate(d, None, K4, np.eye(4))
    out = vol.render_depth(K4, np.eye(4), H, W, z_far=4.0)
    assert (out[:, : W // 2 - 2] == 0).all() and (out[: H // 2 - 2, W // 2 + 2:] == 0).all()
    assert (out[H // 2 + 4: -8, W // 2 + 4: -8] > 0).all()


def test_extracted_cloud_lies_on_the_surface_with_the_integrated_colour():
    z0 = 2.0
    vol = make_volume(z0=z0)
    rgb = np.empty((H, W, 3), np.float32)
    rgb[...] = np.array([200, 100, 50]) / 127.5 - 1.0
    vol.integrate(plane_depth(z0), rgb, K4, np.eye(4))
    xyz, col = vol.extract_point_cloud()
    assert len(xyz) > 1000
    # zero crossings along z sit on the plane; crossings along x / y can only occur within a voxel of it
    assert np.abs(xyz[:, 2] - z0).max() < 0.01
    assert np.abs(col * 255 - np.array([200, 100, 50])).max() < 1e-3


def test_tilted_plane_depth_error_is_sub_voxel():
    # plane n.x = c seen from the identity pose: depth(u,v) = c / (n . ray(u,v))
    n = np.array([0.2, -0.1, 1.0]); n /= np.linalg.norm(n)
    c = 2.0
    vs, us = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    rays = np.stack([(us - K4[2]) / K4[0], (vs - K4[3]) / K4[1], np.ones_like(us, float)], -1)
    depth = (c / (rays @ n)).astype(np.float32)
    vox, unit = 0.01, 0.16
    lo = np.floor(np.array([-0.8, -0.8, 1.5]) / unit).astype(int)
    hi = np.floor(np.array([0.8, 0.8, 2.6]) / unit).astype(int)
    vol = native.TsdfVolume(vox, 0.03, lo, hi - lo + 1, with_color=False)
    vol.integrate(depth, None, K4, np.eye(4))
    out = vol.render_depth(K4, np.eye(4), H, W, pixel_center=0.0, z_far=4.0)
    inner = (slice(8, -8), slice(8, -8))
    assert (out[inner] > 0).mean() > 0.99
    err = np.abs(out[inner] - depth[inner])[out[inner] > 0]
    assert err.max() < 1.0 * vox and err.mean() < 0.3 * vox


def test_empty_frame_opens_nothing_and_renders_nothing():
    vol = make_volume()
    vol.integrate(np.zeros((H, W), np.float32), None, K4, np.eye(4))
    assert not vol.stamp.any() and not vol.vol.any()
    assert not vol.render_depth(K4, np.eye(4), H, W, z_far=4.0).any()
    xyz, col = vol.extract_point_cloud()
    assert xyz.shape == (0, 3) and col.shape == (0, 3)


def test_surface_outside_the_dense_box_is_dropped():
    """Re-design (1): units outside the caller's box are ignored (Open3D's hash would allocate them)."""
    vol = make_volume(z0=2.0)
    vol.integrate(plane_depth(6.0), None, K4, np.eye(4))               # the plane lies far behind the box
    assert not vol.stamp.any()
