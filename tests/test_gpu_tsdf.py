"""GPU parity of the RGB-D integration kernels (csrc/tsdf.cu) against oracle/csrc/tsdf_oracle.c: bit-exact volume
(stamps, tsdf, weights, colours), bit-exact ray-cast depth, identical extracted cloud; plus the scene loop with
use_rgbd_integration=True running on the device volume (BASELINE.json configs[2]).  Parity with Open3D itself is
unpinned."""
import numpy as np
import pytest
import torch

from oracle import native

pytestmark = pytest.mark.gpu

H = W = 128
K = np.array([[248.88887, 0, 64.0], [0, 248.88887, 64.0], [0, 0, 1.0]])


def scene(seed, n_frames=3):
    """A bumpy surface around z = 2.4 seen from cameras stepping along +y with a 30 degree tilt (GoogleEarth-like)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    frames = []
    c, s = np.cos(0.5), np.sin(0.5)
    for i in range(n_frames):
        depth = (2.4 + 0.5 * np.sin(5 * xx + i) * np.cos(4 * yy) + 0.003 * rng.standard_normal((H, W))).astype(np.float32)
        depth[rng.random((H, W)) < 0.02] = 0.0                           # holes
        depth[:4, :4] = 30.0                                             # beyond depth_trunc
        rgb = (rng.integers(0, 256, (H, W, 3)) / 127.5 - 1.0).astype(np.float32)
        c2w = np.eye(4)
        c2w[:3, :3] = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        c2w[:3, 3] = [0.1 * i, 0.06 * i, 0.0]
        frames.append((depth, rgb, np.linalg.inv(c2w)))
    return frames


def volumes(frames, with_color=True):
    from sgam_neurips22_b200.tsdf import TSDFVolume, frustum_box
    lo, hi = frustum_box(K, [f[2] for f in frames], H, W, 3.3, pad=0.2)
    dev = TSDFVolume(0.01, 0.03, lo, hi, with_color=with_color)
    ref = native.TsdfVolume(0.01, 0.03, dev.origin, dev.dims, with_color=with_color)
    return dev, ref


@pytest.mark.parametrize("with_color", [True, False])
def test_integrate_is_bit_exact(with_color):
    frames = scene(0)
    dev, ref = volumes(frames, with_color)
    for depth, rgb, T in frames:
        dev.integrate(torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), K, T)
        ref.integrate(depth, rgb, (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), T)
    torch.cuda.synchronize()
    assert np.array_equal(dev.stamp.cpu().numpy().view(np.uint32), ref.stamp)
    assert (ref.stamp > 0).sum() > 50
    assert np.array_equal(dev.vol.cpu().numpy(), ref.vol)
    assert (ref.vol[..., 1] > 1).any()                                   # overlapping observations were averaged
    if with_color:
        assert np.array_equal(dev.color.cpu().numpy(), ref.color)
    else:
        assert dev.color is None


def test_raycast_and_extract_match_the_oracle():
    frames = scene(1)
    dev, ref = volumes(frames)
    k4 = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    for depth, rgb, T in frames:
        dev.integrate(torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), K, T)
        ref.integrate(depth, rgb, k4, T)
    # a novel view between / beyond the integrated ones
    c2w = np.linalg.inv(frames[1][2]).copy()
    c2w[:3, 3] += [0.03, 0.05, -0.02]
    T_new = np.linalg.inv(c2w)
    for pc in (0.5, 0.0):
        d_dev = dev.render_depth(K, T_new, H, W, pixel_center=pc, z_far=4.0).cpu().numpy()
        d_ref = ref.render_depth(k4, T_new, H, W, pixel_center=pc, z_far=4.0)
        assert (d_ref > 0).mean() > 0.8
        assert np.array_equal(d_dev, d_ref)
    xyz_d, col_d = dev.extract_point_cloud()
    xyz_r, col_r = ref.extract_point_cloud()
    assert len(xyz_r) > 10000
    assert np.array_equal(xyz_d.cpu().numpy(), xyz_r) and np.array_equal(col_d.cpu().numpy(), col_r)


def test_rendered_depth_tracks_the_integrated_surface():
    """Size-independent property: re-rendering an integrated view returns its own depth to sub-voxel accuracy."""
    frames = scene(2, n_frames=1)
    depth, rgb, T = frames[0]
    dev, _ = volumes(frames, with_color=False)
    dev.integrate(torch.from_numpy(depth).cuda(), None, K, T)
    out = dev.render_depth(K, T, H, W, pixel_center=0.0, z_far=4.0).cpu().numpy()
    ok = (out > 0) & (depth > 0) & (depth < 20)
    assert ok.mean() > 0.7
    err = np.abs(out - depth)[ok]
    assert np.median(err) < 0.01 and np.percentile(err, 90) < 0.03       # 1 voxel / 3 voxels (noisy, holed surface)


def test_scene_loop_with_rgbd_integration_runs_on_the_device_volume(tmp_path, monkeypatch):
    """configs[2]: GoogleEarth loop with use_rgbd_integration=True and NO Open3D: the pipeline's integrated target depth
    equals an oracle volume fed the same frames, the step consumes it (pre-warped get_x), the final cloud is written."""
    from sgam_neurips22_b200 import synthetic
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration
    from sgam_neurips22_b200.model import VQModel
    from sgam_neurips22_b200.tsdf import TSDFVolume
    monkeypatch.chdir(tmp_path)
    ds = "google_earth"
    model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
    rng = np.random.default_rng(3)
    yy, xx = np.meshgrid(np.linspace(0, 1, 256), np.linspace(0, 1, 256), indexing="ij")
    seed = (rng.integers(0, 256, (256, 256, 3)).astype(np.uint8),
            (1.4 + 2.4 * (0.5 + 0.3 * np.sin(3 * xx) * np.cos(2 * yy))).astype(np.float32))
    pipe = InfiniteSceneGeneration(model, ds, seed_frame=seed, output_dim=(4, 1), use_rgbd_integration=True)
    assert isinstance(pipe.volume, TSDFVolume)
    ref = native.TsdfVolume(0.01, 0.03, pipe.volume.origin, pipe.volume.dims)
    k4 = (pipe.K[0, 0], pipe.K[1, 1], pipe.K[0, 2], pipe.K[1, 2])
    captured, inner = {}, pipe.rgbd_integration

    def spy(*a, **kw):
        captured["depth"] = inner(*a, **kw)
        return captured["depth"]
    pipe.rgbd_integration = spy
    for step in range(2):
        tgt = pipe.next_pose(pipe.curr)
        srcs, _ = pipe.get_src_grid_coords(tgt)
        tgt_meta = pipe.transform_grid[tgt[0]][tgt[1]]
        T_tgt = np.eye(4); T_tgt[:3, :3], T_tgt[:3, 3] = tgt_meta["R"], tgt_meta["t"]
        for c in srcs:                                                    # what rgbd_integration will integrate
            n = pipe.transform_grid[c[0]][c[1]]
            rgb, depth = pipe._frames[tuple(c)]
            if tuple(c) == tuple(pipe._ordered_grid_coords[0]):
                depth = pipe._seed_depth_single
            T_src = np.eye(4); T_src[:3, :3], T_src[:3, 3] = n["R"], n["t"]
            ref.integrate(depth.cpu().numpy(), rgb.cpu().numpy(), k4, T_src)
        d_ref = ref.render_depth(k4, T_tgt, 256, 256, z_far=pipe._z_far)
        res = pipe.one_step_prediction(tgt)
        assert np.array_equal(captured["depth"].cpu().numpy(), d_ref)
        assert (d_ref > 0).mean() > 0.5
        x = res["x"].cpu().numpy()
        assert np.array_equal(x[0, 3], native.depth_code(d_ref, (d_ref <= 0).astype(np.uint8), ds))   # holes coded -2
        pipe.curr += 1
    assert np.array_equal(pipe.volume.vol.cpu().numpy(), ref.vol)
    pts, cols = pipe.volume.extract_point_cloud()
    assert len(pts) > 1000 and float(cols.min()) >= 0.0 and float(cols.max()) <= 1.0 + 1e-6


def test_empty_frame_and_empty_cloud():
    """Edge case: a frame without valid depth opens no unit; ray cast and cloud extraction return nothing."""
    from sgam_neurips22_b200.tsdf import TSDFVolume
    dev = TSDFVolume(0.01, 0.03, [-0.5, -0.5, 1.5], [0.5, 0.5, 2.5], with_color=False)
    d = torch.zeros(H, W, device="cuda")
    d[:8] = 40.0                                                         # beyond depth_trunc: ignored as well
    dev.integrate(d, None, K, np.eye(4))
    assert int(dev.stamp.abs().sum()) == 0 and float(dev.vol.abs().sum()) == 0.0
    assert float(dev.render_depth(K, np.eye(4), H, W, z_far=4.0).abs().sum()) == 0.0
    xyz, col = dev.extract_point_cloud()
    assert tuple(xyz.shape) == (0, 3) and tuple(col.shape) == (0, 3)


def test_unit_pool_exhaustion_drops_units_instead_of_faulting():
    """The voxel data lives in a bounded pool behind a page table (<= max_bytes per trajectory).  When the pool runs out the
    remaining units stay closed: nothing is written out of bounds, the dropped units are counted, the opened ones keep
    their exact values, and ray casting / extraction still work."""
    from sgam_neurips22_b200.tsdf import TSDFVolume, frustum_box
    frames = scene(4, n_frames=2)
    lo, hi = frustum_box(K, [f[2] for f in frames], H, W, 3.3, pad=0.2)
    full = TSDFVolume(0.01, 0.03, lo, hi, with_color=False)
    for depth, rgb, T in frames:
        full.integrate(torch.from_numpy(depth).cuda(), None, K, T)
    need = full.units_in_use()
    assert need > 100 and full.dropped_units() == 0 and full.capacity == full.page.numel()     # small box: the pool can hold every unit
    small = TSDFVolume(0.01, 0.03, lo, hi, with_color=False, max_bytes=(need // 2) * 4096 * 8)
    assert small.capacity == need // 2
    for depth, rgb, T in frames:
        small.integrate(torch.from_numpy(depth).cuda(), None, K, T)
    torch.cuda.synchronize()
    assert small.units_in_use() == small.capacity and small.dropped_units() > 0
    assert int((small.page > 0).sum()) == small.capacity and int(small.page.max()) == small.capacity
    d = small.render_depth(K, frames[0][2], H, W, z_far=4.0)
    assert bool(torch.isfinite(d).all())
    xyz, _ = small.extract_point_cloud()
    assert 0 < xyz.shape[0] < full.extract_point_cloud()[0].shape[0]


def test_integrate_once_is_a_switch_and_the_default_reintegrates(tmp_path, monkeypatch):
    """inference_pipeline.py:771-777 fuses every selected source at every step (default here too); integrate_once=True fuses a
    frame only the first time it is selected: the volume then holds weight 1 where the default holds the selection count."""
    from sgam_neurips22_b200 import synthetic
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration
    from sgam_neurips22_b200.model import VQModel
    monkeypatch.chdir(tmp_path)
    ds = "google_earth"
    model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
    rng = np.random.default_rng(5)
    yy, xx = np.meshgrid(np.linspace(0, 1, 256), np.linspace(0, 1, 256), indexing="ij")
    seed = (rng.integers(0, 256, (256, 256, 3)).astype(np.uint8), (2.0 + 0.5 * np.sin(3 * xx) * np.cos(2 * yy)).astype(np.float32))
    wmax = {}
    for once in (False, True):
        pipe = InfiniteSceneGeneration(model, ds, seed_frame=seed, output_dim=(4, 1), use_rgbd_integration=True, integrate_once=once,
                                       output_root=str(tmp_path / f"once{int(once)}"))
        for _ in range(3):
            pipe.one_step_prediction(pipe.next_pose(pipe.curr), save_res_to_disk=False)
            pipe.curr += 1
        wmax[once] = float(pipe.volume.pool_vol[..., 1].max())
        assert pipe.volume.dropped_units() == 0
    assert wmax[True] <= 3.0 and wmax[False] > wmax[True]          # 3 frames fused once each vs the seed fused at every step
