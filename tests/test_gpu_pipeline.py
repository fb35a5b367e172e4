"""GPU tests of the scene loop mirror: one_step_prediction against the oracle,
the device-resident frame store against the files it writes, the grid_res/ layout, the merged point cloud, and the
use_rgbd_integration=True branch (inverse warp + pre-warped get_x) with a supplied target depth."""
import os

import numpy as np
import pytest
import torch

from oracle import model as omodel

pytestmark = pytest.mark.gpu


def seed_frame(rng, lo, hi, res=256):
    rgb = rng.integers(0, 256, (res, res, 3)).astype(np.uint8)
    yy, xx = np.meshgrid(np.linspace(0, 1, res), np.linspace(0, 1, res), indexing="ij")
    depth = (lo + (hi - lo) * (0.5 + 0.3 * np.sin(3 * xx) * np.cos(2 * yy))).astype(np.float32)
    return rgb, depth


@pytest.fixture(scope="module")
def models():
    from sgam_neurips22_b200 import synthetic
    from sgam_neurips22_b200.model import VQModel
    cache = {}

    def get(ds):
        if ds not in cache:
            cache[ds] = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
        return cache[ds]
    return get


def test_one_step_matches_oracle_and_frame_store_matches_disk(models, tmp_path, monkeypatch):
    from PIL import Image
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration
    monkeypatch.chdir(tmp_path)
    ds = "google_earth"
    model = models(ds)
    rng = np.random.default_rng(0)
    pipe = InfiniteSceneGeneration(model, ds, seed_frame=seed_frame(rng, 1.4, 3.8), output_dim=(4, 1))
    tgt = pipe.next_pose(pipe.curr)
    srcs, _ = pipe.get_src_grid_coords(tgt)
    assert srcs == [(0, 0)]
    tgt_meta = pipe.transform_grid[tgt[0]][tgt[1]]
    batch = pipe.prepare_batch_data(tgt_meta, [pipe.transform_grid[c[0]][c[1]] for c in srcs], pipe.num_src)
    batch_np = {k: v.cpu().numpy() for k, v in batch.items() if torch.is_tensor(v)}
    res = pipe.one_step_prediction(tgt)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = omodel.scene_step(sd, batch_np, ds)
    assert np.array_equal(res["x"].cpu().numpy(), ref["x"])                       # splat + depth code: bit-exact
    assert np.array_equal(res["warped_depth"].cpu().numpy(), ref["warped_depth"])
    dec = res["rgbd"].cpu().numpy()
    assert np.linalg.norm(dec - ref["dec"][0]) / np.linalg.norm(ref["dec"][0]) < 1e-3
    assert tuple(res["feature"].shape) == (256, 16, 16) and tuple(res["pre_quantized_features"].shape) == (256, 16, 16)
    # files of the reference layout, and the resident frame == what a reload from disk would give (inference_pipeline.py:534-536)
    base = tmp_path / "grid_res" / "google_earth_seed0"
    for stem in ("im_00001_01_00.png", "dm_00001_01_00.npy", "R_00001_01_00.npy", "t_00001_01_00.npy"):
        assert (base / stem).exists(), stem
    rgb_store, depth_store = pipe._frames[(1, 0)]
    png = np.array(Image.open(base / "im_00001_01_00.png"))
    assert np.array_equal(rgb_store.cpu().numpy(), (png / 127.5 - 1.0).astype(np.float32))
    assert np.array_equal(depth_store.cpu().numpy(), np.load(base / "dm_00001_01_00.npy"))
    d = np.abs(png.astype(int) - ref["rgb_u8"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 5e-3
    assert pipe.transform_grid[1][0]["visited"] and pipe.transform_grid[1][0]["rgb_path"].endswith("im_00001_01_00.png")


def test_scene_expansion_writes_reference_layout_and_point_cloud(models, tmp_path, monkeypatch):
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration
    monkeypatch.chdir(tmp_path)
    ds = "clevr-infinite"
    rng = np.random.default_rng(1)
    pipe = InfiniteSceneGeneration(models(ds), ds, seed_frame=seed_frame(rng, 8.0, 15.0), output_dim=(2, 3))
    pipe.scene_expansion()
    base = tmp_path / "grid_res" / "clevr-infinite_seed0"
    assert len(list(base.glob("im_0000[1-5]_*.png"))) == 5 and len(list(base.glob("R_*_*_*.npy"))) == 5
    assert all(n["visited"] for row in pipe.transform_grid for n in row)
    with open(base / "merged_pcds.ply", "rb") as f:
        header = f.read(300).decode("latin1")
    assert "element vertex %d" % (5 * 256 * 256) in header and "property double x" in header
    xyz, col = pipe.unproject_to_color_point_cloud()
    assert xyz.shape == (5 * 65536, 3) and np.isfinite(xyz).all() and col.min() >= 0 and col.max() <= 1
    # later steps really used several sources (zig-zag neighbours within radius 1)
    pipe.curr = 5
    srcs, _ = pipe.get_src_grid_coords((1, 2))
    assert len(srcs) >= 3


def test_device_point_cloud_equals_the_disk_round_trip(models, tmp_path, monkeypatch):
    """unproject_frames_on_device (resident frame store -> sgam_unproject_points) reproduces, bit for bit, the
    reference-style unproject_to_color_point_cloud over the files it wrote (inference_pipeline.py:1038-1062)."""
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration
    monkeypatch.chdir(tmp_path)
    ds = "google_earth"
    rng = np.random.default_rng(4)
    pipe = InfiniteSceneGeneration(models(ds), ds, seed_frame=seed_frame(rng, 1.4, 3.8), output_dim=(3, 1))
    pipe.scene_expansion()
    xyz_disk, col_disk = pipe.unproject_to_color_point_cloud()           # generated frames only (the seed has no R_/t_ file)
    xyz_dev, col_dev = pipe.unproject_frames_on_device()                  # seed frame first, then the generated ones
    assert xyz_dev.shape == (3 * 65536, 3)
    assert np.array_equal(xyz_dev[65536:].cpu().numpy(), xyz_disk)
    assert np.array_equal(col_dev[65536:].cpu().numpy(), col_disk)


def test_rgbd_integration_branch_with_supplied_depth(models, tmp_path, monkeypatch):
    """configs[2]-shaped loop: use_rgbd_integration=True with the target depth supplied (Open3D stand-in)."""
    from oracle import native
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration, forward_splat_depth
    monkeypatch.chdir(tmp_path)
    ds = "google_earth"
    model = models(ds)
    rng = np.random.default_rng(2)
    captured = {}

    def depth_fn(pipe, src_nodes, T_tgt):
        d = forward_splat_depth(pipe, src_nodes, T_tgt)
        captured["depth"] = d.clone()
        return d

    pipe = InfiniteSceneGeneration(model, ds, seed_frame=seed_frame(rng, 1.4, 3.8), output_dim=(3, 1),
                                   use_rgbd_integration=True, tsdf_depth_fn=depth_fn)
    assert model.use_rgbd_integration is True
    tgt = pipe.next_pose(1)
    tgt_meta = pipe.transform_grid[tgt[0]][tgt[1]]
    src_meta = [pipe.transform_grid[0][0]]
    batch = pipe.prepare_batch_data(tgt_meta, src_meta, pipe.num_src)
    assert "warped_tgt_features" in batch and tuple(batch["warped_tgt_features"].shape) == (1, 3, 256, 256)
    # inverse warp parity against the oracle on the very tensors the pipeline used
    T_tgt = np.eye(4); T_tgt[:3, :3], T_tgt[:3, 3] = tgt_meta["R"], tgt_meta["t"]
    T_src = np.eye(4); T_src[:3, :3], T_src[:3, 3] = src_meta[0]["R"], src_meta[0]["t"]
    T_t2s = torch.from_numpy(np.linalg.inv(T_tgt @ np.linalg.inv(T_src)).astype(np.float32))[None]
    Ks = batch["Ks"]
    proj = (Ks.view(-1, 3, 3) @ T_t2s[:, :3]).numpy()
    src_rgb = batch["src_imgs"].permute(0, 1, 4, 2, 3).contiguous().cpu().numpy()
    o_out, _ = native.inverse_warp(src_rgb, pipe._seed_depth_single[None, None].cpu().numpy(), captured["depth"][None].cpu().numpy(),
                                   Ks[:, 0].inverse().numpy(), proj)
    assert np.array_equal(batch["warped_tgt_features"].cpu().numpy(), o_out)
    res = pipe.one_step_prediction(tgt)
    x = res["x"].cpu().numpy()
    mask = (captured["depth"].cpu().numpy() <= 0)
    assert np.array_equal(x[0, 3], native.depth_code(captured["depth"].cpu().numpy(), mask.astype(np.uint8), ds))
    assert np.array_equal(x[0, :3], o_out[0])
    pipe.curr += 1
    pipe.one_step_prediction(pipe.next_pose(pipe.curr))                           # second step: two sources available
    assert len(pipe._frames) == 3


def test_vqmodel_forward_graph_replay_equals_eager(models):
    """forward() replays a cached CUDA graph per input shape; results must equal the op-by-op path bit for bit and
    must not alias between calls."""
    model = models("google_earth")
    g = torch.Generator().manual_seed(0)
    xs = [torch.rand(1, 4, 64, 64, generator=g).cuda() * 2 - 1 for _ in range(3)]
    ms = [(torch.rand(1, 1, 64, 64, generator=g) < 0.3).cuda() for _ in range(3)]
    model.use_cuda_graph = False
    eager = [model(x, topk=1, extrapolation_mask=m, get_pre_quantized_feature=True, get_quantized_feature=True) for x, m in zip(xs, ms)]
    model.use_cuda_graph = True
    model._graphs = {}
    graphed = [model(x, topk=1, extrapolation_mask=m, get_pre_quantized_feature=True, get_quantized_feature=True) for x, m in zip(xs, ms)]
    assert len(model._graphs) == 1
    for e, r in zip(eager, graphed):
        assert torch.equal(e[0][0], r[0][0]) and torch.equal(e[2], r[2]) and torch.equal(e[3], r[3])
        assert tuple(r[0][0].shape) == (1, 1, 4, 64, 64) and tuple(r[3].shape) == (1, 1, 256, 4, 4)
    dec_plain, loss = model(xs[0], extrapolation_mask=ms[0])               # topk=None signature (config 1): [dec, emb_loss]
    assert torch.equal(dec_plain, eager[0][0][0][0]) and loss.dim() == 0


def test_get_x_returns_the_coded_target_like_the_reference(models):
    """model.py:238: get_x always returns x_dst = cat(dst_img, coded dst_depth); only the scene loop's all-zero placeholders
    (marked `_dst_placeholder` by prepare_batch_data) skip that work."""
    from sgam_neurips22_b200 import synthetic
    ds = "clevr-infinite"
    model = models(ds)
    b = synthetic.scene_step_batch(ds, res=64, batch=2, seed=1)
    rng = np.random.default_rng(2)
    b["dst_img"] = rng.uniform(-1, 1, (2, 64, 64, 3)).astype(np.float32)
    b["dst_depth"] = rng.uniform(7, 16, (2, 64, 64)).astype(np.float32)
    batch = {k: torch.from_numpy(v) for k, v in b.items()}
    x, x_dst = model.get_x(dict(batch), ds)
    x2, x_dst2, mask, wd = model.get_x(dict(batch), ds, return_extrapolation_mask=True, no_depth_range=True)
    assert torch.equal(x, x2) and torch.equal(x_dst, x_dst2) and tuple(x_dst.shape) == (2, 4, 64, 64)
    inv = (1 / torch.from_numpy(b["dst_depth"]) - 1 / 16) / (1 / 7 - 1 / 16)
    ref = torch.cat([torch.from_numpy(b["dst_img"]).permute(0, 3, 1, 2), (2 * inv - 1)[:, None]], 1)
    assert torch.allclose(x_dst.cpu(), ref, atol=1e-6)
    batch["_dst_placeholder"] = True
    assert model.get_x(dict(batch), ds, return_extrapolation_mask=True, no_depth_range=True)[1] is None
