"""CPU-only tests of the host-side logic: drop-in import surface, checkpoint layout, config shim, pose grid /
source selection against the reference's constants, synthetic inputs, and the world_size-2 gloo path of the
final-map all-gather."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_drop_in_import_surface():
    """main_scene_generation.py:6-9 imports exactly these names."""
    ns = {}
    exec("from data.utils.utils import *", ns)
    assert "torch" in ns and "OmegaConf" in ns and "instantiate_from_config" in ns
    from sgam.inference_pipeline import InfiniteSceneGeneration  # noqa: F401
    from sgam.generative_sensing_module.model import VQModel  # noqa: F401
    from sgam.point_rendering.warp import render_projection_from_srcs_fast, median_blur  # noqa: F401
    from sgam.generative_sensing_module.modules.vqvae.quantize import VectorQuantizer2  # noqa: F401


def test_config_shim_matches_entry_point_usage(tmp_path):
    from sgam_neurips22_b200.config import _OmegaConfShim as OC
    p = tmp_path / "config.yaml"
    p.write_text("model:\n  params:\n    n_embed: 64\n    online_kmeans_config:\n      do_online_kmeans_clustering: false\n"
                 "data:\n  params:\n    dataset: clevr-infinite\n")
    cfg = OC.load(str(p))
    cfg.model.params.data_config = cfg.data.params                     # main_scene_generation.py:24
    kw = dict(**cfg.model["params"])                                    # :25
    assert kw["data_config"]["dataset"] == "clevr-infinite" and kw["n_embed"] == 64
    assert kw["online_kmeans_config"]["kmean_init_codebook_path"] is None   # model.py:60 relies on this


def test_vqmodel_checkpoint_layout(state_dicts, tmp_path):
    from oracle import recipes
    from sgam_neurips22_b200 import synthetic
    from sgam_neurips22_b200.model import VQModel
    kw = synthetic.model_kwargs("google_earth")
    m = VQModel(**kw)
    ref_shapes = recipes.hot_path_param_shapes(4096)
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(ref_shapes.keys())
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in ref_shapes.items())
    # Lightning-style checkpoint with training-only keys: filtered by prefix, loaded non-strictly (model.py:87-104)
    full = dict(state_dicts("google_earth"))
    full["loss.discriminator.main.0.weight"] = torch.zeros(3)
    full["loss.perceptual_loss.lin0.model.1.weight"] = torch.zeros(3)
    full["perceptual_loss.net.slice1.0.weight"] = torch.zeros(3)
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": full}, path)
    m2 = VQModel(**kw, ckpt_path=str(path))
    for k, v in state_dicts("google_earth").items():
        assert torch.equal(m2.state_dict()[k], v), k
    with pytest.raises(RuntimeError, match="CUDA"):
        m2.engine                                                       # no CPU fallback


def test_pose_grid_and_source_selection_match_reference_constants(tmp_path, monkeypatch):
    from oracle import recipes
    from sgam_neurips22_b200 import synthetic
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration
    from sgam_neurips22_b200.model import VQModel
    monkeypatch.chdir(tmp_path)
    m = VQModel(**synthetic.model_kwargs("clevr-infinite"))
    m.__dict__["_fake_device"] = True
    rgb = np.zeros((256, 256, 3), np.uint8)
    depth = np.full((256, 256), 10.0, np.float32)
    pipe = object.__new__(InfiniteSceneGeneration)
    # run the constructor logic without a GPU: frame store on CPU
    InfiniteSceneGeneration.__init__.__wrapped__ if hasattr(InfiniteSceneGeneration.__init__, "__wrapped__") else None
    type(m).device = property(lambda self: torch.device("cpu"))
    try:
        pipe = InfiniteSceneGeneration(m, "clevr-infinite", seed_frame=(rgb, depth), output_dim=(4, 5))
    finally:
        type(m).device = property(lambda self: next(self.parameters()).device)
    assert pipe.num_src == 5 and pipe.output_dim == (4, 5)
    for (i, j) in [(0, 0), (2, 3), (3, 4)]:
        R, t = recipes.grid_pose("clevr-infinite", i, j)
        node = pipe.transform_grid[i][j]
        assert np.allclose(node["R"], R) and np.allclose(node["t"], t)
    order = pipe._ordered_grid_coords
    assert order[:6] == [(0, 0), (0, 1), (1, 0), (2, 0), (1, 1), (0, 2)] and len(order) == 20
    # after visiting the first three poses, (2,0) selects the visited poses within radius 1, nearest first
    for c in order[1:3]:
        pipe.transform_grid[c[0]][c[1]]["visited"] = True
    pipe.curr = 3
    srcs, _ = pipe.get_src_grid_coords((2, 0))
    assert srcs == [(1, 0), (0, 0), (0, 1)]                            # distances 0.408, 0.816, 0.913 <= 1
    assert (tmp_path / "grid_res" / "clevr-infinite_seed0" / "im_00000_00_00.png").exists()


def test_synthetic_inputs_are_deterministic_and_well_formed():
    from sgam_neurips22_b200 import synthetic
    a = synthetic.scene_step_batch("clevr-infinite", res=32, batch=2, seed=3)
    b = synthetic.scene_step_batch("clevr-infinite", res=32, batch=2, seed=3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["src_imgs"].shape == (2, 5, 32, 32, 3) and a["src_depths"].shape == (2, 5, 32, 32)
    assert a["src_depths"].min() >= 7 and a["src_depths"].max() <= 16
    u8 = (a["src_imgs"].astype(np.float64) + 1) * 127.5
    assert np.abs(u8 - np.round(u8)).max() < 1e-4                        # on the uint8 lattice like PNG-loaded frames


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from sgam_neurips22_b200 import dist as sdist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    F, H, W = 3, 8, 6
    g = torch.Generator().manual_seed(100 + rank)
    rgb = torch.randint(0, 256, (F, H, W, 3), generator=g, dtype=torch.uint8)
    depth = torch.rand(F, H, W, generator=g)
    poses = torch.rand(F, 12, generator=g, dtype=torch.float64)
    all_rgb, all_depth, all_poses = sdist.gather_scene_map(rgb, depth, poses)
    ok = all_rgb.shape[0] == world * F
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        e_rgb = torch.randint(0, 256, (F, H, W, 3), generator=g, dtype=torch.uint8)
        e_depth = torch.rand(F, H, W, generator=g)
        e_poses = torch.rand(F, 12, generator=g, dtype=torch.float64)
        ok = ok and torch.equal(all_rgb[r * F:(r + 1) * F], e_rgb) and torch.equal(all_depth[r * F:(r + 1) * F], e_depth) \
            and torch.equal(all_poses[r * F:(r + 1) * F], e_poses)
    ok = ok and sdist.shard(7, rank, world) == [t for t in range(7) if t % world == rank]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_final_map_allgather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _records(r, F, H=8, W=6):
    g = torch.Generator().manual_seed(100 + r)
    return (torch.randint(0, 256, (F, H, W, 3), generator=g, dtype=torch.uint8), torch.rand(F, H, W, generator=g),
            torch.rand(F, 12, generator=g, dtype=torch.float64))


def _gloo_uneven_worker(rank, world, port, q):
    """shard() hands out unequal frame counts when trajectories % world != 0 and none at all to surplus ranks."""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from sgam_neurips22_b200 import dist as sdist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    counts = [3, 0, 2][:world]
    all_rgb, all_depth, all_poses = sdist.gather_scene_map(*_records(rank, counts[rank]))
    exp = [_records(r, counts[r]) for r in range(world)]
    ok = torch.equal(all_rgb, torch.cat([e[0] for e in exp])) and torch.equal(all_depth, torch.cat([e[1] for e in exp])) \
        and torch.equal(all_poses, torch.cat([e[2] for e in exp])) and all_rgb.shape[0] == sum(counts)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_final_map_allgather_uneven_counts_and_an_empty_rank_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_uneven_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True), (2, True)]


def test_plain_config_conversion_does_not_duck_type_omegaconf():
    """under omegaconf 2.0 `hasattr(cfg, "to_container")` is True and the attribute is None."""
    from sgam_neurips22_b200.config import ConfigNode, _wrap
    from sgam_neurips22_b200.model import _plain

    class FakeDictConfig(dict):                       # answers None for every missing attribute, like DictConfig 2.0
        def __getattr__(self, k):
            return None
    assert _plain(FakeDictConfig(ch=128, ch_mult=[1, 2])) == {"ch": 128, "ch_mult": [1, 2]}
    node = _wrap({"a": {"b": [1, {"c": 2}]}})
    assert isinstance(node, ConfigNode) and _plain(node) == {"a": {"b": [1, {"c": 2}]}}


def test_file_names_are_derived_from_the_name_not_the_directory(tmp_path):
    """str.replace over the whole path rewrote directories containing 'R', 'dm', 'im' or 'npy'."""
    from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration as ISG
    d = tmp_path / "RESULTS_dm_im.npy_dir"
    assert ISG._sibling(d / "R_00003_01_02.npy", "t") == d / "t_00003_01_02.npy"
    assert ISG._sibling(d / "R_00003_01_02.npy", "im", ".png") == d / "im_00003_01_02.png"
    assert ISG._sibling(d / "dm_00000_00_00.npy", "im", ".png") == d / "im_00000_00_00.png"


def test_unproject_records_matches_prepare_pcd():
    from oracle import native
    from sgam_neurips22_b200 import dist as sdist
    rng = np.random.default_rng(0)
    depth = rng.uniform(1, 5, (2, 6, 7)).astype(np.float32)
    rgb = rng.integers(0, 256, (2, 6, 7, 3)).astype(np.uint8)
    K = np.array([[50., 0, 3.5], [0, 50., 3.0], [0, 0, 1]])
    poses = []
    for f in range(2):
        A = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        poses.append(np.concatenate([A.reshape(-1), rng.standard_normal(3)]))
    poses = np.stack(poses)
    xyz, col = sdist.unproject_records(torch.from_numpy(rgb), torch.from_numpy(depth), torch.from_numpy(poses), K)
    for f in range(2):
        Rt = np.eye(4)
        Rt[:3, :3], Rt[:3, 3] = poses[f, :9].reshape(3, 3), poses[f, 9:]
        ref = native.unproject_world(depth[f], np.linalg.inv(K), np.linalg.inv(Rt))
        assert np.allclose(xyz[f * 42:(f + 1) * 42].numpy(), ref, atol=1e-9)
    assert np.allclose(col.numpy(), rgb.reshape(-1, 3) / 255.0)


def test_frustum_box_contains_every_point_the_cameras_can_see():
    from sgam_neurips22_b200.tsdf import frustum_box
    rng = np.random.default_rng(3)
    K = np.array([[248.88887, 0, 128], [0, 248.88887, 128], [0, 0, 1.0]])
    poses = []
    for i in range(5):
        c, s = np.cos(0.5), np.sin(0.5)
        c2w = np.eye(4)
        c2w[:3, :3] = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        c2w[:3, 3] = [-3.0, -6.0 + 0.06 * i, 2.0]
        poses.append(np.linalg.inv(c2w @ np.diag([1., -1., -1., 1.])))
    lo, hi = frustum_box(K, poses, 256, 256, z_far=5.0, pad=0.1)
    for T in poses:
        uv = rng.uniform(0, 256, (200, 2))
        z = rng.uniform(0.05, 5.0, 200)
        cam = np.stack([(uv[:, 0] - 128) / K[0, 0] * z, (uv[:, 1] - 128) / K[1, 1] * z, z, np.ones(200)])
        world = (np.linalg.inv(T) @ cam)[:3]
        assert (world.min(1) >= lo).all() and (world.max(1) <= hi).all()


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses ONE JSON line from `bench.py --impl reference`; it runs the oracle port on the host cores only."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
