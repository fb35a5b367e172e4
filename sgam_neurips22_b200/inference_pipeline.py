"""Mirror of the reference's scene loop (sgam/inference_pipeline.py:21-1062) driving the B200 kernels.

Same class name, constructor keywords, public methods (`scene_expansion`, `one_step_prediction`, `next_pose`,
`get_src_grid_coords`, `prepare_batch_data`, `inverse_warping`, `save_to_disk`, `unproject_to_color_point_cloud`) and
on-disk layout (`grid_res/<data>_seed<k>/{im,dm,R,t}_<idx:05d>_<i:02d>_<j:02d>.*`), so main_scene_generation.py runs
unchanged.  What is re-designed:

  * generated frames stay resident on the GPU in a frame store (uint8-lattice fp32 colours + fp32 depth: bit-identical
    to what the reference re-reads from its PNG / NPY files, inference_pipeline.py:534-536), so a step does no disk
    read, no PIL decode and no large host->device copy; files are still written when `save_res_to_disk` is set;
  * the per-source Python z-test loop of `inverse_warping` (:725-737) is one kernel;
  * no matplotlib (`plt.show()` per frame, :903-904), no tqdm prints in the step.

use_rgbd_integration=True (:745-838; the README default) runs on a device-resident TSDF volume (tsdf.py / csrc/tsdf.cu):
the selected source frames are integrated with Open3D 0.15.2's ScalableTSDFVolume arithmetic and the target depth is
ray-cast from the volume -- no Open3D, no mesh extraction, no off-screen renderer, no host round trip.  `tsdf_depth_fn`
lets the caller substitute any other integrated-depth source (e.g. the real Open3D where it is installed).
"""
import os
import shutil
from pathlib import Path

import numpy as np
import torch

from . import ops
from .model import VQModel
from .tsdf import TSDFVolume, frustum_box

# (voxel_length, sdf_trunc) of the reference's ScalableTSDFVolume (:119-131) and the far end of each dataset's depth
# code (model.py:211-226), which bounds the dense grid
TSDF_PARAMS = {"clevr-infinite": (0.05, 0.5), "google_earth": (0.01, 0.03)}
DEPTH_FAR = {"clevr-infinite": 16.0, "google_earth": 14.765625 - 10.0}


def _ray_to_z(depth, K):
    """CLEVR depth is stored along the ray; convert to z (inference_pipeline.py:71-79), float64 like numpy."""
    h, w = depth.shape[:2]
    xs, ys = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    return depth * K[0][0] / np.sqrt(K[0][0] ** 2 + (K[0][2] - ys - 0.5) ** 2 + (K[1][2] - xs - 0.5) ** 2)


def write_ply(path, xyz, rgb):
    """Binary little-endian PLY with double xyz + uchar rgb (what o3d.io.write_point_cloud emits, :441-444)."""
    xyz = np.asarray(xyz, np.float64)
    rgb = np.clip(np.asarray(rgb) * 255.0 + 0.5, 0, 255).astype(np.uint8) if np.asarray(rgb).dtype != np.uint8 else rgb
    rec = np.empty(len(xyz), dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["r"], rec["g"], rec["b"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    with open(path, "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(xyz)}\nproperty double x\n"
                 "property double y\nproperty double z\nproperty uchar red\nproperty uchar green\n"
                 "property uchar blue\nend_header\n").encode())
        rec.tofile(f)


def forward_splat_depth(pipe, src_nodes, T_tgt):
    """An Open3D-free `tsdf_depth_fn`: the target depth is the z-buffered (nearest-depth) forward splat of the
    selected source frames with 3x3 hole fill -- the same fused kernel as stage (i) with SGAM_SPLAT_ZMIN.  It stands
    in for ScalableTSDFVolume + mesh rendering (inference_pipeline.py:745-838) where Open3D is unavailable; it is NOT
    numerically equivalent to TSDF fusion (that depth is parity-unpinned)."""
    dev = pipe.device
    frames = [pipe._frames[tuple(n["grid_coord"])] for n in src_nodes]
    rgb = torch.stack([f[0] for f in frames])[None].contiguous()
    dm = torch.stack([f[1] for f in frames])[None].contiguous()
    N = len(src_nodes)
    K = torch.from_numpy(pipe.K.astype(np.float32))
    T = torch.eye(4).repeat(1, N, 1, 1)
    for i, n in enumerate(src_nodes):
        T_src = np.eye(4)
        T_src[:3, :3], T_src[:3, 3] = n["R"], n["t"]
        T[0, i] = torch.from_numpy((T_tgt @ np.linalg.inv(T_src)).astype(np.float32))
    Kinv = K.inverse()[None, None].repeat(1, N, 1, 1).contiguous()
    out = ops.splat_forward(rgb, dm, ops.h2d(K[None], dev), ops.h2d(Kinv, dev), ops.h2d(T, dev), pipe.data, channels_last=True,
                            policy=ops.SPLAT_ZMIN, want_merge_depth=True)
    return out["merge_depth"][0, 0]


class InfiniteSceneGeneration:

    def __init__(self,
                 dynamic_model, data, topk=1, step_size_denom=2, use_rgbd_integration=False, use_discriminator_loss=False,
                 discriminator_loss_weight=0, recon_on_visible=False, offscreen_rendering=True, output_dim=None, seed_index=0,
                 num_src=None, tsdf_depth_fn=None, template_root="templates", output_root="grid_res", seed_frame=None,
                 image_resolution=(256, 256), integrate_once=False, tsdf_max_bytes=2 << 30):
        self.use_discriminator_loss = use_discriminator_loss
        self.offscreen_rendering = offscreen_rendering
        self.discriminator_loss_weight = discriminator_loss_weight
        self.seed_index = seed_index
        self.topk = topk
        self.recon_on_visible = recon_on_visible
        self.use_rgbd_integration = use_rgbd_integration
        self.step_size_denom = step_size_denom
        self.dynamic_model = dynamic_model
        self.data = data
        self.tsdf_depth_fn = tsdf_depth_fn
        # The reference re-integrates EVERY selected source frame at EVERY step (:771-777), so a frame that stays within
        # the source radius for k steps enters the volume k times (its observations weigh k).  integrate_once=True is a
        # flagged deviation: each frame is fused the first time it is selected and never again (one integration per
        # step in steady state instead of num_src, every observation weighs 1).  Default: the reference's behaviour.
        self.integrate_once = bool(integrate_once)
        self.tsdf_max_bytes = int(tsdf_max_bytes)
        self._integrated = set()
        if data not in ("clevr-infinite", "google_earth"):
            raise NotImplementedError                                      # inference_pipeline.py:55-56
        # :42,47 hard-code 256x256; BASELINE.json configs[4] runs GoogleEarth at 512x512, so it is a keyword here
        self.image_resolution = tuple(int(v) for v in image_resolution)
        if self.image_resolution[0] % 16 or self.image_resolution[1] % 16:
            raise ValueError("image_resolution must be a multiple of 16 (the VQGAN down-samples by 16)")
        self.output_dim = self.default_output_dim(data) if output_dim is None else output_dim
        is_vq = isinstance(dynamic_model, VQModel)
        self.K = self.intrinsics_for(data, self.image_resolution)
        if data == "clevr-infinite":
            self.num_src = (5 if num_src is None else num_src) if is_vq else 1                 # :68
        else:
            self.num_src = (3 if num_src is None else num_src) if is_vq else 1                 # :90
        self.K_inv = np.linalg.inv(self.K)
        self.curr = 1
        self.trajectory_shape = "grid"
        self.global_pcd = []
        self.total_inconsistency = 0

        name = data + "_seed" + str(seed_index)
        self.grid_transform_path = Path(f"{output_root}/{name}")
        self.grid_transform_mask_path = self.grid_transform_path / "masks"
        seed_rgb, seed_depth = self._stage_seed(template_root, seed_frame)
        os.makedirs(self.grid_transform_mask_path, exist_ok=True)

        self.prepare_grid(self.output_dim, self.get_known_map(), self.grid_transform_path)
        self._ordered_grid_coords = self.zig_zag_order()
        self.dynamic_model.use_rgbd_integration = self.use_rgbd_integration    # :117

        # device-resident frame store: grid coord -> (rgb [H,W,3] fp32 on the uint8 lattice, depth [H,W] fp32)
        self.device = dynamic_model.device if hasattr(dynamic_model, "device") else torch.device("cuda:0")
        self._frames = {}
        first = self._ordered_grid_coords[0]
        self._store_seed(first, seed_rgb, seed_depth)
        self._K_host = torch.from_numpy(self.K.astype(np.float32))

        self.volume = None
        if self.use_rgbd_integration and tsdf_depth_fn is None:
            self._init_volume()

    @staticmethod
    def default_output_dim(data):
        return (20, 20) if data == "clevr-infinite" else (100, 1)

    @staticmethod
    def intrinsics_for(data, image_resolution=(256, 256)):
        """:61-65 (CLEVR, given for 256x256) and :83-89 (GoogleEarth, given for 512x512), scaled to the working resolution."""
        if data == "clevr-infinite":
            K, base = np.array([[355.5555, 0, 128], [0, 355.5555, 128], [0, 0, 1]]), 256
        else:
            K, base = np.array([[497.77774, 0, 256], [0, 497.77774, 256], [0, 0, 1]]), 512
        K[0] = K[0] * image_resolution[1] / base
        K[1] = K[1] * image_resolution[0] / base
        return K

    # ------------------------------------------------------------------------------------ seed / files
    def _stage_seed(self, template_root, seed_frame):
        """Recreate the output directory from the templates (:37-54) and return the seed frame (uint8 RGB at the
        working resolution, fp32 depth)."""
        out = self.grid_transform_path
        if os.path.exists(out):
            shutil.rmtree(out)
        if seed_frame is not None:                                          # synthetic seed (benchmarks, tests)
            os.makedirs(out, exist_ok=True)
            rgb, depth = seed_frame
            np.save(out / "dm_00000_00_00.npy", np.asarray(depth, np.float32))
            self._save_png(out / "im_00000_00_00.png", np.asarray(rgb, np.uint8))
        elif self.data == "clevr-infinite":
            shutil.copytree(f"{template_root}/clevr-infinite", out)
            for dm_path in sorted(out.glob("dm*")):                         # :71-79 in-place ray -> z conversion
                np.save(dm_path, _ray_to_z(np.load(dm_path), self.K))
        else:
            os.makedirs(out, exist_ok=True)
            img_fn = sorted(Path(f"{template_root}/google_earth/seed{self.seed_index}").glob("im*"))[0]
            stem = img_fn.name[len("im"):-len(".png")]                       # "_00000"
            shutil.copy(img_fn, out / f"im{stem}_00_00.png")
            shutil.copy(img_fn.with_name(f"dm{stem}.npy"), out / f"dm{stem}_00_00.npy")
        dm_file = sorted(out.glob("dm_*_00_00.npy"))[0]
        return self._load_rgb(self._sibling(dm_file, "im", ".png")), np.load(dm_file)

    @staticmethod
    def _sibling(path, prefix, suffix=None):
        """`<dir>/<old prefix>_<rest>.<ext>` -> `<dir>/<prefix>_<rest><suffix>`: the reference derives its file names by
        str.replace on a relative path (e.g. :191); here roots are configurable, so only the NAME is rewritten."""
        path = Path(path)
        rest = path.name[path.name.index("_"):]
        if suffix is not None:
            rest = rest[:rest.rindex(".")] + suffix
        return path.with_name(prefix + rest)

    @staticmethod
    def _save_png(path, rgb):
        from PIL import Image
        Image.fromarray(rgb).save(str(path), format="png")

    def _load_rgb(self, path):
        from PIL import Image
        im = Image.open(path).convert("RGB")
        return np.array(im.resize((self.image_resolution[1], self.image_resolution[0]), resample=Image.LANCZOS))   # :534

    def _store_seed(self, coord, rgb_u8, depth):
        H, W = self.image_resolution
        d = torch.from_numpy(np.asarray(depth))[None, None]
        d = torch.nn.functional.interpolate(d, size=(H, W))[0, 0].numpy()                  # :536 nearest resize
        self._seed_depth_single = torch.from_numpy(d.astype(np.float32)).to(self.device)   # what the TSDF / inverse-warp path sees
        if self.data == "clevr-infinite":
            d = _ray_to_z(d, self.K)                                                       # :582-590: seed frame converted again
        rgb = (np.asarray(rgb_u8, np.float64) / 127.5 - 1.0).astype(np.float32)
        self._frames[tuple(coord)] = (torch.from_numpy(rgb).to(self.device), torch.from_numpy(d.astype(np.float32)).to(self.device))

    def get_known_map(self):
        known = {}
        for f in Path(self.grid_transform_path).glob("dm*"):
            parts = f.name[3:-4].split("_")
            known[(int(parts[1]), int(parts[2]))] = {"rgb_path": str(self._sibling(f, "im", ".png")),
                                                     "depth_path": str(f), "orig_frame_idx": int(parts[0])}
        return known

    # ------------------------------------------------------------------------------------ pose grid
    def prepare_grid(self, grid_size, known_map, output_folder):
        """Camera poses of the (rows x cols) grid (:157-204)."""
        if self.data == "google_earth":
            start = np.array([[1., 0., 0., -3.], [0., 0.86602527, -0.50000024, -6.],
                              [0., 0.50000024, 0.86602527, 2.], [0., 0., 0., 1.]])
            step_i = np.array([0., 0.11878788, 0.]) / self.step_size_denom
            step_j = np.array([0.12, 0, 0.]) / self.step_size_denom
        else:
            start = np.array([[1., 0., 0., -20.], [0., 0.95533651, -0.29552022, -20.],
                              [0., 0.29552022, 0.95533651, 0.], [0., 0., 0., 1.]])
            step_j = np.array([0.81632614, 0, 0.]) / self.step_size_denom
            step_i = np.array([0, 0.81632614, 0.]) / self.step_size_denom
        flip = np.diag([1., -1., -1., 1.])
        self.transform_grid, self.anchor_poses = [], {}
        for i in range(grid_size[0]):
            row = []
            for j in range(grid_size[1]):
                c2w = np.eye(4)
                c2w[:3, :3] = start[:3, :3]
                c2w[:3, 3] = start[:3, 3] + step_j * j + step_i * i
                w2c = np.linalg.inv(c2w @ flip)
                R, t = w2c[:3, :3], w2c[:3, 3]
                k = known_map.get((i, j))
                node = {"R": R, "t": t, "K": self.K, "position": -R.T @ t,
                        "rgb_path": k["rgb_path"] if k else f"{output_folder}/im_{i * grid_size[1] + j:05d}.png",
                        "depth_path": k["depth_path"] if k else f"{output_folder}/dm_{i * grid_size[1] + j:05d}.npy",
                        "R_path": f"{output_folder}/R_{i:05d}.npy", "K_path": f"{output_folder}/K_{i:05d}.npy",
                        "t_path": f"{output_folder}/t_{i:05d}.npy", "visited": k is not None, "grid_coord": (i, j)}
                if k:
                    self.anchor_poses[(i, j)] = node
                row.append(node)
            self.transform_grid.append(row)

    def zig_zag_order(self):
        """Anti-diagonal zig-zag over the grid (:452-475)."""
        rows, cols = self.output_dim
        diags = [[] for _ in range(rows + cols - 1)]
        for i in range(rows):
            for j in range(cols):
                if (i + j) % 2 == 0:
                    diags[i + j].insert(0, (i, j))
                else:
                    diags[i + j].append((i, j))
        order = [c for d in diags for c in d]
        self.transform_grid[order[0][0]][order[0][1]]["visited"] = True
        return order

    def next_pose(self, curr):
        return self._ordered_grid_coords[curr]

    def get_src_grid_coords(self, tgt_grid_coord):
        """Visited poses within the radius, nearest first, at most num_src (:507-531)."""
        tgt = self.transform_grid[tgt_grid_coord[0]][tgt_grid_coord[1]]
        radius = 1 if self.data == "clevr-infinite" else 0.3
        cands = []
        for c in self._ordered_grid_coords[:self.curr]:
            node = self.transform_grid[c[0]][c[1]]
            dist = np.linalg.norm(node["position"] - tgt["position"])
            if node["visited"] and dist <= radius:
                cands.append((c, dist))
        cands.sort(key=lambda x: x[1])                                     # stable, like sorted() in the reference
        return [c for c, _ in cands[:self.num_src]], None

    # ------------------------------------------------------------------------------------ batch
    def prepare_batch_data(self, tgt_node, src_nodes, num_src):
        """:533-609 from the device frame store: returns the reference's batch dict (device tensors for the images,
        host tensors for the 3x3 / 4x4 matrices so that no device->host sync is needed downstream)."""
        H, W = self.image_resolution
        frames = [self._frames[tuple(n["grid_coord"])] for n in src_nodes]
        src_imgs = torch.stack([f[0] for f in frames])[None]                # [1,N,H,W,3]
        src_depths = torch.stack([f[1] for f in frames])[None]              # [1,N,H,W]
        T_tgt = np.eye(4)
        T_tgt[:3, :3], T_tgt[:3, 3] = tgt_node["R"], tgt_node["t"]
        R_rels, t_rels, T_tgt2srcs = [], [], []
        for n in src_nodes:
            T_src = np.eye(4)
            T_src[:3, :3], T_src[:3, 3] = n["R"], n["t"]
            T_rel = T_tgt @ np.linalg.inv(T_src)                            # :556-569, float64
            T_tgt2srcs.append(np.linalg.inv(T_rel))
            R_rels.append(T_rel[:3, :3])
            t_rels.append(T_rel[:3, 3])
        N = len(src_nodes)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        batch = {"Ks": f32(np.stack([self.K] * N))[None], "K_invs": f32(np.stack([np.linalg.inv(self.K)] * N))[None],
                 "R_rels": f32(np.stack(R_rels))[None], "t_rels": f32(np.stack(t_rels))[None],
                 "dst_img": self._dst_zeros(H, W)[0], "dst_depth": self._dst_zeros(H, W)[1],       # :596-597 placeholders
                 "_dst_placeholder": True,                  # tells VQModel.get_x not to code the (all-zero) target frame
                 "src_imgs": src_imgs, "src_depths": src_depths}
        if self.use_rgbd_integration:
            first = tuple(self._ordered_grid_coords[0])                     # the seed's second ray->z conversion happens
            dm_iw = torch.stack([self._seed_depth_single if tuple(n["grid_coord"]) == first else self._frames[tuple(n["grid_coord"])][1]
                                 for n in src_nodes])[None]                 # after the inverse warp in the reference (:570-590)
            tgt_depth = self.rgbd_integration(src_nodes, T_tgt, src_depths=dm_iw[0])   # [H,W] fp32 on device
            warped = self.inverse_warping(src_imgs.permute(0, 1, 4, 2, 3), dm_iw, tgt_depth[None],
                                          batch["Ks"], f32(self.K)[None], f32(np.stack(T_tgt2srcs))[None], as_numpy=False)
            batch["warped_tgt_features"] = warped[None]
            batch["warped_tgt_depth"] = tgt_depth[None]
        return batch

    def _dst_zeros(self, H, W):
        z = getattr(self, "_dst_zero_cache", None)
        if z is None or z[0].shape[1:3] != (H, W):
            z = self._dst_zero_cache = (torch.zeros(1, H, W, 3), torch.zeros(1, H, W))
        return z

    def inverse_warping(self, src_imgs, src_depths, tgt_depth, src_intrinsics, tgt_intrinsic, T_tgt2srcs,
                        padding_mode='zeros', depth_threshold=100, as_numpy=True):
        """:662-743 as one kernel.  src_imgs [B,N,3,H,W] (or a permuted view of [B,N,H,W,3]), src_depths [B,N,H,W],
        tgt_depth [B,H,W], src_intrinsics [B,N,3,3], tgt_intrinsic [B,3,3], T_tgt2srcs [B,N,4,4]."""
        dev = self.device
        Ks = torch.as_tensor(src_intrinsics).detach().to("cpu", torch.float32)
        T = torch.as_tensor(T_tgt2srcs).detach().to("cpu", torch.float32)
        B, N = Ks.shape[:2]
        proj = (Ks.reshape(-1, 3, 3) @ T.reshape(-1, 4, 4)[:, :3]).reshape(B, N, 3, 4)       # :696 (host, fp32)
        Kinv_tgt = torch.as_tensor(tgt_intrinsic).detach().to("cpu", torch.float32).inverse()    # :693
        src_imgs = torch.as_tensor(src_imgs).to(dev, torch.float32)
        channels_last = not src_imgs.is_contiguous() and src_imgs.permute(0, 1, 3, 4, 2).is_contiguous()
        rgb = src_imgs.permute(0, 1, 3, 4, 2) if channels_last else src_imgs.contiguous()
        out = ops.inverse_warp(rgb, torch.as_tensor(src_depths).to(dev, torch.float32).contiguous(),
                               torch.as_tensor(tgt_depth).to(dev, torch.float32).contiguous(),
                               ops.h2d(Kinv_tgt, dev), ops.h2d(proj, dev), channels_last=channels_last)
        return out[0].cpu().numpy() if as_numpy else out[0]

    def _init_volume(self):
        """:119-131: the TSDF volume, as a dense device grid over the box the trajectory's view frusta can reach."""
        H, W = self.image_resolution
        vox, trunc = TSDF_PARAMS[self.data]
        self._z_far = 1.1 * DEPTH_FAR[self.data]
        poses = []
        for row in self.transform_grid:
            for node in row:
                T = np.eye(4)
                T[:3, :3], T[:3, 3] = node["R"], node["t"]
                poses.append(T)
        lo, hi = frustum_box(self.K, poses, H, W, self._z_far, pad=trunc + 16 * vox)
        self.volume = TSDFVolume(vox, trunc, lo, hi, device=self.device, with_color=True, max_bytes=self.tsdf_max_bytes)

    def rgbd_integration(self, src_nodes, T_tgt, src_depths=None):
        """Integrated target depth [H,W] on the device (:745-838): integrate the selected source frames into the
        volume (again at every step they are selected, like the reference), then render the target view's depth."""
        if self.tsdf_depth_fn is not None:
            d = self.tsdf_depth_fn(self, src_nodes, T_tgt)
            return torch.as_tensor(d).to(self.device, torch.float32).contiguous()
        H, W = self.image_resolution
        for i, n in enumerate(src_nodes):
            key = tuple(n["grid_coord"])
            if self.integrate_once:
                if key in self._integrated:
                    continue
                self._integrated.add(key)
            rgb, depth = self._frames[key]
            if src_depths is not None:
                depth = src_depths[i]
            T_src = np.eye(4)
            T_src[:3, :3], T_src[:3, 3] = n["R"], n["t"]
            self.volume.integrate(depth.contiguous(), rgb, self.K, T_src, depth_trunc=20.0)      # :771-777
        return self.volume.render_depth(self.K, T_tgt, H, W, z_far=self._z_far)                  # :786-827

    # ------------------------------------------------------------------------------------ the step
    @torch.no_grad()
    def one_step_prediction(self, tgt_pose_grid_coord, save_res_to_disk=True):
        """One generated frame (:860-926)."""
        src_coords, _ = self.get_src_grid_coords(tgt_pose_grid_coord)
        tgt_meta = self.transform_grid[tgt_pose_grid_coord[0]][tgt_pose_grid_coord[1]]
        src_metas = [self.transform_grid[c[0]][c[1]] for c in src_coords]
        batch = self.prepare_batch_data(tgt_meta, src_metas, self.num_src)
        batch['src_depths'] = batch['src_depths'][..., None]                                  # :870
        x, x_dst, extrapolation_mask, warped_depth = self.dynamic_model.get_x(
            batch, self.data, return_extrapolation_mask=True, no_depth_range=True, parallel=True)
        x_sample_dets, _, pre_quantized_features, quantized_features = self.dynamic_model(
            x, topk=self.topk, extrapolation_mask=extrapolation_mask, get_pre_quantized_feature=True,
            get_quantized_feature=True, sample_number=1)
        x_sample_dets = x_sample_dets[0]                                                      # sample number is 1
        rgb_u8, depth, src_rgb = ops.frame_outputs(x_sample_dets[0][:1].contiguous(), self.data, want_src_rgb=True)
        self._frames[tuple(tgt_pose_grid_coord)] = (src_rgb[0], depth[0])                     # stays on the device
        if save_res_to_disk:
            self.save_to_disk(tgt_pose_grid_coord, rgb_u8[0].cpu().numpy(), depth[0].cpu().numpy())
        else:
            tgt_meta["visited"] = True
        return {"rgbd": x_sample_dets.squeeze().detach(), "feature": quantized_features.squeeze().detach(),
                "pre_quantized_features": pre_quantized_features.squeeze().detach(), "fixed": False, "x": x.detach(),
                "batch_src_imgs": batch['src_imgs'], "batch_src_depths": batch['src_depths'],
                "batch_R_rels": batch['R_rels'], "batch_t_rels": batch['t_rels'], "warped_depth": warped_depth}

    def save_to_disk(self, tgt_pose_grid_coord, rgb, depth):
        """:928-959."""
        index = self.curr
        node = self.transform_grid[tgt_pose_grid_coord[0]][tgt_pose_grid_coord[1]]
        suffix = f"_{tgt_pose_grid_coord[0]:02d}_{tgt_pose_grid_coord[1]:02d}"
        base = self.grid_transform_path
        np.save(str(base / f"R_{index:05d}{suffix}.npy"), node['R'])
        np.save(str(base / f"t_{index:05d}{suffix}.npy"), node['t'])
        np.save(str(base / f"dm_{index:05d}{suffix}.npy"), depth)
        self._save_png(base / f"im_{index:05d}{suffix}.png", rgb)
        node['visited'] = True
        node["rgb_path"] = str(base / f"im_{index:05d}{suffix}.png")
        node['depth_path'] = str(base / f"dm_{index:05d}{suffix}.npy")
        node['R_path'] = str(base / f"R_{index:05d}{suffix}.npy")
        node['K_path'] = str(base / f"K_{index:05d}{suffix}.npy")
        node['t_path'] = str(base / f"t_{index:05d}{suffix}.npy")

    def scene_expansion(self, return_hs=False):
        """:433-450."""
        for _ in range(self.output_dim[0] * self.output_dim[1] - 1):
            self.one_step_prediction(self.next_pose(self.curr))
            self.curr += 1
        print(f"Successfully unrolling, results saved at {self.grid_transform_path}")
        xyz, rgb = self.unproject_to_color_point_cloud()
        merged = str(self.grid_transform_path / "merged_pcds.ply")
        write_ply(merged, xyz, rgb)
        print(f"Merged per-view point cloud is saved at {merged}")
        if self.use_rgbd_integration and isinstance(self.volume, TSDFVolume):                 # :446-450
            pts, cols = self.volume.extract_point_cloud()
            path = str(self.grid_transform_path / "rgbd_integrated_mesh.ply")
            write_ply(path, pts.cpu().numpy(), cols.cpu().numpy())
            print(f"RGB-D integrated point cloud is saved at {path}")

    # ------------------------------------------------------------------------------------ final map
    def prepare_pcd(self, depth, color, K, Rt):
        """:1014-1036: world-space points (float64) and colours in [0,1] of one frame."""
        h, w = depth.shape
        xs, ys = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
        pix = np.stack([xs.reshape(-1), ys.reshape(-1), np.ones(h * w)])
        cam = (np.linalg.inv(K) @ pix) * depth.reshape(1, -1)
        world = np.linalg.inv(Rt) @ np.concatenate([cam, np.ones((1, h * w))], 0)
        return world[:3].T, color.reshape(h * w, 3) / 255.

    def unproject_frames_on_device(self):
        """:1038-1062 without the disk round trip: every visited frame of the resident store, in zig-zag order, through
        the sgam_unproject_points kernel.  Returns device tensors (xyz [P,3] f64, colours [P,3] f64 in [0,1]).  The
        seed frame carries the depth the splat path uses (for CLEVR: after the second ray->z conversion, :582-590)."""
        coords = [c for c in self._ordered_grid_coords if tuple(c) in self._frames]
        rgb = torch.stack([torch.round((self._frames[tuple(c)][0] + 1.0) * 127.5).to(torch.uint8) for c in coords])
        depth = torch.stack([self._frames[tuple(c)][1] for c in coords]).contiguous()
        Rt = np.tile(np.eye(4), (len(coords), 1, 1))
        for k, c in enumerate(coords):
            node = self.transform_grid[c[0]][c[1]]
            Rt[k, :3, :3], Rt[k, :3, 3] = node["R"], node["t"]
        return ops.unproject_points(depth, rgb.contiguous(), self.K, Rt)

    def unproject_to_color_point_cloud(self):
        """:1038-1062 from the files on disk (sorted by R_* name)."""
        from PIL import Image
        pts, cols = [], []
        for R_path in sorted(self.grid_transform_path.glob("R_*_*_*.npy")):
            Rt = np.eye(4)
            Rt[:3, :3] = np.load(str(R_path))
            Rt[:3, 3] = np.load(str(self._sibling(R_path, "t")))
            depth = np.load(str(self._sibling(R_path, "dm")))
            color = np.array(Image.open(str(self._sibling(R_path, "im", ".png"))).convert("RGB"))
            p, c = self.prepare_pcd(depth, color, self.K, Rt)
            pts.append(p)
            cols.append(c)
        return np.concatenate(pts), np.concatenate(cols)
