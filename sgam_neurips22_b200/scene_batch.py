"""Many independent trajectories through the scene loop at once (BASELINE.json configs[3] and [4]).

The reference generates ONE trajectory per process, one frame at a time (sgam/inference_pipeline.py:433-440; batch 1
is hard-coded in quantize.py:368).  Trajectories with different seed frames share nothing but the read-only weights,
so on a B200 they are advanced in lock-step: every step gathers each trajectory's `prepare_batch_data` dictionary
(:533-609), concatenates them along the batch dimension and runs ONE `get_x` + ONE `VQModel.forward` for the whole
micro-batch -- the same two calls `one_step_prediction` makes (:872-877) -- then hands every trajectory its frame back.
Each trajectory is a full `InfiniteSceneGeneration` (own pose grid, frame store, TSDF volume, output directory), so
everything else of the reference's behaviour is unchanged.

Multi-GPU: trajectory t lives on rank t mod world (`dist.shard`); there is no collective in the loop; `gather_map`
is the single all-gather of the compact per-frame records at the end.  Results do not depend
on the number of ranks as long as `micro_batch` is the same (kernels choose tilings by batch size).
"""
import numpy as np
import torch

from . import dist as sdist
from . import ops
from .inference_pipeline import InfiniteSceneGeneration


class TrajectoryBatch:

    def __init__(self, dynamic_model, data, seed_frames, micro_batch=8, rank=0, world_size=1, output_root="grid_res",
                 **pipeline_kwargs):
        """seed_frames: one (rgb uint8 [H,W,3], depth fp32 [H,W]) per trajectory of the WHOLE job; this rank builds the
        ones `dist.shard` assigns to it.  pipeline_kwargs go to every InfiniteSceneGeneration (output_dim,
        use_rgbd_integration, num_src, ...)."""
        self.model, self.data = dynamic_model, data
        self.micro_batch = int(micro_batch)
        self.ids = sdist.shard(len(seed_frames), rank, world_size)
        self.pipes = [InfiniteSceneGeneration(dynamic_model, data, seed_frame=seed_frames[t], seed_index=t,
                                              output_root=output_root, **pipeline_kwargs) for t in self.ids]
        # a rank may own no trajectory (fewer trajectories than ranks): it still takes part in gather_map
        self.image_resolution = tuple(pipeline_kwargs.get("image_resolution", (256, 256)))
        self.K = InfiniteSceneGeneration.intrinsics_for(data, self.image_resolution)
        dim = pipeline_kwargs.get("output_dim") or InfiniteSceneGeneration.default_output_dim(data)
        self.n_steps = dim[0] * dim[1] - 1

    @torch.no_grad()
    def step(self, save_res_to_disk=False):
        """Advance every local trajectory by one frame (one_step_prediction, :860-926, batched)."""
        for lo in range(0, len(self.pipes), self.micro_batch):
            chunk = self.pipes[lo:lo + self.micro_batch]
            coords, batches = [], []
            for p in chunk:
                tgt = p.next_pose(p.curr)
                src_coords, _ = p.get_src_grid_coords(tgt)
                tgt_meta = p.transform_grid[tgt[0]][tgt[1]]
                batches.append(p.prepare_batch_data(tgt_meta, [p.transform_grid[c[0]][c[1]] for c in src_coords], p.num_src))
                coords.append(tgt)
            n_src = {b["src_depths"].shape[1] for b in batches}
            if len(n_src) != 1:
                raise RuntimeError(f"trajectories of one micro-batch must select the same number of sources, got {sorted(n_src)}")
            batch = {k: (torch.cat([b[k] for b in batches], 0) if torch.is_tensor(batches[0][k]) else batches[0][k]) for k in batches[0]}
            batch["src_depths"] = batch["src_depths"][..., None]                                   # :870
            x, _, mask, _ = self.model.get_x(batch, self.data, return_extrapolation_mask=True, no_depth_range=True, parallel=True)
            decs, _, _, _ = self.model(x, topk=chunk[0].topk, extrapolation_mask=mask, get_pre_quantized_feature=True,
                                       get_quantized_feature=True, sample_number=1)
            rgb_u8, depth, src_rgb = ops.frame_outputs(decs[0][0].contiguous(), self.data, want_src_rgb=True)
            for i, (p, tgt) in enumerate(zip(chunk, coords)):
                p._frames[tuple(tgt)] = (src_rgb[i], depth[i])
                if save_res_to_disk:
                    p.save_to_disk(tgt, rgb_u8[i].cpu().numpy(), depth[i].cpu().numpy())
                else:
                    p.transform_grid[tgt[0]][tgt[1]]["visited"] = True
                p.curr += 1

    def scene_expansion(self, save_res_to_disk=False):
        for _ in range(self.n_steps):
            self.step(save_res_to_disk)

    # ------------------------------------------------------------------------------------------ final map
    def local_records(self):
        """Compact records of every local frame in (trajectory, zig-zag) order: uint8 RGB [F,H,W,3], fp32 depth
        [F,H,W], float64 poses [F,12] (R row-major, t).  The seed frame's depth is the one the splat path uses."""
        rgbs, depths, poses = [], [], []
        if not self.pipes:
            H, W = self.image_resolution
            dev = self.model.device
            return (torch.empty(0, H, W, 3, dtype=torch.uint8, device=dev), torch.empty(0, H, W, device=dev),
                    torch.empty(0, 12, dtype=torch.float64))
        for p in self.pipes:
            for c in p._ordered_grid_coords[:p.curr]:
                rgb, d = p._frames[tuple(c)]
                rgbs.append(torch.round((rgb + 1.0) * 127.5).to(torch.uint8))
                depths.append(d)
                node = p.transform_grid[c[0]][c[1]]
                poses.append(np.concatenate([np.asarray(node["R"], np.float64).reshape(-1), np.asarray(node["t"], np.float64)]))
        return torch.stack(rgbs), torch.stack(depths), torch.from_numpy(np.stack(poses))

    def gather_map(self, group=None):
        """The job's fused map on every rank: one all_gather_into_tensor of the records (NCCL over NVLink), then
        `unproject_records` (prepare_pcd, :1014-1036).  Returns (xyz float64 [P,3], rgb [P,3] in [0,1], records)."""
        rgb, depth, poses = sdist.gather_scene_map(*self.local_records(), group=group)
        xyz, col = sdist.unproject_records(rgb, depth, poses.to(depth.device), self.K)
        return xyz, col, (rgb, depth, poses)
