"""Multi-GPU plumbing: independent trajectories shard across ranks (one process per GPU, weights replicated), no
collective inside the scene loop; ONE all-gather at the end assembles the compact per-frame records
(uint8 RGB + fp32 depth + float64 world-to-camera pose) of every trajectory on every rank.
Works on NCCL (GPU tensors) and gloo (CPU tensors; used by the world_size-2 tests)."""
import numpy as np
import torch
import torch.distributed as dist

POSE_BYTES = 12 * 8          # R (9) + t (3) as float64


def shard(n_units, rank, world_size):
    """Trajectory t runs on rank t mod world_size."""
    return [t for t in range(n_units) if t % world_size == rank]


def record_bytes(H, W):
    return H * W * 3 + H * W * 4 + POSE_BYTES


def pack_records(rgb_u8, depth, poses):
    """rgb_u8 [F,H,W,3] uint8, depth [F,H,W] fp32, poses [F,12] float64 -> [F, record_bytes] uint8 (same device)."""
    F, H, W = depth.shape
    return torch.cat([rgb_u8.reshape(F, H * W * 3), depth.contiguous().view(torch.uint8).reshape(F, H * W * 4),
                      poses.to(rgb_u8.device).contiguous().view(torch.uint8).reshape(F, POSE_BYTES)], dim=1).contiguous()


def unpack_records(buf, H, W):
    F = buf.shape[0]
    n_rgb, n_d = H * W * 3, H * W * 4
    rgb = buf[:, :n_rgb].reshape(F, H, W, 3)
    depth = buf[:, n_rgb:n_rgb + n_d].contiguous().view(torch.float32).reshape(F, H, W)
    poses = buf[:, n_rgb + n_d:].contiguous().view(torch.float64).reshape(F, 12)
    return rgb, depth, poses


def gather_scene_map(rgb_u8, depth, poses, group=None):
    """Every rank contributes its F_r frames (F_r may differ between ranks and may be 0: `shard` gives unequal counts
    whenever the trajectory count is not a multiple of the world size); returns (rgb [sum F_r,...], depth, poses)
    ordered by rank.  The frame counts are exchanged first (one 8-byte all-gather), the records are padded to the largest
    count, ONE all_gather_into_tensor moves them over NVLink / NVSwitch (gloo on CPU), and the padding is dropped."""
    H, W = depth.shape[-2:]
    local = pack_records(rgb_u8, depth, poses)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_records(local, H, W)
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c) for c in counts.tolist()]
    fmax = max(counts)
    if fmax == 0:
        return unpack_records(local, H, W)
    if local.shape[0] < fmax:
        local = torch.cat([local, local.new_zeros(fmax - local.shape[0], local.shape[1])])
    out = torch.empty((world * fmax, local.shape[1]), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    if any(c != fmax for c in counts):
        out = torch.cat([out[r * fmax:r * fmax + c] for r, c in enumerate(counts)])
    return unpack_records(out, H, W)


def unproject_records(rgb_u8, depth, poses, K):
    """prepare_pcd (inference_pipeline.py:1014-1036) for a stack of gathered records: float64 world points [F*H*W,3]
    and colours in [0,1].  Device records go through the sgam_unproject_points kernel (bit-identical to the
    reference's numpy); host records (the gloo tests) through the same formula in float64 torch ops."""
    F, H, W = depth.shape
    dev = depth.device
    if depth.is_cuda:
        from . import ops
        Rt = np.tile(np.eye(4), (F, 1, 1))
        pn = poses.detach().cpu().numpy()
        Rt[:, :3, :3] = pn[:, :9].reshape(F, 3, 3)
        Rt[:, :3, 3] = pn[:, 9:]
        return ops.unproject_points(depth.contiguous(), rgb_u8.contiguous(), K, Rt)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64, device=dev),
                            torch.arange(W, dtype=torch.float64, device=dev), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=torch.float64, device=dev)])
    Kinv = torch.from_numpy(np.linalg.inv(np.asarray(K, np.float64))).to(dev)
    cam = (Kinv @ pix)[None] * depth.reshape(F, 1, -1).double()                       # [F,3,HW]
    Rt = torch.eye(4, dtype=torch.float64, device=dev).repeat(F, 1, 1)
    Rt[:, :3, :3] = poses[:, :9].reshape(F, 3, 3)
    Rt[:, :3, 3] = poses[:, 9:]
    world = torch.linalg.inv(Rt) @ torch.cat([cam, torch.ones(F, 1, H * W, dtype=torch.float64, device=dev)], 1)
    return world[:, :3].permute(0, 2, 1).reshape(-1, 3), rgb_u8.reshape(-1, 3).double() / 255.0
