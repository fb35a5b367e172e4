"""Mirror of the reference's operator module sgam/point_rendering/warp.py on B200 kernels: same function names,
argument meaning and return structure for the three functions the hot path uses."""
import torch

from . import ops


def median_blur(input, kernel_size=(3, 3)):
    """warp.py:307-347 (3x3 only: the only size the reference ever passes, warp.py:274-275)."""
    if not isinstance(input, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(input)}")
    if not len(input.shape) == 4:
        raise ValueError(f"Invalid input shape, we expect BxCxHxW. Got: {input.shape}")
    if tuple(kernel_size) != (3, 3):
        raise NotImplementedError("median_blur: only the 3x3 window of the hot path is implemented")
    return ops.median_blur3(input.contiguous())


@torch.no_grad()
def render_projection_from_srcs_fast(src_features, src_depths, tgt_intrinsic, src_intrinsics, src2tgt_transform,
                                     src_num, dynamic_masks=None, depth_range=None, parallel=False,
                                     policy=ops.SPLAT_LAST_WRITER):
    """warp.py:193-286.  src_features [B,N,3,H,W], src_depths [B,N,H,W], tgt_intrinsic [B,3,3],
    src_intrinsics [B,N,3,3], src2tgt_transform [B,N,4,4] (CUDA fp32).
    Returns the reference's 7-tuple (merge_depths, merge_feats, extrapolation_mask, inbounds_mask,
    fused_features, idx, projected_features); the two debugging members the pipeline never reads
    (fused_features, idx) are None.  Deterministic for either value of `parallel`."""
    if dynamic_masks is not None or depth_range is not None:
        raise NotImplementedError("dynamic_masks / depth_range are training-time options of the reference")
    B, N, H, W = src_depths.shape
    dev = src_depths.device
    Kinv = src_intrinsics.detach().reshape(-1, 3, 3).to("cpu", torch.float32).inverse().to(dev)   # warp.py:212, CPU LAPACK
    out = ops.splat_forward(src_features.contiguous(), src_depths.contiguous(), tgt_intrinsic.contiguous(),
                            Kinv.contiguous(), src2tgt_transform.contiguous(), "clevr-infinite",
                            policy=policy, want_merge_depth=True, want_proj=True, want_inbounds=True)
    merge_feats = out["x"][:, :3]
    inbounds = out["inbounds"].view(torch.bool).reshape(-1)
    return (out["merge_depth"], merge_feats, out["mask"].view(torch.bool), inbounds, None, None, out["proj"][:, :3])
