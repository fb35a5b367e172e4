"""Restatement of VQModel.get_x / encode / decode / forward and the per-frame output
conversion of InfiniteSceneGeneration.one_step_prediction.  TEST INFRASTRUCTURE ONLY.

Reference: /root/reference/sgam/generative_sensing_module/model.py:106-269,
           /root/reference/sgam/generative_sensing_module/modules/vqvae/quantize.py:275-381,
           /root/reference/sgam/inference_pipeline.py:860-926.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import native, network
from .recipes import DDCONFIG


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def src2tgt_transforms(R_rels, t_rels):
    """model.py:188-195: 4x4 [B,N,4,4] from R_rels[B,N,3,3], t_rels[B,N,3]."""
    B, N = R_rels.shape[:2]
    T = np.tile(np.eye(4, dtype=np.float32), (B, N, 1, 1))
    T[..., :3, :3] = R_rels
    T[..., :3, 3] = t_rels
    return T


def splat(batch, zmin=False):
    """model.py:201-209 -> warp.py:193-286 on a prepare_batch_data-style numpy batch."""
    src = np.ascontiguousarray(np.transpose(batch["src_imgs"], (0, 1, 4, 2, 3)))      # model.py:169-177
    Ks = batch["Ks"]
    Kinv = torch.from_numpy(Ks.reshape(-1, 3, 3)).inverse().numpy().reshape(Ks.shape)  # warp.py:212
    T = src2tgt_transforms(batch["R_rels"], batch["t_rels"])
    return native.splat_forward(src, batch["src_depths"], Ks[:, 0], Kinv, T, zmin=zmin)


def get_x(batch, dataset, zmin=False):
    """model.py:179-269 with return_extrapolation_mask=True, no_depth_range=True.
    Returns x[B,4,H,W], extrapolation_mask[B,1,H,W] (bool), warped_depth code [B,1,H,W] (numpy)."""
    if "warped_tgt_features" in batch:                                                  # model.py:196-199
        rgb = np.asarray(batch["warped_tgt_features"], np.float32)
        depth = np.asarray(batch["warped_tgt_depth"], np.float32)[:, None]
        mask = (depth <= 0).astype(np.uint8)
    else:
        s = splat(batch, zmin=zmin)
        rgb, depth, mask = s["merge_rgb"], s["merge_depth"], s["mask"]
    code = native.depth_code(depth, mask, dataset)
    x = np.concatenate([rgb, code], 1)
    return x, mask.astype(bool), code


def vq_distances_torch(z_flat, codebook):
    """quantize.py:285-287 / :348-350, the reference's own ATen expression."""
    return torch.sum(z_flat ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) - 2 * \
        torch.einsum('bd,dn->bn', z_flat, codebook.permute(1, 0))


def quantize(sd, pre_quant, canonical=False):
    """VectorQuantizer2.forward (topk=None) and get_multiple_codewords(topk=1, sample_number=1):
    both reduce to arg-min + codebook gather (with topk=1 the multinomial draws index 0 and the
    mask pinning is the identity, quantize.py:352-367).  Returns z_q [B,D,h,w], idx [B,h,w] int64."""
    E = sd["quantize.embedding.weight"]
    B, D, h, w = pre_quant.shape
    z = pre_quant.permute(0, 2, 3, 1).contiguous().view(-1, D)
    if canonical:
        idx = torch.from_numpy(native.vq_nearest(z.numpy(), E.numpy())[0])
    else:
        idx = torch.argmin(vq_distances_torch(z, E), dim=1)
    z_q = F.embedding(idx, E).view(B, h, w, D).permute(0, 3, 1, 2).contiguous()
    return z_q, idx.view(B, h, w)


def encode(sd, x, mask=None, dd=DDCONFIG):
    """model.py:106-124 up to pre_quantized_f."""
    if mask is None:
        mask = torch.zeros(x.shape[0], 1, *x.shape[2:])
    x5 = torch.cat([x, mask.to(x.dtype)], 1)
    h = F.conv2d(x5, sd["conv_in.weight"], sd["conv_in.bias"])
    h = network.encoder(sd, h, dd)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def decode(sd, quant, dd=DDCONFIG):
    """model.py:131-134"""
    return network.decoder(sd, F.conv2d(quant, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]), dd)


@torch.no_grad()
def forward(sd, x, mask=None, dd=DDCONFIG, canonical_vq=False):
    """model.py:141-167 -> (dec[B,4,H,W], pre_quant[B,D,h,w], z_q[B,D,h,w], idx[B,h,w])."""
    x = _t(x) if not torch.is_tensor(x) else x
    if mask is not None and not torch.is_tensor(mask):
        mask = torch.from_numpy(np.ascontiguousarray(mask)).to(torch.float32)
    pre = encode(sd, x, mask, dd)
    z_q, idx = quantize(sd, pre, canonical=canonical_vq)
    dec = decode(sd, z_q, dd)
    return dec, pre, z_q, idx


def frame_outputs(dec, dataset):
    """inference_pipeline.py:893-911: uint8 RGB [H,W,3] (clip + truncate) and metric depth [H,W] of sample 0."""
    dec = dec.numpy() if torch.is_tensor(dec) else dec
    return native.pack_u8(dec[0, :3]), native.depth_decode(dec[0, 3], dataset)


@torch.no_grad()
def scene_step(sd, batch, dataset, dd=DDCONFIG, zmin=False):
    """One one_step_prediction without disk / plt (inference_pipeline.py:860-926)."""
    x, mask, code = get_x(batch, dataset, zmin=zmin)
    dec, pre, z_q, idx = forward(sd, x, mask, dd)
    rgb_u8, depth = frame_outputs(dec[:1], dataset)
    return dict(x=x, mask=mask, warped_depth=code, dec=dec.numpy(), pre_quant=pre.numpy(), z_q=z_q.numpy(),
                idx=idx.numpy(), rgb_u8=rgb_u8, depth=depth)
