"""Tensor-level wrappers over the C ABI (include/sgam_b200.h).

PyTorch is plumbing here: it owns the device memory and the stream; every function below checks its inputs,
allocates the outputs and enqueues the library's kernels on `torch.cuda.current_stream()`.  Nothing falls
back to ATen: a missing library or a CPU tensor raises.
"""
import ctypes

import torch

from . import _lib

DATASET_ID = {"clevr-infinite": 0, "google_earth": 1}
SPLAT_LAST_WRITER, SPLAT_ZMIN = 0, 1


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype=torch.float32, name="tensor"):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"{name}: expected a CUDA tensor (the SGAM hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def h2d(t, device, dtype=None):
    """Upload a small host tensor (intrinsics, poses) WITHOUT stalling the host: `.to(device)` from pageable memory is a
    cudaMemcpyAsync + stream synchronize, i.e. the host waits for every kernel already queued, which serialises the scene
    loop's host work with the previous frame's GPU work.  Staging through the caching pinned allocator keeps the copy
    asynchronous and stream ordered (the allocator holds the pinned block until the copy has run)."""
    t = torch.as_tensor(t)
    if dtype is not None:
        t = t.to(dtype)
    if t.is_cuda:
        return t.to(device)
    return t.contiguous().pin_memory().to(device, non_blocking=True)


def dataset_id(dataset):
    if dataset not in DATASET_ID:
        raise NotImplementedError(dataset)        # same error the reference raises (model.py:230-231)
    return DATASET_ID[dataset]


# ---------------------------------------------------------------------------------------------- stage (i)
def splat_forward(src_rgb, src_depth, K_tgt, Kinv_src, T_src2tgt, dataset, channels_last=False,
                  policy=SPLAT_LAST_WRITER, want_merge_depth=False, want_proj=False, want_inbounds=False,
                  workspace=None):
    """Fused forward splat + hole fill + mask + inverse-depth code (warp.py:193-286, model.py:210-237).

    src_rgb [B,N,3,H,W] (or [B,N,H,W,3] with channels_last), src_depth [B,N,H,W], K_tgt [B,3,3],
    Kinv_src [B,N,3,3], T_src2tgt [B,N,4,4].  Returns dict(x, mask, merge_depth?, proj?, inbounds?, winner)."""
    lib = _lib.load()
    B, N, H, W = src_depth.shape
    _chk(src_rgb, name="src_rgb"), _chk(src_depth, name="src_depth"), _chk(K_tgt, name="K_tgt")
    _chk(Kinv_src, name="Kinv_src"), _chk(T_src2tgt, name="T_src2tgt")
    exp = (B, N, H, W, 3) if channels_last else (B, N, 3, H, W)
    if tuple(src_rgb.shape) != exp or tuple(K_tgt.shape) != (B, 3, 3) or Kinv_src.numel() != B * N * 9 \
            or T_src2tgt.numel() != B * N * 16:
        raise RuntimeError(f"splat_forward: inconsistent shapes {tuple(src_rgb.shape)} {tuple(src_depth.shape)}")
    cs, ps = (1, 3) if channels_last else (H * W, 1)
    dev = src_depth.device
    if workspace is None:
        workspace = torch.empty(lib.sgam_splat_workspace_bytes(B, H, W) // 8, dtype=torch.int64, device=dev)
    x = torch.empty(B, 4, H, W, device=dev)
    mask = torch.empty(B, 1, H, W, dtype=torch.uint8, device=dev)
    md = torch.empty(B, 1, H, W, device=dev) if want_merge_depth else None
    proj = torch.empty(B, 4, H, W, device=dev) if want_proj else None
    inb = torch.empty(B, H * W * N, dtype=torch.uint8, device=dev) if want_inbounds else None
    _lib.check(lib.sgam_splat_forward(src_rgb.data_ptr(), cs, ps, src_depth.data_ptr(), K_tgt.data_ptr(),
                                      Kinv_src.data_ptr(), T_src2tgt.data_ptr(), B, N, H, W, policy,
                                      dataset_id(dataset), workspace.data_ptr(), x.data_ptr(), mask.data_ptr(),
                                      _ptr(md), _ptr(proj), _ptr(inb), _stream()), "sgam_splat_forward")
    return dict(x=x, mask=mask, merge_depth=md, proj=proj, inbounds=inb, winner=workspace)


def median_blur3(x):
    """3x3 zero-padded median per plane (warp.py:289-347)."""
    lib = _lib.load()
    _chk(x, name="input")
    if x.dim() != 4:
        raise ValueError(f"Invalid input shape, we expect BxCxHxW. Got: {x.shape}")   # warp.py:326-327
    out = torch.empty_like(x)
    H, W = x.shape[-2:]
    _lib.check(lib.sgam_median_blur3(x.data_ptr(), out.data_ptr(), x.numel() // (H * W), H, W, _stream()),
               "sgam_median_blur3")
    return out


def depth_code(rgb, depth, dataset):
    """get_x on a pre-warped input (model.py:196-199, 210-237): rgb [B,3,H,W], depth [B,H,W] -> x, mask."""
    lib = _lib.load()
    _chk(rgb, name="rgb"), _chk(depth, name="depth")
    B, _, H, W = rgb.shape
    x = torch.empty(B, 4, H, W, device=rgb.device)
    mask = torch.empty(B, 1, H, W, dtype=torch.uint8, device=rgb.device)
    _lib.check(lib.sgam_depth_code(rgb.data_ptr(), depth.data_ptr(), B, H, W, dataset_id(dataset), x.data_ptr(),
                                   mask.data_ptr(), _stream()), "sgam_depth_code")
    return x, mask


def inverse_warp(src_rgb, src_depth, tgt_depth, Kinv_tgt, proj, channels_last=False, want_best=False):
    """inference_pipeline.py:662-743.  src_rgb [B,N,3,H,W], src_depth [B,N,H,W], tgt_depth [B,H,W],
    Kinv_tgt [B,3,3], proj [B,N,3,4] -> warped [B,3,H,W] (and best_src [B,H,W] int32)."""
    lib = _lib.load()
    for n, t in (("src_rgb", src_rgb), ("src_depth", src_depth), ("tgt_depth", tgt_depth), ("Kinv_tgt", Kinv_tgt),
                 ("proj", proj)):
        _chk(t, name=n)
    B, N, H, W = src_depth.shape
    cs, ps = (1, 3) if channels_last else (H * W, 1)
    out = torch.empty(B, 3, H, W, device=src_depth.device)
    best = torch.empty(B, H, W, dtype=torch.int32, device=src_depth.device) if want_best else None
    _lib.check(lib.sgam_inverse_warp(src_rgb.data_ptr(), cs, ps, src_depth.data_ptr(), tgt_depth.data_ptr(),
                                     Kinv_tgt.data_ptr(), proj.data_ptr(), B, N, H, W, out.data_ptr(), _ptr(best),
                                     _stream()), "sgam_inverse_warp")
    return (out, best) if want_best else out


def unproject_points(depth, rgb_u8, K, Rt):
    """prepare_pcd (inference_pipeline.py:1014-1036) for a stack of frames on the device: depth [F,H,W] fp32,
    rgb_u8 [F,H,W,3] uint8 or None, K 3x3, Rt [F,4,4] world->camera (host, float64) -> xyz [F*H*W,3] f64 (and
    colours [F*H*W,3] f64 in [0,1])."""
    import numpy as np
    lib = _lib.load()
    _chk(depth, name="depth")
    F, H, W = depth.shape
    if rgb_u8 is not None:
        _chk(rgb_u8, torch.uint8, "rgb_u8")
    Kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(K, np.float64)))
    Rt = np.asarray(Rt, np.float64).reshape(F, 4, 4)
    Rt_inv = h2d(torch.from_numpy(np.ascontiguousarray(np.stack([np.linalg.inv(m)[:3] for m in Rt]))), depth.device)
    xyz = torch.empty(F * H * W, 3, dtype=torch.float64, device=depth.device)
    col = torch.empty(F * H * W, 3, dtype=torch.float64, device=depth.device) if rgb_u8 is not None else None
    for f0 in range(0, F, 65535):                                           # frames ride in grid.y
        f1 = min(F, f0 + 65535)
        _lib.check(lib.sgam_unproject_points(depth[f0:f1].data_ptr(), None if rgb_u8 is None else rgb_u8[f0:f1].data_ptr(),
                                             Kinv.ctypes.data, Rt_inv[f0:f1].data_ptr(), f1 - f0, H, W,
                                             xyz[f0 * H * W:].data_ptr(), None if col is None else col[f0 * H * W:].data_ptr(),
                                             _stream()), "sgam_unproject_points")
    return (xyz, col) if rgb_u8 is not None else xyz


def frame_outputs(dec, dataset, rgb_u8=None, depth=None, want_src_rgb=False):
    """inference_pipeline.py:893-911: dec [B,4,H,W] -> uint8 RGB [B,H,W,3], metric depth [B,H,W]
    (+ the fp32 source image u8/127.5-1 a later step would re-load, inference_pipeline.py:534)."""
    lib = _lib.load()
    _chk(dec, name="dec")
    B, C, H, W = dec.shape
    if C != 4:
        raise RuntimeError("frame_outputs: dec must be [B,4,H,W]")
    if rgb_u8 is None:
        rgb_u8 = torch.empty(B, H, W, 3, dtype=torch.uint8, device=dec.device)
    if depth is None:
        depth = torch.empty(B, H, W, device=dec.device)
    src = torch.empty(B, H, W, 3, device=dec.device) if want_src_rgb else None
    _lib.check(lib.sgam_frame_outputs(dec.data_ptr(), B, H, W, dataset_id(dataset), rgb_u8.data_ptr(),
                                      depth.data_ptr(), _ptr(src), _stream()), "sgam_frame_outputs")
    return (rgb_u8, depth, src) if want_src_rgb else (rgb_u8, depth)


# --------------------------------------------------------------------------------------------- stage (ii)
def vq_nearest(z_tokens, codebook, want_dmin=False, workspace=None):
    """z_tokens [T,D] (NHWC latent flattened), codebook [n_e,D] -> idx [T] int64, z_q [T,D] (, dmin [T])."""
    lib = _lib.load()
    _chk(z_tokens, name="z"), _chk(codebook, name="codebook")
    T, D = z_tokens.shape
    dev = z_tokens.device
    if workspace is None:
        workspace = torch.empty(T, dtype=torch.int64, device=dev)
    idx = torch.empty(T, dtype=torch.int64, device=dev)
    z_q = torch.empty(T, D, device=dev)
    dmin = torch.empty(T, device=dev) if want_dmin else None
    _lib.check(lib.sgam_vq_nearest(z_tokens.data_ptr(), codebook.data_ptr(), T, codebook.shape[0], D,
                                   workspace.data_ptr(), idx.data_ptr(), z_q.data_ptr(), _ptr(dmin), _stream()),
               "sgam_vq_nearest")
    return (idx, z_q, dmin) if want_dmin else (idx, z_q)


def vq_topk_sample(z_tokens, codebook, topk, samples, lat_hw, mask=None, seed=0, row0_probs=True):
    """get_multiple_codewords for topk > 1 (quantize.py:344-381).  z_tokens [T,D] = B images of lat_hw = (h, w) tokens;
    mask [B,1,H,W] / [B,H,W] uint8 or None.  Returns dict(idx [T,S], z_q [T,S,D], topk_idx [T,k], topk_p [T,k])."""
    lib = _lib.load()
    _chk(z_tokens, name="z"), _chk(codebook, name="codebook")
    T, D = z_tokens.shape
    h, w = lat_hw
    n_e = codebook.shape[0]
    dev = z_tokens.device
    mb, mw, fy, fx = 0, 0, 1, 1
    if mask is not None:
        _chk(mask, torch.uint8, "mask")
        H, W = mask.shape[-2:]
        mb, mw, fy, fx = H * W, W, H // h, W // w
    ws = torch.empty(lib.sgam_vq_topk_workspace_bytes(T, n_e) // 4, device=dev)
    out = dict(idx=torch.empty(T, samples, dtype=torch.int64, device=dev), z_q=torch.empty(T, samples, D, device=dev),
               topk_idx=torch.empty(T, topk, dtype=torch.int64, device=dev), topk_p=torch.empty(T, topk, device=dev))
    _lib.check(lib.sgam_vq_topk_sample(z_tokens.data_ptr(), codebook.data_ptr(), _ptr(mask), mb, mw, fy, fx, w, h * w, T, n_e, D,
                                       topk, samples, seed & 0xFFFFFFFFFFFFFFFF, int(row0_probs), ws.data_ptr(),
                                       out["topk_idx"].data_ptr(), out["topk_p"].data_ptr(), out["idx"].data_ptr(),
                                       out["z_q"].data_ptr(), _stream()), "sgam_vq_topk_sample")
    return out


def vq_norms(x):
    """Canonical (sequential fma) squared row norms of x [R,D]."""
    lib = _lib.load()
    _chk(x, name="x")
    out = torch.empty(x.shape[0], device=x.device)
    _lib.check(lib.sgam_vq_norms(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _stream()), "sgam_vq_norms")
    return out


class CodebookTC:
    """Per-codebook constants of the tensor-core search: split-bf16 planes, canonical norms and their maximum."""

    def __init__(self, codebook):
        self.E = _chk(codebook, name="codebook")
        self.hi, self.lo = split_weight(codebook)
        self.ee = vq_norms(codebook)
        self.ee_max = float(self.ee.max().item())


def vq_nearest_tc(z_tokens, cb, want_dmin=False):
    """Tensor-core codebook search; same results as vq_nearest (bit-identical idx, z_q, dmin)."""
    lib = _lib.load()
    _chk(z_tokens, name="z")
    T, D = z_tokens.shape
    n_e = cb.E.shape[0]
    dev = z_tokens.device
    z_hi, z_lo = split_bf16(z_tokens.view(1, 1, T, D))
    ws = torch.empty(lib.sgam_vq_tc_workspace_bytes(T, n_e) // 4, device=dev)
    idx = torch.empty(T, dtype=torch.int64, device=dev)
    z_q = torch.empty(T, D, device=dev)
    dmin = torch.empty(T, device=dev) if want_dmin else None
    _lib.check(lib.sgam_vq_nearest_tc(z_tokens.data_ptr(), z_hi.data_ptr(), z_lo.data_ptr(), cb.E.data_ptr(), cb.hi.data_ptr(),
                                      cb.lo.data_ptr(), cb.ee.data_ptr(), cb.ee_max, T, n_e, D, ws.data_ptr(), idx.data_ptr(),
                                      z_q.data_ptr(), _ptr(dmin), _stream()), "sgam_vq_nearest_tc")
    return (idx, z_q, dmin) if want_dmin else (idx, z_q)


# -------------------------------------------------------------------------------------------- stage (iii)
def stem_conv(x, mask, w, bias):
    """cat(x, mask) -> 1x1 conv 5->4 (model.py:106-113).  x [B,4,H,W] NCHW -> [B,H,W,4] NHWC."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(w, name="w"), _chk(bias, name="bias")
    if mask is not None:
        _chk(mask, torch.uint8, "mask")
    B, _, H, W = x.shape
    y = torch.empty(B, H, W, 4, device=x.device)
    _lib.check(lib.sgam_stem_conv(x.data_ptr(), _ptr(mask), w.data_ptr(), bias.data_ptr(), B, H, W, y.data_ptr(),
                                  _stream()), "sgam_stem_conv")
    return y


def conv_out_hw(H, W, ksize, stride, pad_mode, upsample):
    Hl, Wl = H << upsample, W << upsample
    if pad_mode == 0:
        p = ksize // 2
        return (Hl + 2 * p - ksize) // stride + 1, (Wl + 2 * p - ksize) // stride + 1
    return (Hl + 1 - ksize) // stride + 1, (Wl + 1 - ksize) // stride + 1


def conv2d(x, w, bias, residual=None, ksize=3, stride=1, pad_mode=0, upsample=0, out_nchw=False, out=None):
    """x [B,H,W,Cin] NHWC, w [Cout, k*k*Cin] (K-major), bias [Cout] -> [B,Ho,Wo,Cout] (or NCHW for the head)."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(w, name="w"), _chk(bias, name="bias")
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    if w.shape[1] != ksize * ksize * Cin:
        raise RuntimeError(f"conv2d: weight {tuple(w.shape)} does not match k={ksize} Cin={Cin}")
    Ho, Wo = conv_out_hw(H, W, ksize, stride, pad_mode, upsample)
    if out is None:
        out = torch.empty((B, Cout, Ho, Wo) if out_nchw else (B, Ho, Wo, Cout), device=x.device)
    if residual is not None:
        _chk(residual, name="residual")
    _lib.check(lib.sgam_conv2d(x.data_ptr(), w.data_ptr(), bias.data_ptr(), _ptr(residual), out.data_ptr(), B, H, W,
                               Cin, Cout, ksize, stride, pad_mode, upsample, int(out_nchw), _stream()), "sgam_conv2d")
    return out


def groupnorm(x, gamma, beta, swish, out=None, workspace=None):
    """GroupNorm(32, C, eps=1e-6) (+ swish) on NHWC x [B,H,W,C] (model.py:29-35)."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(gamma, name="gamma"), _chk(beta, name="beta")
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    S = lib.sgam_gn_splits(HW)
    if workspace is None:
        workspace = torch.empty(B * S * 64, dtype=torch.float64, device=x.device)
    if out is None:
        out = torch.empty_like(x)
    _lib.check(lib.sgam_groupnorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(),
                                  workspace.data_ptr(), B, HW, C, int(swish), _stream()), "sgam_groupnorm")
    return out


def gemm_nt(A, Bm, bias_m=None, alpha=1.0, out=None):
    """C = alpha * A . B^T (+ bias_m per row).  A [batch,M,K] or [M,K]; B [batch,N,K] or [N,K]."""
    lib = _lib.load()
    _chk(A, name="A"), _chk(Bm, name="B")
    squeeze = A.dim() == 2 and Bm.dim() == 2
    A3 = A.unsqueeze(0) if A.dim() == 2 else A
    B3 = Bm.unsqueeze(0) if Bm.dim() == 2 else Bm
    batch = max(A3.shape[0], B3.shape[0])
    _, M, K = A3.shape
    N = B3.shape[1]
    if B3.shape[2] != K or A3.shape[0] not in (1, batch) or B3.shape[0] not in (1, batch):
        raise RuntimeError(f"gemm_nt: shapes {tuple(A.shape)} x {tuple(Bm.shape)}")
    sA = M * K if A3.shape[0] == batch else 0          # a 2-D operand is shared by every batch element
    sB = N * K if B3.shape[0] == batch else 0
    if out is None:
        out = torch.empty(batch, M, N, device=A.device)
    _lib.check(lib.sgam_gemm_nt(A3.data_ptr(), B3.data_ptr(), out.data_ptr(), _ptr(bias_m), batch, M, N, K,
                                sA, sB, M * N, float(alpha), _stream()), "sgam_gemm_nt")
    return out[0] if squeeze and out.dim() == 3 else out


def softmax_rows_(x):
    """In-place softmax over the last dimension (model.py:181)."""
    lib = _lib.load()
    _chk(x, name="x")
    cols = x.shape[-1]
    _lib.check(lib.sgam_softmax_rows(x.data_ptr(), x.numel() // cols, cols, _stream()), "sgam_softmax_rows")
    return x


# ------------------------------------------------------------------------- stage (iii), tensor-core path
def _bf16_pair(shape, device):
    return (torch.empty(shape, dtype=torch.bfloat16, device=device), torch.empty(shape, dtype=torch.bfloat16, device=device))


def stem_conv_split(x, mask, w, bias, cpad=64):
    """Stem (model.py:106-113) -> split bf16 NHWC [B,H,W,cpad] with channels 4.. zero (feeds the tensor-core conv_in)."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(w, name="w"), _chk(bias, name="bias")
    if mask is not None:
        _chk(mask, torch.uint8, "mask")
    B, _, H, W = x.shape
    hi, lo = _bf16_pair((B, H, W, cpad), x.device)
    _lib.check(lib.sgam_stem_conv_split(x.data_ptr(), _ptr(mask), w.data_ptr(), bias.data_ptr(), B, H, W, cpad, hi.data_ptr(),
                                        lo.data_ptr(), _stream()), "sgam_stem_conv_split")
    return hi, lo


def split_bf16(x, upsample=0):
    """fp32 NHWC [B,H,W,C] -> (hi, lo) bf16 [B,H<<up,W<<up,C] with x = hi + lo (+ fused nearest x2 up-sampling)."""
    lib = _lib.load()
    _chk(x, name="x")
    B, H, W, C = x.shape
    hi, lo = _bf16_pair((B, H << upsample, W << upsample, C), x.device)
    _lib.check(lib.sgam_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), B, H, W, C, int(upsample), _stream()),
               "sgam_split_bf16")
    return hi, lo


def split_weight(w, pad_rows_to=1):
    """Host-side one-off split of a weight matrix [N, K] fp32 -> (hi, lo) bf16 (same rounding as the device split);
    rows are zero-padded to a multiple of `pad_rows_to` (32 for conv weights: the 4-channel head)."""
    if pad_rows_to > 1 and w.shape[-2] % pad_rows_to:
        pad = pad_rows_to - w.shape[-2] % pad_rows_to
        w = torch.cat([w, w.new_zeros(*w.shape[:-2], pad, w.shape[-1])], dim=-2)
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def groupnorm_split(x, gamma, beta, swish, workspace=None):
    """GroupNorm(32, C, eps=1e-6) (+ swish) -> (hi, lo) bf16 NHWC."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(gamma, name="gamma"), _chk(beta, name="beta")
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    hi, lo = _bf16_pair(tuple(x.shape), x.device)
    partial = getattr(x, "gn_partial", None)
    if partial is not None and x.dim() == 4:
        _lib.check(lib.sgam_groupnorm_split_fused(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                                  partial.data_ptr(), B, x.shape[1], x.shape[2], C, int(swish), _stream()),
                   "sgam_groupnorm_split_fused")
        return hi, lo
    partial64 = getattr(x, "gn_partial64", None)
    if partial64 is not None:
        _lib.check(lib.sgam_groupnorm_split_apply(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                                  partial64.data_ptr(), B, HW, C, int(swish), _stream()), "sgam_groupnorm_split_apply")
        return hi, lo
    if workspace is None:
        workspace = torch.empty(B * lib.sgam_gn_splits(HW) * 64, dtype=torch.float64, device=x.device)
    _lib.check(lib.sgam_groupnorm_split(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), hi.data_ptr(), lo.data_ptr(),
                                        workspace.data_ptr(), B, HW, C, int(swish), _stream()), "sgam_groupnorm_split")
    return hi, lo


def softmax_split(s):
    """Row softmax of fp32 scores [..., cols] -> (hi, lo) bf16 probabilities."""
    lib = _lib.load()
    _chk(s, name="scores")
    cols = s.shape[-1]
    hi, lo = _bf16_pair(tuple(s.shape), s.device)
    _lib.check(lib.sgam_softmax_split(s.data_ptr(), hi.data_ptr(), lo.data_ptr(), s.numel() // cols, cols, _stream()),
               "sgam_softmax_split")
    return hi, lo


def tc_supported_conv(H, W, Cin, Cout, ksize, stride):
    return bool(_lib.load().sgam_tc_supported_conv(H, W, Cin, Cout, ksize, stride))


def conv2d_tc_splitk_floats(B, H, W, Cin, Cout, ksize, stride=1):
    """Workspace floats conv2d_tc would use for a split K loop at this shape (0 = the K loop is not split)."""
    return int(_lib.load().sgam_conv2d_tc_splitk_floats(B, H, W, Cin, Cout, ksize, stride))


def conv2d_tc(x, w, bias, residual=None, ksize=3, stride=1, cout=None, out_nchw=False, out_f32=True, out_split=False, nsplit=3,
              gn_stats=False):
    """Conv on tcgen05 (stride 1 symmetric pad, or stride 2 = Downsample).  x = (hi, lo) bf16 [B,H,W,Cin];
    w = (hi, lo) bf16 [ceil32(Cout), k*k*Cin]; bias fp32 [Cout].  Returns fp32 y [B,Ho,Wo,Cout] (or [B,Cout,Ho,Wo] with
    out_nchw) and/or the (hi, lo) pair, as requested."""
    lib = _lib.load()
    x_hi, x_lo = x
    w_hi, w_lo = w
    for n, t in (("x_hi", x_hi), ("x_lo", x_lo), ("w_hi", w_hi), ("w_lo", w_lo)):
        _chk(t, torch.bfloat16, n)
    B, H, W, Cin = x_hi.shape
    Cout = w_hi.shape[0] if cout is None else cout
    if w_hi.shape[1] != ksize * ksize * Cin or w_hi.shape[0] != (Cout + 31) // 32 * 32:
        raise RuntimeError(f"conv2d_tc: weight {tuple(w_hi.shape)} does not match k={ksize} Cin={Cin} Cout={Cout}")
    Ho, Wo = H // stride, W // stride
    y = torch.empty((B, Cout, Ho, Wo) if out_nchw else (B, Ho, Wo, Cout), device=x_hi.device) if out_f32 else None
    pair = _bf16_pair((B, Ho, Wo, Cout), x_hi.device) if out_split else (None, None)
    if residual is not None:
        _chk(residual, name="residual")
    partial = splitk = partial64 = None
    if out_f32 and not out_split and not out_nchw and Cout % 32 == 0:
        n_ws = lib.sgam_conv2d_tc_splitk_floats(B, H, W, Cin, Cout, ksize, stride)
        if n_ws > 0:                    # under-filled grid: split the K loop; the reduce kernel also takes the statistics
            splitk = torch.empty(n_ws, device=x_hi.device)
            if gn_stats and Cout % 128 == 0 and Cout <= 1024:
                partial64 = torch.empty(B * lib.sgam_gn_splits(Ho * Wo) * 64, dtype=torch.float64, device=x_hi.device)
    if splitk is None and gn_stats and out_f32 and not out_nchw and Cout % 128 == 0 and Cout <= 512:
        partial = torch.empty(lib.sgam_tc_gn_partial_floats(B, Ho, Wo), device=x_hi.device)
    written = ctypes.c_int(0)
    _lib.check(lib.sgam_conv2d_tc(x_hi.data_ptr(), x_lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), _ptr(bias),
                                  _ptr(residual), _ptr(y), _ptr(pair[0]), _ptr(pair[1]), B, H, W, Cin, Cout, ksize,
                                  stride, int(out_nchw), nsplit, _ptr(partial), _ptr(splitk), _ptr(partial64),
                                  ctypes.byref(written), _stream()), "sgam_conv2d_tc")
    # the library reports which statistics buffer it filled; attach exactly that one (never a guess)
    if written.value == 1:
        y.gn_partial = partial          # GroupNorm statistics of y, fused into the epilogue (consumed by groupnorm_split)
    elif written.value == 2:
        y.gn_partial64 = partial64      # ... or into the split-K reduction
    if out_f32 and out_split:
        return y, pair
    return y if out_f32 else pair


def _rows_pitch(t, name):
    """Row pitch (elements) of a bf16 operand [batch, R, K] / [R, K] whose rows may be column slices of a wider tensor."""
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.bfloat16):
        raise RuntimeError(f"{name}: expected a CUDA bf16 tensor")
    ld = t.stride(-2)
    if t.stride(-1) != 1 or ld < t.shape[-1] or ld % 8 or t.data_ptr() % 16 or (t.dim() == 3 and t.shape[0] > 1 and t.stride(0) != t.shape[1] * ld):
        raise RuntimeError(f"{name}: shape {tuple(t.shape)} strides {t.stride()} is not a row-pitched operand")
    return int(ld)


def gemm_nt_tc(a, b, bias_m=None, alpha=1.0, out_f32=True, out_split=False, nsplit=3):
    """C = alpha * A . B^T (+ bias_m per row) on tcgen05.  a = (hi, lo) [batch, M, K] or [M, K] (shared);
    b = (hi, lo) [batch, N, K] or [N, K] (shared); rows may be column slices of a wider tensor (one pitch per operand).
    Output [batch, M, N]."""
    lib = _lib.load()
    a_hi, a_lo = a
    b_hi, b_lo = b
    lda, ldb = _rows_pitch(a_hi, "a_hi"), _rows_pitch(b_hi, "b_hi")
    if _rows_pitch(a_lo, "a_lo") != lda or _rows_pitch(b_lo, "b_lo") != ldb:
        raise RuntimeError("gemm_nt_tc: the hi and lo planes of an operand must share one row pitch")
    a_b, b_b = a_hi.dim() == 3, b_hi.dim() == 3
    batch = a_hi.shape[0] if a_b else (b_hi.shape[0] if b_b else 1)
    M, K = a_hi.shape[-2:]
    N = b_hi.shape[-2]
    if b_hi.shape[-1] != K or (a_b and b_b and a_hi.shape[0] != b_hi.shape[0]):
        raise RuntimeError(f"gemm_nt_tc: shapes {tuple(a_hi.shape)} x {tuple(b_hi.shape)}")
    dev = a_hi.device
    y = torch.empty(batch, M, N, device=dev) if out_f32 else None
    pair = _bf16_pair((batch, M, N), dev) if out_split else (None, None)
    _lib.check(lib.sgam_gemm_nt_tc(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), _ptr(bias_m),
                                   _ptr(y), _ptr(pair[0]), _ptr(pair[1]), batch, M, N, K, int(a_b), int(b_b), float(alpha),
                                   nsplit, 0 if lda == K else lda, 0 if ldb == K else ldb, _stream()), "sgam_gemm_nt_tc")
    if out_f32 and out_split:
        return y, pair
    return y if out_f32 else pair


def attention_tc_supported(B, T, C):
    return bool(_lib.load().sgam_attention_tc_supported(B, T, C))


def attention_tc(q, k, vT, scale, kv_splits=None):
    """Fused softmax(q . k^T * scale) . v on tcgen05 (AttnBlock, diffusionmodules/model.py:168-192): the [B,T,T] scores
    never leave the SM.  q, k = (hi, lo) [B,T,C] -- contiguous, or the column halves of one [B,T,2C] tensor (qkv_tc);
    vT = (hi, lo) [B,C,T]; returns o = (hi, lo) [B,T,C].  kv_splits: key splits per query tile (None = the library's
    choice: 1 unless the batch is too small to fill the SM pairs)."""
    lib = _lib.load()
    B, T, C = q[0].shape
    ld = q[0].stride(1)
    for n, t in (("q_hi", q[0]), ("q_lo", q[1]), ("k_hi", k[0]), ("k_lo", k[1])):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.bfloat16):
            raise RuntimeError(f"attention_tc: {n} must be a CUDA bf16 tensor")
        if tuple(t.shape) != (B, T, C) or t.stride(2) != 1 or t.stride(1) != ld or t.stride(0) != T * ld or ld % 8 or t.data_ptr() % 16:
            raise RuntimeError(f"attention_tc: {n} shape {tuple(t.shape)} strides {t.stride()} (rows must share one pitch, a multiple of 8)")
    for n, t in (("vT_hi", vT[0]), ("vT_lo", vT[1])):
        _chk(t, torch.bfloat16, n)
    if tuple(vT[0].shape) != (B, C, T):
        raise RuntimeError(f"attention_tc: shapes q {tuple(q[0].shape)} k {tuple(k[0].shape)} vT {tuple(vT[0].shape)}")
    if kv_splits is None:
        kv_splits = lib.sgam_attention_tc_splits(B, T)
    o = _bf16_pair((B, T, C), q[0].device)
    ws = None
    if kv_splits > 1:
        ws = torch.empty(lib.sgam_attention_tc_workspace_bytes(B, T, kv_splits) // 4, device=q[0].device)
    _lib.check(lib.sgam_attention_tc(q[0].data_ptr(), q[1].data_ptr(), k[0].data_ptr(), k[1].data_ptr(), vT[0].data_ptr(),
                                     vT[1].data_ptr(), o[0].data_ptr(), o[1].data_ptr(), B, T, C, float(scale), int(kv_splits),
                                     _ptr(ws), 0 if ld == C else int(ld), _stream()), "sgam_attention_tc")
    return o


def gn_conv2d_tc_supported(B, H, W, Cin, Cout):
    return bool(_lib.load().sgam_gn_conv2d_tc_supported(B, H, W, Cin, Cout))


def gn_conv2d_tc(x, gamma, beta, w, bias, residual=None, out_f32=True, out_split=False, nsplit=3, gn_stats=True):
    """GroupNorm + swish + 3x3 conv in one tcgen05 kernel (the normalisation runs inside the conv's operand path).  x fp32 NHWC
    [B,H,W,Cin] carrying the partial sums of its producer (`x.gn_partial`); w = (hi, lo) [Cout, 9*Cin]; returns like conv2d_tc."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(gamma, name="gamma"), _chk(beta, name="beta"), _chk(w[0], torch.bfloat16, "w_hi"), _chk(w[1], torch.bfloat16, "w_lo")
    partial_in = getattr(x, "gn_partial", None)
    if partial_in is None:
        raise RuntimeError("gn_conv2d_tc: x carries no fused GroupNorm statistics (gn_partial)")
    B, H, W, Cin = x.shape
    Cout = w[0].shape[0]
    if w[0].shape[1] != 9 * Cin:
        raise RuntimeError(f"gn_conv2d_tc: weight {tuple(w[0].shape)} does not match Cin={Cin}")
    if residual is not None:
        _chk(residual, name="residual")
    y = torch.empty(B, H, W, Cout, device=x.device) if out_f32 else None
    pair = _bf16_pair((B, H, W, Cout), x.device) if out_split else (None, None)
    partial = torch.empty(lib.sgam_tc_gn_partial_floats(B, H, W), device=x.device) if (gn_stats and out_f32 and Cout % 128 == 0 and Cout <= 512) else None
    _lib.check(lib.sgam_gn_conv2d_tc(x.data_ptr(), partial_in.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w[0].data_ptr(), w[1].data_ptr(),
                                     _ptr(bias), _ptr(residual), _ptr(y), _ptr(pair[0]), _ptr(pair[1]), B, H, W, Cin, Cout, int(nsplit),
                                     _ptr(partial), _stream()), "sgam_gn_conv2d_tc")
    if partial is not None:
        y.gn_partial = partial
    if out_f32 and out_split:
        return y, pair
    return y if out_f32 else pair


def qkv_tc_supported(B, H, W, C):
    return bool(_lib.load().sgam_qkv_tc_supported(B, H, W, C))


def qkv_tc(xs, w, bias, nsplit=3):
    """The q / k / v projections of an AttnBlock (diffusionmodules/model.py:158-175) as one tcgen05 GEMM.  xs = (hi, lo)
    [B,H,W,C]; w = (hi, lo) [3C, C] (rows of q, k, v); bias [3C].  Returns q, k = (hi, lo) views [B,T,C] of one [B,T,2C]
    tensor and vT = (hi, lo) [B,C,T]."""
    lib = _lib.load()
    x_hi, x_lo = xs
    _chk(x_hi, torch.bfloat16, "x_hi"), _chk(x_lo, torch.bfloat16, "x_lo"), _chk(w[0], torch.bfloat16, "w_hi"), _chk(w[1], torch.bfloat16, "w_lo")
    _chk(bias, name="bias")
    B, H, W, C = x_hi.shape
    if tuple(w[0].shape) != (3 * C, C) or bias.numel() != 3 * C:
        raise RuntimeError(f"qkv_tc: weight {tuple(w[0].shape)} / bias {tuple(bias.shape)} for C = {C}")
    T = H * W
    qk = _bf16_pair((B, T, 2 * C), x_hi.device)
    vT = _bf16_pair((B, C, T), x_hi.device)
    _lib.check(lib.sgam_qkv_tc(x_hi.data_ptr(), x_lo.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), bias.data_ptr(), qk[0].data_ptr(),
                               qk[1].data_ptr(), vT[0].data_ptr(), vT[1].data_ptr(), B, H, W, C, int(nsplit), _stream()), "sgam_qkv_tc")
    q = (qk[0][..., :C], qk[1][..., :C])
    k = (qk[0][..., C:], qk[1][..., C:])
    return q, k, vT


def gn_head_conv(x, gamma, beta, w_t, bias):
    """decoder.norm_out + swish + decoder.conv_out (diffusionmodules/model.py:534-538) in one fp32 kernel.  x fp32 NHWC
    [B,H,W,128] carrying the GroupNorm partial sums of its producing conv (`x.gn_partial`); w_t [9,128,4]; -> [B,4,H,W]."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(gamma, name="gamma"), _chk(beta, name="beta"), _chk(w_t, name="w_t"), _chk(bias, name="bias")
    partial = getattr(x, "gn_partial", None)
    if partial is None:
        raise RuntimeError("gn_head_conv: x carries no fused GroupNorm statistics (gn_partial)")
    B, H, W, C = x.shape
    Cout = w_t.shape[-1]
    y = torch.empty(B, Cout, H, W, device=x.device)
    _lib.check(lib.sgam_gn_head_conv(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), partial.data_ptr(), w_t.data_ptr(),
                                     bias.data_ptr(), y.data_ptr(), B, H, W, C, Cout, _stream()), "sgam_gn_head_conv")
    return y


def subpixel_weights(w, cin):
    """Packed 3x3 weights [Cout, 9*Cin] (K index (kh*3+kw)*Cin+ci) -> the four parity matrices of the sub-pixel form of
    `nearest x2 then conv3x3` [4, Cout, 4*Cin]: parity (py, px), K index (dy*2+dx)*Cin+ci.  Rows: py = 0 reads low-res rows
    {i-1: kh 0, i: kh 1 + kh 2}; py = 1 reads {i: kh 0 + kh 1, i+1: kh 2}; columns alike.  Taps are summed in fp32."""
    cout = w.shape[0]
    w9 = w.view(cout, 3, 3, cin)
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = w.new_zeros(4, cout, 2, 2, cin)
    for py in (0, 1):
        for px in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    acc = None
                    for kh in groups[py][dy]:
                        for kw in groups[px][dx]:
                            acc = w9[:, kh, kw] if acc is None else acc + w9[:, kh, kw]
                    out[2 * py + px, :, dy, dx] = acc
    return out.view(4, cout, 4 * cin).contiguous()


def conv2d_tc_up2_supported(B, H, W, Cin, Cout):
    return bool(_lib.load().sgam_conv2d_tc_up2_supported(B, H, W, Cin, Cout))


def conv2d_tc_up2(x, w, bias, nsplit=3, gn_stats=True):
    """Upsample (nearest x2 + 3x3 conv) in sub-pixel form on tcgen05.  x = (hi, lo) bf16 [B,H,W,Cin] at LOW resolution;
    w = (hi, lo) bf16 [4, Cout, 4*Cin] from `subpixel_weights`; returns fp32 y [B,2H,2W,Cout] (+ fused GroupNorm statistics)."""
    lib = _lib.load()
    x_hi, x_lo = x
    w_hi, w_lo = w
    for n, t in (("x_hi", x_hi), ("x_lo", x_lo), ("w_hi", w_hi), ("w_lo", w_lo)):
        _chk(t, torch.bfloat16, n)
    B, H, W, Cin = x_hi.shape
    Cout = w_hi.shape[1]
    if tuple(w_hi.shape) != (4, Cout, 4 * Cin):
        raise RuntimeError(f"conv2d_tc_up2: weight {tuple(w_hi.shape)} does not match Cin={Cin}")
    y = torch.empty(B, 2 * H, 2 * W, Cout, device=x_hi.device)
    partial = None
    if gn_stats and Cout % 128 == 0 and Cout <= 512:
        partial = torch.empty(lib.sgam_tc_gn_partial_floats(B, 2 * H, 2 * W), device=x_hi.device)
    _lib.check(lib.sgam_conv2d_tc_up2(x_hi.data_ptr(), x_lo.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), _ptr(bias), y.data_ptr(),
                                      B, H, W, Cin, Cout, nsplit, _ptr(partial), _stream()), "sgam_conv2d_tc_up2")
    if partial is not None:
        y.gn_partial = partial
    return y


def stem_conv_in(x, mask, w1, b1, w3, b3, gn_stats=True):
    """Stem (model.py:106-113) + encoder.conv_in (3x3, 4 -> 128) + GroupNorm partial sums in one fp32 kernel.
    x [B,4,H,W] NCHW, mask [B,H,W] uint8 or None -> fp32 NHWC [B,H,W,128] (with `.gn_partial`)."""
    lib = _lib.load()
    _chk(x, name="x"), _chk(w1, name="w1"), _chk(b1, name="b1"), _chk(w3, name="w3"), _chk(b3, name="b3")
    if mask is not None:
        _chk(mask, torch.uint8, "mask")
    B, _, H, W = x.shape
    Cout = w3.shape[0]
    y = torch.empty(B, H, W, Cout, device=x.device)
    partial = torch.empty(lib.sgam_tc_gn_partial_floats(B, H, W), device=x.device) if gn_stats else None
    _lib.check(lib.sgam_stem_conv_in(x.data_ptr(), _ptr(mask), w1.data_ptr(), b1.data_ptr(), w3.data_ptr(), b3.data_ptr(), y.data_ptr(),
                                     _ptr(partial), B, H, W, Cout, _stream()), "sgam_stem_conv_in")
    if partial is not None:
        y.gn_partial = partial
    return y
