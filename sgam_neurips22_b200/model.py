"""Host-side mirror of the reference's `VQModel` (sgam/generative_sensing_module/model.py:18-269) on B200 kernels.

Same constructor keywords, same `get_x` / `encode` / `decode` / `decode_code` / `forward` / `init_from_ckpt`
signatures, return structures and checkpoint layout, so `main_scene_generation.py` and
`InfiniteSceneGeneration` drive it unchanged.  What differs is underneath: no Lightning, no LPIPS / discriminator
(training only; their checkpoint keys are tolerated and ignored), parameters are plain tensors named exactly like
the checkpoint, and every operator is a libsgam_b200 kernel.  There is no CPU path: calling the model before
`.to('cuda')` raises.
"""
import math
import os

import numpy as np
import torch

from . import arch, ops
from .vqgan import VQGANEngine


def _get(cfg, key, default=None):
    """Config access that tolerates dicts, OmegaConf-like objects and missing keys (the reference relies on
    OmegaConf returning None for absent keys: model.py:60)."""
    if cfg is None:
        return default
    try:
        v = cfg[key]
    except (KeyError, IndexError, TypeError, AttributeError):
        v = getattr(cfg, key, default)
    return default if v is None else v


def _plain(cfg):
    """Config -> plain dict / list.  The real OmegaConf is detected explicitly: under the reference-pinned
    omegaconf==2.0.0 a DictConfig answers None for ANY missing attribute, so duck-typing on `to_container` would call
    None; with omegaconf >= 2.1 it raises.  Only the built-in ConfigNode shim has a `to_container` method."""
    try:
        from omegaconf import OmegaConf as _OC  # noqa: WPS433 (optional dependency, absent from the target image)
        if _OC.is_config(cfg):
            return _OC.to_container(cfg, resolve=True)
    except ImportError:
        pass
    from .config import ConfigNode
    if isinstance(cfg, ConfigNode):
        return cfg.to_container()
    if isinstance(cfg, dict):
        return {k: _plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_plain(v) for v in cfg]
    return cfg


class VQModel(torch.nn.Module):
    def __init__(self,
                 ddconfig,
                 data_config,
                 lossconfig,
                 n_embed,
                 embed_dim,
                 phase=None,
                 ckpt_path=None,
                 ignore_keys=['loss.discriminator'],
                 image_key="image",
                 colorize_nlabels=None,
                 logdir=None,
                 use_extrapolation_mask=True,
                 vq_step_threshold=0,
                 monitor=None,
                 remap=None,
                 sane_index_shape=False,
                 online_kmeans_config=None,
                 batch_size=None,
                 depth_range=None
                 ):
        super().__init__()
        self.phase = phase
        self.online_kmeans_config = online_kmeans_config
        self.data_config = data_config
        self.logdir = logdir
        self.depth_range = depth_range
        self.n_embed = int(n_embed)
        self.embed_dim = int(embed_dim)
        self.use_extrapolation_mask = use_extrapolation_mask
        self.vq_step_threshold = vq_step_threshold
        self.image_key = image_key
        self.global_step = 0
        self.use_rgbd_integration = False
        self.splat_policy = ops.SPLAT_LAST_WRITER          # reference semantics; ops.SPLAT_ZMIN = z-buffered splat
        self.engine_mode = os.environ.get("SGAM_ENGINE_MODE", "tc")   # "tc": tcgen05 split-bf16 convs; "simt": exact fp32
        # forward() replays one CUDA graph per input shape (~340 kernel nodes) instead of launching op by op
        self.use_cuda_graph = os.environ.get("SGAM_CUDA_GRAPH", "1") != "0"
        self._graphs = {}
        if monitor is not None:
            self.monitor = monitor
        if remap is not None:
            raise NotImplementedError("index remapping is not used by any reference config")
        if not use_extrapolation_mask:
            raise NotImplementedError("every reference config sets use_extrapolation_mask=True")
        self.ddconfig = _plain(ddconfig)
        self._names = []
        gen = torch.Generator().manual_seed(torch.initial_seed() & 0x7fffffff)
        for name, shape in arch.param_shapes(self.ddconfig, self.n_embed, self.embed_dim).items():
            self._register(name, self._init_tensor(name, shape, gen))
        self._engine = None
        self._kinv_cache = {}
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    # ----------------------------------------------------------------------------------- parameters
    def _init_tensor(self, name, shape, gen):
        """Same distributions as torch's defaults for the reference's modules (Conv2d, GroupNorm, the
        U(+-1/n_e) embedding of quantize.py:233)."""
        if name == "quantize.embedding.weight":
            return (torch.rand(shape, generator=gen) * 2 - 1) / self.n_embed
        if ".norm" in name:
            return torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        fan_in = None
        if name.endswith(".weight"):
            fan_in = int(np.prod(shape[1:]))
            self._last_fan_in = fan_in
        bound = 1.0 / math.sqrt(fan_in if fan_in else self._last_fan_in)
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound

    def _register(self, dotted, tensor):
        """Register `tensor` under the checkpoint's dotted name by growing a tree of bare containers, so
        state_dict()/load_state_dict() speak the reference layout."""
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if not hasattr(node, part):
                node.add_module(part, torch.nn.Module())
            node = getattr(node, part)
        node.register_parameter(parts[-1], torch.nn.Parameter(tensor, requires_grad=False))
        self._names.append(dotted)

    def _apply(self, fn, *a, **k):
        self._engine = None            # .to() / .cuda() / .float(): weights moved, repack lazily
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._engine = None
        self._graphs = {}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def init_from_ckpt(self, path, ignore_keys=['loss'], only_keep_keys=[]):
        """model.py:87-104: prefix-filter the Lightning checkpoint, then load non-strictly."""
        sd = torch.load(path, map_location="cpu")["state_dict"]
        sd = {k: v for k, v in sd.items() if not any(k.startswith(ik) for ik in ignore_keys)}
        if only_keep_keys:
            sd = {k: v for k, v in sd.items() if all(ik in k for ik in only_keep_keys)}
        own = set(self._names)
        self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        print(f"Restored from {path}")

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def engine(self):
        if self._engine is None:
            if self.device.type != "cuda":
                raise RuntimeError("VQModel: move the model to a CUDA device first (.to('cuda:0')); "
                                   "the B200 engine has no CPU fallback")
            self._engine = VQGANEngine(self.state_dict(), self.ddconfig, self.device, mode=self.engine_mode)
        return self._engine

    def use_vq(self):
        return self.global_step >= self.vq_step_threshold

    # ----------------------------------------------------------------------------------- get_x
    def get_input(self, key, batch):
        """model.py:169-177 (NHWC -> NCHW)."""
        x = batch[key]
        if len(x.shape) == 3:
            x = x[..., None]
        if len(x.shape) == 4:
            x = x.permute(0, 3, 1, 2).to(memory_format=torch.contiguous_format)
        elif len(x.shape) == 5:
            x = x.permute(0, 1, 4, 2, 3).to(memory_format=torch.contiguous_format)
        return x

    def _kinv(self, Ks):
        """fp32 LAPACK inverse on the host, exactly what `src_intrinsics.inverse()` gives the reference on CPU
        (warp.py:212); cached per distinct K because the scene loop re-uses one intrinsic matrix."""
        Kh = Ks.detach().to("cpu", torch.float32).contiguous()
        key = Kh.numpy().tobytes()
        hit = self._kinv_cache.get(key)
        if hit is None or hit.device != self.device:
            if len(self._kinv_cache) > 64:
                self._kinv_cache.clear()
            hit = ops.h2d(Kh.reshape(-1, 3, 3).inverse().reshape(Kh.shape), self.device)
            self._kinv_cache[key] = hit
        return hit

    def get_x(self, batch, dataset, return_extrapolation_mask=False, no_depth_range=False, parallel=True):
        """model.py:179-269.  Returns (x, x_dst) or (x, x_dst, extrapolation_mask, warped_depth).
        `parallel` is accepted for signature parity: the kernel is deterministic (last writer in reference order)."""
        dev = self.device
        if not no_depth_range and self.depth_range is not None:
            raise NotImplementedError("depth_range clipping (warp.py:280-283) is a training-time option; the "
                                      "inference pipeline calls get_x(no_depth_range=True)")
        f32 = lambda t: (torch.as_tensor(t) if not torch.is_tensor(t) else t).to(dev, torch.float32, non_blocking=True)
        if 'warped_tgt_features' in batch:                                     # model.py:196-199
            rgb = f32(batch['warped_tgt_features']).contiguous()
            depth = f32(batch['warped_tgt_depth']).contiguous()
            x, mask = ops.depth_code(rgb, depth, dataset)
        else:
            src = f32(batch["src_imgs"])                                        # [B,N,H,W,3] as prepare_batch_data builds it
            dm = f32(batch["src_depths"])
            if dm.dim() == 5:                                                   # pipeline adds a trailing 1 (:870)
                dm = dm[..., 0]
            B, N = dm.shape[:2]
            Ks = batch["Ks"]
            Ks_host = Ks if (torch.is_tensor(Ks) and not Ks.is_cuda) else torch.as_tensor(Ks).cpu()
            Kinv = self._kinv(Ks_host)
            K_tgt = ops.h2d(Ks_host[:, 0], dev, torch.float32)
            R = torch.as_tensor(batch["R_rels"]).detach().to("cpu", torch.float32)
            t = torch.as_tensor(batch["t_rels"]).detach().to("cpu", torch.float32)
            T = torch.eye(4).repeat(B, N, 1, 1)                                  # model.py:188-195
            T[..., :3, :3] = R
            T[..., :3, 3] = t
            channels_last = src.shape[-1] == 3 and src.dim() == 5 and src.shape[2] != 3
            out = ops.splat_forward(src.contiguous(), dm.contiguous(), K_tgt, Kinv, ops.h2d(T, dev), dataset,
                                    channels_last=channels_last, policy=self.splat_policy)
            x, mask = out["x"], out["mask"]
        extrapolation_mask = mask.view(torch.bool)
        # model.py:238: the coded target frame.  The scene loop feeds all-zero placeholders (inference_pipeline.py:596-597)
        # and never reads x_dst; its batches carry "_dst_placeholder" so that the per-frame path stays free of ATen work.
        x_dst = None
        if "dst_img" in batch and "dst_depth" in batch and not batch.get("_dst_placeholder", False):
            x_dst = self._x_dst(batch, dataset)
        if not return_extrapolation_mask:
            return x, x_dst
        return x, x_dst, extrapolation_mask, x[:, 3:4]

    def _x_dst(self, batch, dataset):
        """model.py:211-213 / 221-223, 238 (target RGB-D coding; unused by the inference pipeline)."""
        dev = self.device
        x_dst = self.get_input("dst_img", {"dst_img": torch.as_tensor(batch["dst_img"]).to(dev, torch.float32)})
        d = self.get_input("dst_depth", {"dst_depth": torch.as_tensor(batch["dst_depth"]).to(dev, torch.float32)})
        if dataset == 'google_earth':
            inv = (1 / (d + 10) - 1 / 14.765625) / (1 / 10.099975586 - 1 / 14.765625)
        elif dataset == 'clevr-infinite':
            inv = (1 / d - 1 / 16) / (1 / 7 - 1 / 16)
        else:
            raise NotImplementedError
        return torch.cat([x_dst, 2 * inv - 1], 1)

    # ----------------------------------------------------------------------------------- encode / decode
    @staticmethod
    def _mask_u8(mask):
        if mask is None:
            return None
        if mask.dtype == torch.bool:
            return mask.contiguous().view(torch.uint8)
        return (mask != 0).to(torch.uint8).contiguous()

    @torch.no_grad()
    def encode(self, x, topk=None, encoding_indices=None, extrapolation_mask=None, use_old=False, sample_number=1):
        """model.py:106-124 -> (quant, emb_loss, info, pre_quantized_f); NCHW views of NHWC device buffers."""
        eng = self.engine
        pre = eng.encode(x.contiguous(), self._mask_u8(extrapolation_mask))     # NHWC [B,h,w,D]
        pre_nchw = pre.permute(0, 3, 1, 2)
        if not self.use_vq():
            return pre_nchw
        if topk is not None and (topk != 1 or sample_number != 1):
            # quantize.py:352-367 on the device: seeded counter-based sampler (parity with torch's global generator is
            # distributional, not bitwise); the reference's row-0-probabilities quirk is kept unless switched off
            B, h, w, D = pre.shape
            self._sample_calls = getattr(self, "_sample_calls", 0) + 1
            seed = (torch.initial_seed() * 0x9E3779B1 + self._sample_calls) & 0xFFFFFFFFFFFFFFFF
            out = ops.vq_topk_sample(pre.view(B * h * w, D), eng.p["quantize.embedding.weight"], int(topk), int(sample_number),
                                     (h, w), mask=self._mask_u8(extrapolation_mask), seed=seed,
                                     row0_probs=getattr(self, "sample_row0_probs", True))
            quants = out["z_q"].view(B, h, w, sample_number, D).permute(0, 3, 4, 1, 2)            # [B,S,D,h,w]
            idx = out["idx"].view(B, h, w, sample_number).permute(0, 3, 1, 2)
            return quants, None, (None, None, idx), pre_nchw
        if encoding_indices is not None:
            idx = encoding_indices.reshape(pre.shape[:3]).to(torch.int64)
            z_q = eng.embed_code(idx)
        else:
            idx, z_q = eng.quantize(pre)
        self._last_zq_nhwc = z_q
        if topk is None:                                                        # VectorQuantizer2.forward
            diff = z_q - pre
            emb_loss = (1.0 + 0.25) * torch.mean(diff * diff)                   # legacy loss, beta=0.25 (quantize.py:300-301)
            return z_q.permute(0, 3, 1, 2), emb_loss, (None, None, idx), pre_nchw
        quants = z_q.permute(0, 3, 1, 2).unsqueeze(1)                          # [B, S=1, D, h, w] (quantize.py:368-369)
        return quants, None, (None, None, idx.unsqueeze(1)), pre_nchw

    @torch.no_grad()
    def decode(self, quant):
        """model.py:131-134: quant [B,D,h,w] -> dec [B,4,H,W]."""
        return self.engine.decode(quant.permute(0, 2, 3, 1).contiguous())

    @torch.no_grad()
    def decode_code(self, code_b):
        """model.py:136-139 (the reference calls a method VectorQuantizer2 does not have; this is the intent)."""
        return self.engine.decode(self.engine.embed_code(code_b.to(torch.int64)))

    def _graphed_forward(self, x, mask_u8):
        """encode -> quantize -> decode as one CUDA-graph replay.  -> (dec NCHW, pre NHWC, z_q NHWC, idx), fresh tensors."""
        eng = self.engine
        key = (tuple(x.shape), mask_u8 is not None, x.device.index)
        g = self._graphs.get(key)
        if g is None:
            sx = x.clone()
            sm = mask_u8.clone() if mask_u8 is not None else None
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eng.forward(sx, sm)                                   # warm-up: allocator, one-time kernel attributes
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = eng.forward(sx, sm)
            g = self._graphs[key] = (graph, sx, sm, outs)
        graph, sx, sm, outs = g
        sx.copy_(x, non_blocking=True)
        if sm is not None:
            sm.copy_(mask_u8, non_blocking=True)
        graph.replay()
        return tuple(o.clone() for o in outs)

    @torch.no_grad()
    def forward(self, input, topk=None, extrapolation_mask=None, sample_number=1, get_codebook_count=False,
                get_pre_quantized_feature=False, get_quantized_feature=False):
        """model.py:141-167."""
        if self.use_cuda_graph and self.use_vq() and topk in (None, 1) and sample_number == 1 and input.is_cuda:
            dec, pre, z_q, idx = self._graphed_forward(input.contiguous(), self._mask_u8(extrapolation_mask))
            pre_nchw = pre.permute(0, 3, 1, 2)
            if topk is None:
                diff = z_q - pre
                out = [dec, (1.0 + 0.25) * torch.mean(diff * diff)]
                info, quants = (None, None, idx), z_q.permute(0, 3, 1, 2)
            else:
                out = [[dec[None,]], None]
                info, quants = (None, None, idx.unsqueeze(1)), z_q.permute(0, 3, 1, 2).unsqueeze(1)
            if get_codebook_count:
                out.append(info[-1])
            if get_pre_quantized_feature:
                out.append(pre_nchw)
            if get_quantized_feature:
                out.append(quants)
            return out
        res = self.encode(input, topk=topk, encoding_indices=None, extrapolation_mask=extrapolation_mask,
                          sample_number=sample_number)
        if not self.use_vq():
            pre_quant = res
            dec = self.decode(pre_quant)
            return dec, torch.tensor(0).to(dec.device), pre_quant
        quants, diff, info, pre_quant = res
        if topk is None:
            decs = self.decode(quants)
        else:
            decs = [self.decode(quants[:, i])[None,] for i in range(sample_number)]
        out = [decs, diff]
        if get_codebook_count:
            out.append(info[-1] if len(info) else {})
        if get_pre_quantized_feature:
            out.append(pre_quant)
        if get_quantized_feature:
            out.append(quants)
        return out
