"""Host-side executor of the conditional VQGAN (VQModel.encode / quantize / decode) on libsgam_b200 kernels.

Mirrors the module structure of the reference checkpoint -- encoder.down.{l}.block.{b},
decoder.up.{l}.attn.{b}, ... -- but holds no nn.Module: weights are repacked once (OIHW -> [Cout, kh*kw*Cin],
K-major for NHWC implicit GEMM) and every layer is one or two C-ABI calls on the current CUDA stream.

Reference: sgam/generative_sensing_module/modules/diffusionmodules/model.py:29-192 (blocks), :342-433 (Encoder),
:437-539 (Decoder); sgam/generative_sensing_module/model.py:106-139 (encode / decode).
"""
import torch

from . import ops

HOT_PATH_PREFIXES = ("conv_in.", "encoder.", "decoder.", "quantize.embedding.", "quant_conv.", "post_quant_conv.")


def _pack_conv(w):
    """OIHW -> [Cout, kh*kw*Cin] with K index (kh*k + kw)*Cin + ci."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


class VQGANEngine:
    def __init__(self, state_dict, ddconfig, device="cuda:0"):
        self.dd = dict(ddconfig)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VQGANEngine runs on a CUDA device only (no CPU fallback)")
        self.p = {}
        self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        """Take the reference's checkpoint layout; tolerate / ignore loss.* and perceptual_loss.* keys
        (model.py:87-104 loads with strict=False)."""
        p = {}
        for k, v in sd.items():
            if not k.startswith(HOT_PATH_PREFIXES):
                continue
            v = v.detach().to(device=self.device, dtype=torch.float32)
            if v.dim() == 4:
                v = _pack_conv(v)
            p[k] = v.contiguous()
        missing = [k for k in ("conv_in.weight", "encoder.conv_in.weight", "decoder.conv_out.weight",
                               "quantize.embedding.weight", "quant_conv.weight", "post_quant_conv.weight") if k not in p]
        if missing:
            raise KeyError(f"state_dict lacks hot-path tensors: {missing}")
        self.p = p
        self.n_embed, self.embed_dim = p["quantize.embedding.weight"].shape

    def has(self, name):
        return f"{name}.weight" in self.p

    # ------------------------------------------------------------------ blocks (NHWC)
    def conv(self, name, x, **kw):
        return ops.conv2d(x, self.p[f"{name}.weight"], self.p[f"{name}.bias"], **kw)

    def norm(self, name, x, swish):
        return ops.groupnorm(x, self.p[f"{name}.weight"], self.p[f"{name}.bias"], swish)

    def resnet_block(self, name, x):
        """model.py:78-137 with temb=None, dropout 0."""
        h = self.conv(f"{name}.conv1", self.norm(f"{name}.norm1", x, True), ksize=3)
        h = self.norm(f"{name}.norm2", h, True)
        if self.has(f"{name}.nin_shortcut"):
            x = self.conv(f"{name}.nin_shortcut", x, ksize=1)
        return self.conv(f"{name}.conv2", h, ksize=3, residual=x)

    def attn_block(self, name, x):
        """model.py:140-192: single-head spatial self-attention over H*W tokens."""
        B, H, W, C = x.shape
        h_ = self.norm(f"{name}.norm", x, False)
        q = self.conv(f"{name}.q", h_, ksize=1).view(B, H * W, C)
        k = self.conv(f"{name}.k", h_, ksize=1).view(B, H * W, C)
        # V^T [B, C, tokens] = W_v . h^T + b_v (per row), so that P.V is another A.B^T product
        vT = ops.gemm_nt(self.p[f"{name}.v.weight"], h_.view(B, H * W, C), bias_m=self.p[f"{name}.v.bias"])
        s = ops.gemm_nt(q, k, alpha=float(int(C) ** (-0.5)))          # [B, tokens, tokens]
        ops.softmax_rows_(s)
        o = ops.gemm_nt(s, vT).view(B, H, W, C)
        return self.conv(f"{name}.proj_out", o, ksize=1, residual=x)

    # ------------------------------------------------------------------ encoder / decoder
    def encoder(self, h):
        dd = self.dd
        nres, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
        h = self.conv("encoder.conv_in", h, ksize=3)
        for l in range(nres):
            for b in range(nrb):
                h = self.resnet_block(f"encoder.down.{l}.block.{b}", h)
                if self.has(f"encoder.down.{l}.attn.{b}.norm"):
                    h = self.attn_block(f"encoder.down.{l}.attn.{b}", h)
            if l != nres - 1:
                h = self.conv(f"encoder.down.{l}.downsample.conv", h, ksize=3, stride=2, pad_mode=1)
        h = self.resnet_block("encoder.mid.block_1", h)
        h = self.attn_block("encoder.mid.attn_1", h)
        h = self.resnet_block("encoder.mid.block_2", h)
        return self.conv("encoder.conv_out", self.norm("encoder.norm_out", h, True), ksize=3)

    def decoder(self, z):
        dd = self.dd
        nres, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
        h = self.conv("decoder.conv_in", z, ksize=3)
        h = self.resnet_block("decoder.mid.block_1", h)
        h = self.attn_block("decoder.mid.attn_1", h)
        h = self.resnet_block("decoder.mid.block_2", h)
        for l in reversed(range(nres)):
            for b in range(nrb + 1):
                h = self.resnet_block(f"decoder.up.{l}.block.{b}", h)
                if self.has(f"decoder.up.{l}.attn.{b}.norm"):
                    h = self.attn_block(f"decoder.up.{l}.attn.{b}", h)
            if l != 0:
                h = self.conv(f"decoder.up.{l}.upsample.conv", h, ksize=3, upsample=1)
        h = self.norm("decoder.norm_out", h, True)
        return self.conv("decoder.conv_out", h, ksize=3, out_nchw=True)        # [B, out_ch, H, W]

    # ------------------------------------------------------------------ VQModel pieces
    def encode(self, x, mask=None):
        """model.py:106-116: x [B,4,H,W] NCHW (+ mask [B,1,H,W] uint8) -> pre-quantised latent NHWC [B,h,w,D]."""
        if mask is not None:
            mask = mask.reshape(mask.shape[0], *mask.shape[-2:])
        h = ops.stem_conv(x, mask, self.p["conv_in.weight"], self.p["conv_in.bias"])
        h = self.encoder(h)
        return self.conv("quant_conv", h, ksize=1)

    def quantize(self, pre_quant):
        """quantize.py:275-319 / :344-381 (topk=1): NHWC latent -> idx [B,h,w] int64, z_q NHWC."""
        B, h, w, D = pre_quant.shape
        idx, z_q = ops.vq_nearest(pre_quant.view(B * h * w, D), self.p["quantize.embedding.weight"])
        return idx.view(B, h, w), z_q.view(B, h, w, D)

    def decode(self, z_q):
        """model.py:131-134: NHWC quantised latent -> dec [B,4,H,W] NCHW."""
        return self.decoder(self.conv("post_quant_conv", z_q, ksize=1))

    def embed_code(self, idx):
        """Codebook gather for decode_code (model.py:136-139): idx [B,h,w] -> NHWC latent."""
        return torch.index_select(self.p["quantize.embedding.weight"], 0, idx.reshape(-1)).view(*idx.shape, -1)

    @torch.no_grad()
    def forward(self, x, mask=None):
        """-> dec [B,4,H,W], pre_quant NHWC, z_q NHWC, idx [B,h,w]."""
        pre = self.encode(x, mask)
        idx, z_q = self.quantize(pre)
        return self.decode(z_q), pre, z_q, idx
