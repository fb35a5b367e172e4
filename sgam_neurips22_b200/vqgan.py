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
    """mode "tc"  : convolutions and attention products on tcgen05 tensor cores with split-bf16 operands
                     (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM); odd-shaped layers stay on the fp32 kernels.
       mode "simt": every layer on the exact-fp32 CUDA-core kernels (the on-device reference of the tc path)."""

    def __init__(self, state_dict, ddconfig, device="cuda:0", mode="tc", nsplit=3):
        self.dd = dict(ddconfig)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VQGANEngine runs on a CUDA device only (no CPU fallback)")
        if mode not in ("tc", "simt"):
            raise ValueError(mode)
        self.mode, self.nsplit = mode, nsplit
        import os
        self.fused_head = os.environ.get("SGAM_FUSED_HEAD", "1") != "0"
        self.fused_qkv = os.environ.get("SGAM_FUSED_QKV", "1") != "0"
        self.fused_gnconv = os.environ.get("SGAM_FUSED_GNCONV", "0") == "1"   # GroupNorm apply inside the 128-channel convs: correct, slower (opt-in)
        self.emit_split = os.environ.get("SGAM_EMIT_SPLIT", "1") != "0"    # producers of Downsample / Upsample inputs write split bf16
        self.subpixel = os.environ.get("SGAM_SUBPIXEL", "1") != "0"
        self.fused_stem = os.environ.get("SGAM_FUSED_STEM", "1") != "0"
        self.p = {}
        self.wsplit = {}
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
        self.wsplit = {}
        self.codebook_tc = None
        if self.mode == "tc":
            if self.embed_dim % 64 == 0 and self.n_embed % 128 == 0:
                self.codebook_tc = ops.CodebookTC(p["quantize.embedding.weight"])
            for k, v in p.items():
                if k.endswith(".weight") and v.dim() == 2 and k != "quantize.embedding.weight" and v.shape[1] % 64 == 0:
                    self.wsplit[k[:-len(".weight")]] = ops.split_weight(v, pad_rows_to=32)
            # AttnBlock projections: q | k | v weight rows stacked for the fused projection GEMM (ops.qkv_tc)
            for k in list(p.keys()):
                if k.endswith(".q.weight") and p[k].dim() == 2 and p[k].shape[1] % 128 == 0:
                    base = k[:-len(".q.weight")]
                    self.wsplit[base + ".qkv"] = ops.split_weight(torch.cat([p[base + ".q.weight"], p[base + ".k.weight"], p[base + ".v.weight"]], 0).contiguous())
                    self.p[base + ".qkv.bias"] = torch.cat([p[base + ".q.bias"], p[base + ".k.bias"], p[base + ".v.bias"]], 0).contiguous()
            # Upsample convs (nearest x2 + 3x3) also in sub-pixel form: four 2x2 parity filters on the low-resolution tensor
            for k, v in p.items():
                if k.endswith(".upsample.conv.weight") and v.dim() == 2 and v.shape[1] % (9 * 64) == 0:
                    self.wsplit[k[:-len(".weight")] + ".subpixel"] = ops.split_weight(ops.subpixel_weights(v, v.shape[1] // 9))
            # encoder.conv_in has Cin = 4: zero-pad each tap's channels to 64 so it runs on the tensor cores as well
            w = p["encoder.conv_in.weight"]
            cin = self.dd["in_channels"]
            if w.shape[1] == 9 * cin and cin <= 64:
                wp = torch.zeros(w.shape[0], 9, 64, device=w.device)
                wp[:, :, :cin] = w.view(w.shape[0], 9, cin)
                self.wsplit["encoder.conv_in.padded"] = ops.split_weight(wp.view(w.shape[0], 9 * 64).contiguous(), pad_rows_to=32)

    def has(self, name):
        return f"{name}.weight" in self.p

    # ------------------------------------------------------------------ blocks (NHWC)
    def conv(self, name, x, **kw):
        """fp32 CUDA-core conv (sgam_conv2d)."""
        return ops.conv2d(x, self.p[f"{name}.weight"], self.p[f"{name}.bias"], **kw)

    def norm(self, name, x, swish):
        return ops.groupnorm(x, self.p[f"{name}.weight"], self.p[f"{name}.bias"], swish)

    def norm_split(self, name, x, swish):
        return ops.groupnorm_split(x, self.p[f"{name}.weight"], self.p[f"{name}.bias"], swish)

    def tc_ok(self, name, x_shape, ksize, stride=1):
        if self.mode != "tc" or name not in self.wsplit:
            return False
        B, H, W, Cin = x_shape
        if H % stride or W % stride:
            return False
        return ops.tc_supported_conv(H // stride, W // stride, Cin, self.p[f"{name}.weight"].shape[0], ksize, stride)

    def unsplit_k(self, name, x_shape, ksize):
        """True when the library runs this conv without a split K loop (under-filled grids split K and reduce into an fp32
        tensor; their output cannot come out of the epilogue as split bf16)."""
        B, H, W, Cin = x_shape
        return ops.conv2d_tc_splitk_floats(B, H, W, Cin, self.p[f"{name}.weight"].shape[0], ksize, 1) == 0

    def conv_tc(self, name, xs, ksize, **kw):
        """tcgen05 conv on a split-bf16 activation pair."""
        kw.setdefault("gn_stats", True)       # nearly every fp32 conv output feeds a GroupNorm: fuse its statistics
        return ops.conv2d_tc(xs, self.wsplit[name], self.p[f"{name}.bias"], ksize=ksize, nsplit=self.nsplit,
                             cout=self.p[f"{name}.weight"].shape[0], **kw)

    def conv_from_f32(self, name, x, ksize, upsample=0, residual=None):
        """conv of an fp32 activation: split (+ fused x2 up-sampling) then tcgen05, or the fp32 kernel."""
        B, H, W, C = x.shape
        if self.tc_ok(name, (B, H << upsample, W << upsample, C), ksize):
            return self.conv_tc(name, ops.split_bf16(x, upsample), ksize, residual=residual)
        return self.conv(name, x, ksize=ksize, upsample=upsample, residual=residual)

    def norm_conv(self, norm_name, conv_name, x, residual=None, want_split=False):
        """GroupNorm + swish + 3x3 conv (the ResnetBlock / norm_out pattern).  want_split: the only consumer is another
        tensor-core conv (Downsample / sub-pixel Upsample): return the (hi, lo) bf16 planes straight from the epilogue instead
        of an fp32 tensor that a separate pass would split (falls back to fp32 where the tensor-core path does not apply)."""
        if self.tc_ok(conv_name, x.shape, 3):
            B, H, W, Cin = x.shape
            if self.fused_gnconv and getattr(x, "gn_partial", None) is not None and \
                    ops.gn_conv2d_tc_supported(B, H, W, Cin, self.p[f"{conv_name}.weight"].shape[0]):
                # 128-channel layers on wide images: GroupNorm + swish + split run inside the conv's operand path
                return ops.gn_conv2d_tc(x, self.p[f"{norm_name}.weight"], self.p[f"{norm_name}.bias"], self.wsplit[conv_name],
                                        self.p[f"{conv_name}.bias"], residual=residual, out_f32=not want_split, out_split=want_split,
                                        nsplit=self.nsplit)
            if want_split and self.unsplit_k(conv_name, x.shape, 3):
                return self.conv_tc(conv_name, self.norm_split(norm_name, x, True), 3, residual=residual, out_f32=False, out_split=True)
            return self.conv_tc(conv_name, self.norm_split(norm_name, x, True), 3, residual=residual)
        return self.conv(conv_name, self.norm(norm_name, x, True), ksize=3, residual=residual)

    def resnet_block(self, name, x, want_split=False):
        """model.py:78-137 with temb=None, dropout 0."""
        h = self.norm_conv(f"{name}.norm1", f"{name}.conv1", x)
        if self.has(f"{name}.nin_shortcut"):
            x = self.conv_from_f32(f"{name}.nin_shortcut", x, 1)
        return self.norm_conv(f"{name}.norm2", f"{name}.conv2", h, residual=x, want_split=want_split)

    def attn_block(self, name, x, want_split=False):
        """model.py:140-192: single-head spatial self-attention over H*W tokens."""
        B, H, W, C = x.shape
        T = H * W
        scale = float(int(C) ** (-0.5))
        if self.mode == "tc" and f"{name}.q" in self.wsplit and T % 8 == 0 and T % 32 == 0 and self.tc_ok(f"{name}.q", x.shape, 1):
            hs = self.norm_split(f"{name}.norm", x, False)
            flat = lambda pair, shape: (pair[0].view(shape), pair[1].view(shape))
            if self.fused_qkv and f"{name}.qkv" in self.wsplit and ops.qkv_tc_supported(B, H, W, C):
                # one GEMM for the three projections; q / k come back as the column halves of one tensor, V already transposed
                q, k, vT = ops.qkv_tc(hs, self.wsplit[f"{name}.qkv"], self.p[f"{name}.qkv.bias"], nsplit=self.nsplit)
            else:
                q = flat(self.conv_tc(f"{name}.q", hs, 1, out_f32=False, out_split=True), (B, T, C))
                k = flat(self.conv_tc(f"{name}.k", hs, 1, out_f32=False, out_split=True), (B, T, C))
                # V^T [B, C, T] = W_v . h^T + b_v (bias per row), so that P.V is another A.B^T product
                vT = ops.gemm_nt_tc(self.wsplit[f"{name}.v"], flat(hs, (B, T, C)), bias_m=self.p[f"{name}.v.bias"],
                                    out_f32=False, out_split=True, nsplit=self.nsplit)
            if self.use_fused_attention(B, T, C):
                # one flash-style kernel: scores / probabilities stay in tensor memory, [B,T,T] is never written
                o = ops.attention_tc(q, k, vT, scale)
            else:
                s = ops.gemm_nt_tc(q, k, alpha=scale, nsplit=self.nsplit)                                       # [B, T, T] fp32
                p = ops.softmax_split(s)
                o = ops.gemm_nt_tc(p, vT, out_f32=False, out_split=True, nsplit=self.nsplit)                    # [B, T, C]
            if want_split and self.unsplit_k(f"{name}.proj_out", x.shape, 1):
                return self.conv_tc(f"{name}.proj_out", flat(o, (B, H, W, C)), 1, residual=x, out_f32=False, out_split=True)
            return self.conv_tc(f"{name}.proj_out", flat(o, (B, H, W, C)), 1, residual=x)
        h_ = self.norm(f"{name}.norm", x, False)
        q = self.conv(f"{name}.q", h_, ksize=1).view(B, T, C)
        k = self.conv(f"{name}.k", h_, ksize=1).view(B, T, C)
        vT = ops.gemm_nt(self.p[f"{name}.v.weight"], h_.view(B, T, C), bias_m=self.p[f"{name}.v.bias"])
        s = ops.gemm_nt(q, k, alpha=scale)
        ops.softmax_rows_(s)
        o = ops.gemm_nt(s, vT).view(B, H, W, C)
        return self.conv(f"{name}.proj_out", o, ksize=1, residual=x)

    def use_fused_attention(self, B, T, C):
        """256-channel blocks whose token count tiles into 256-query pair tiles run the fused kernel at every batch size
        (small batches split the keys of a tile over several SM pairs and merge); SGAM_ATTN=3pass selects the three-pass
        path (QK^T GEMM -> softmax -> PV GEMM), which the 512-channel mid blocks always take."""
        import os
        mode = os.environ.get("SGAM_ATTN", "auto")
        return mode != "3pass" and self.nsplit == 3 and ops.attention_tc_supported(B, T, C)

    # ------------------------------------------------------------------ encoder / decoder
    def encoder(self, h):
        dd = self.dd
        nres, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
        for l in range(nres):                                             # h = encoder.conv_in(stem(x)), see encode()
            down = f"encoder.down.{l}.downsample.conv"
            for b in range(nrb):
                # the level's last tensor feeds only the Downsample conv (diffusionmodules/model.py:418-423: no skip connections):
                # its producer emits the split-bf16 operand directly
                last = self.emit_split and b == nrb - 1 and l != nres - 1 and self.tc_ok(down, h.shape, 3, stride=2)
                has_attn = self.has(f"encoder.down.{l}.attn.{b}.norm")
                h = self.resnet_block(f"encoder.down.{l}.block.{b}", h, want_split=last and not has_attn)
                if has_attn:
                    h = self.attn_block(f"encoder.down.{l}.attn.{b}", h, want_split=last)
            if l != nres - 1:
                if isinstance(h, tuple):
                    h = self.conv_tc(down, h, 3, stride=2)
                elif self.tc_ok(down, h.shape, 3, stride=2):
                    h = self.conv_tc(down, ops.split_bf16(h), 3, stride=2)
                else:
                    h = self.conv(down, h, ksize=3, stride=2, pad_mode=1)
        h = self.resnet_block("encoder.mid.block_1", h)
        h = self.attn_block("encoder.mid.attn_1", h)
        h = self.resnet_block("encoder.mid.block_2", h)
        return self.norm_conv("encoder.norm_out", "encoder.conv_out", h)

    def decoder(self, z):
        dd = self.dd
        nres, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
        h = self.conv_from_f32("decoder.conv_in", z, 3)
        h = self.resnet_block("decoder.mid.block_1", h)
        h = self.attn_block("decoder.mid.attn_1", h)
        h = self.resnet_block("decoder.mid.block_2", h)
        for l in reversed(range(nres)):
            name = f"decoder.up.{l}.upsample.conv"
            for b in range(nrb + 1):
                # the level's last tensor feeds only the Upsample conv (:527-533); in sub-pixel form that conv reads the
                # LOW-resolution split operand, which the producer's epilogue can write directly
                last = False
                if self.emit_split and b == nrb and l != 0 and self.subpixel and f"{name}.subpixel" in self.wsplit:
                    Bh, Hh, Wh, _ = h.shape
                    cmid = self.p[f"decoder.up.{l}.block.{b}.conv2.weight"].shape[0]
                    last = ops.conv2d_tc_up2_supported(Bh, Hh, Wh, cmid, self.p[f"{name}.weight"].shape[0])
                has_attn = self.has(f"decoder.up.{l}.attn.{b}.norm")
                h = self.resnet_block(f"decoder.up.{l}.block.{b}", h, want_split=last and not has_attn)
                if has_attn:
                    h = self.attn_block(f"decoder.up.{l}.attn.{b}", h, want_split=last)
            if l != 0:
                if isinstance(h, tuple):
                    h = ops.conv2d_tc_up2(h, self.wsplit[f"{name}.subpixel"], self.p[f"{name}.bias"], nsplit=self.nsplit)
                    continue
                Bh, Hh, Wh, Ch = h.shape
                if self.subpixel and f"{name}.subpixel" in self.wsplit and \
                        ops.conv2d_tc_up2_supported(Bh, Hh, Wh, Ch, self.p[f"{name}.weight"].shape[0]):
                    h = ops.conv2d_tc_up2(ops.split_bf16(h), self.wsplit[f"{name}.subpixel"], self.p[f"{name}.bias"], nsplit=self.nsplit)
                else:
                    h = self.conv_from_f32(name, h, 3, upsample=1)
        w_out = self.p["decoder.conv_out.weight"]
        if self.mode == "tc" and self.fused_head and getattr(h, "gn_partial", None) is not None and h.shape[-1] == 128 \
                and w_out.shape == (4, 9 * 128):
            # GroupNorm + swish + the 4-channel conv as one fp32 kernel (a 4-column GEMM tile wastes the tensor pipe)
            if "decoder.conv_out.taps" not in self.p:
                self.p["decoder.conv_out.taps"] = w_out.view(4, 9, 128).permute(1, 2, 0).contiguous()        # [tap, cin, cout]
            return ops.gn_head_conv(h, self.p["decoder.norm_out.weight"], self.p["decoder.norm_out.bias"],
                                    self.p["decoder.conv_out.taps"], self.p["decoder.conv_out.bias"])
        if self.tc_ok("decoder.conv_out", h.shape, 3):                          # Cout = 4: zero-padded to one 32-column tile
            return self.conv_tc("decoder.conv_out", self.norm_split("decoder.norm_out", h, True), 3, out_nchw=True)
        h = self.norm("decoder.norm_out", h, True)
        return self.conv("decoder.conv_out", h, ksize=3, out_nchw=True)        # [B, out_ch, H, W]

    # ------------------------------------------------------------------ VQModel pieces
    def encode(self, x, mask=None):
        """model.py:106-116: x [B,4,H,W] NCHW (+ mask [B,1,H,W] uint8) -> pre-quantised latent NHWC [B,h,w,D]."""
        if mask is not None:
            mask = mask.reshape(mask.shape[0], *mask.shape[-2:])
        B, _, H, W = x.shape
        w3 = self.p["encoder.conv_in.weight"]
        if self.mode == "tc" and self.fused_stem and w3.shape == (128, 36) and W >= 4 and ops.tc_supported_conv(H, W, 64, 128, 3, 1):
            # stem + conv_in + GroupNorm statistics as one fp32 kernel (36 real taps are FP32-pipe work, not a padded GEMM)
            h = ops.stem_conv_in(x, mask, self.p["conv_in.weight"], self.p["conv_in.bias"], w3, self.p["encoder.conv_in.bias"])
        elif self.mode == "tc" and "encoder.conv_in.padded" in self.wsplit and \
                ops.tc_supported_conv(H, W, 64, self.p["encoder.conv_in.weight"].shape[0], 3, 1):
            xs = ops.stem_conv_split(x, mask, self.p["conv_in.weight"], self.p["conv_in.bias"], 64)
            h = ops.conv2d_tc(xs, self.wsplit["encoder.conv_in.padded"], self.p["encoder.conv_in.bias"], ksize=3,
                              nsplit=self.nsplit, gn_stats=True)
        else:
            h = self.conv("encoder.conv_in", ops.stem_conv(x, mask, self.p["conv_in.weight"], self.p["conv_in.bias"]), ksize=3)
        h = self.encoder(h)
        return self.conv_from_f32("quant_conv", h, 1)

    def quantize(self, pre_quant):
        """quantize.py:275-319 / :344-381 (topk=1): NHWC latent -> idx [B,h,w] int64, z_q NHWC."""
        B, h, w, D = pre_quant.shape
        if self.codebook_tc is not None:
            idx, z_q = ops.vq_nearest_tc(pre_quant.view(B * h * w, D), self.codebook_tc)
        else:
            idx, z_q = ops.vq_nearest(pre_quant.view(B * h * w, D), self.p["quantize.embedding.weight"])
        return idx.view(B, h, w), z_q.view(B, h, w, D)

    def decode(self, z_q):
        """model.py:131-134: NHWC quantised latent -> dec [B,4,H,W] NCHW."""
        return self.decoder(self.conv_from_f32("post_quant_conv", z_q, 1))

    def embed_code(self, idx):
        """Codebook gather for decode_code (model.py:136-139): idx [B,h,w] -> NHWC latent."""
        return torch.index_select(self.p["quantize.embedding.weight"], 0, idx.reshape(-1)).view(*idx.shape, -1)

    @torch.no_grad()
    def forward(self, x, mask=None):
        """-> dec [B,4,H,W], pre_quant NHWC, z_q NHWC, idx [B,h,w]."""
        pre = self.encode(x, mask)
        idx, z_q = self.quantize(pre)
        return self.decode(z_q), pre, z_q, idx
