"""torch-CPU fp32 restatement of the VQGAN Encoder / Decoder.  TEST INFRASTRUCTURE ONLY.

Functional (no nn.Module): every function takes the reference's checkpoint
`state_dict` and a key prefix, and applies the same ATen
operators, in the same order, as
/root/reference/sgam/generative_sensing_module/modules/diffusionmodules/model.py.
"""
import torch
import torch.nn.functional as F

from .recipes import DDCONFIG


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[f"{name}.weight"], sd[f"{name}.bias"], stride=stride, padding=padding)


def normalize(sd, name, x):
    """diffusionmodules/model.py:34-35  GroupNorm(32, C, eps=1e-6, affine=True)"""
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps=1e-6)


def nonlinearity(x):
    """diffusionmodules/model.py:29-31  swish"""
    return x * torch.sigmoid(x)


def resnet_block(sd, name, x):
    """diffusionmodules/model.py:78-137 (temb=None, dropout p=0)"""
    h = _conv(sd, f"{name}.conv1", nonlinearity(normalize(sd, f"{name}.norm1", x)), padding=1)
    h = _conv(sd, f"{name}.conv2", nonlinearity(normalize(sd, f"{name}.norm2", h)), padding=1)
    if f"{name}.nin_shortcut.weight" in sd:
        x = _conv(sd, f"{name}.nin_shortcut", x)
    return x + h


def attn_block(sd, name, x):
    """diffusionmodules/model.py:140-192 single-head spatial self-attention"""
    h_ = normalize(sd, f"{name}.norm", x)
    q, k, v = (_conv(sd, f"{name}.{n}", h_) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, f"{name}.proj_out", h_)


def downsample(sd, name, x):
    """diffusionmodules/model.py:56-75: pad right/bottom by one, 3x3 stride 2"""
    return _conv(sd, f"{name}.conv", F.pad(x, (0, 1, 0, 1), mode="constant", value=0), stride=2)


def upsample(sd, name, x):
    """diffusionmodules/model.py:38-53: nearest x2, 3x3"""
    return _conv(sd, f"{name}.conv", F.interpolate(x, scale_factor=2.0, mode="nearest"), padding=1)


def encoder(sd, x, dd=DDCONFIG, prefix="encoder"):
    """diffusionmodules/model.py:405-433"""
    nres, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
    h = _conv(sd, f"{prefix}.conv_in", x, padding=1)
    for l in range(nres):
        for b in range(nrb):
            h = resnet_block(sd, f"{prefix}.down.{l}.block.{b}", h)
            if f"{prefix}.down.{l}.attn.{b}.norm.weight" in sd:
                h = attn_block(sd, f"{prefix}.down.{l}.attn.{b}", h)
        if l != nres - 1:
            h = downsample(sd, f"{prefix}.down.{l}.downsample", h)
    h = resnet_block(sd, f"{prefix}.mid.block_1", h)
    h = attn_block(sd, f"{prefix}.mid.attn_1", h)
    h = resnet_block(sd, f"{prefix}.mid.block_2", h)
    h = nonlinearity(normalize(sd, f"{prefix}.norm_out", h))
    return _conv(sd, f"{prefix}.conv_out", h, padding=1)


def decoder(sd, z, dd=DDCONFIG, prefix="decoder"):
    """diffusionmodules/model.py:508-539"""
    nres, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
    h = _conv(sd, f"{prefix}.conv_in", z, padding=1)
    h = resnet_block(sd, f"{prefix}.mid.block_1", h)
    h = attn_block(sd, f"{prefix}.mid.attn_1", h)
    h = resnet_block(sd, f"{prefix}.mid.block_2", h)
    for l in reversed(range(nres)):
        for b in range(nrb + 1):
            h = resnet_block(sd, f"{prefix}.up.{l}.block.{b}", h)
            if f"{prefix}.up.{l}.attn.{b}.norm.weight" in sd:
                h = attn_block(sd, f"{prefix}.up.{l}.attn.{b}", h)
        if l != 0:
            h = upsample(sd, f"{prefix}.up.{l}.upsample", h)
    h = nonlinearity(normalize(sd, f"{prefix}.norm_out", h))
    return _conv(sd, f"{prefix}.conv_out", h, padding=1)
