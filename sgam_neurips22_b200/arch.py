"""Parameter inventory of the conditional VQGAN in the reference's checkpoint layout.

Walks the same module tree as sgam/generative_sensing_module/modules/diffusionmodules/model.py:342-539 (Encoder,
Decoder) and sgam/generative_sensing_module/model.py:54-63 (conv_in, quantize, quant_conv, post_quant_conv) and
yields `name -> shape`; used to create / validate weights without instantiating torch modules.
"""
from collections import OrderedDict


def _res(p, n, cin, cout):
    p[f"{n}.norm1.weight"] = p[f"{n}.norm1.bias"] = (cin,)
    p[f"{n}.conv1.weight"], p[f"{n}.conv1.bias"] = (cout, cin, 3, 3), (cout,)
    p[f"{n}.norm2.weight"] = p[f"{n}.norm2.bias"] = (cout,)
    p[f"{n}.conv2.weight"], p[f"{n}.conv2.bias"] = (cout, cout, 3, 3), (cout,)
    if cin != cout:
        p[f"{n}.nin_shortcut.weight"], p[f"{n}.nin_shortcut.bias"] = (cout, cin, 1, 1), (cout,)


def _att(p, n, c):
    p[f"{n}.norm.weight"] = p[f"{n}.norm.bias"] = (c,)
    for leaf in ("q", "k", "v", "proj_out"):
        p[f"{n}.{leaf}.weight"], p[f"{n}.{leaf}.bias"] = (c, c, 1, 1), (c,)


def param_shapes(ddconfig, n_embed, embed_dim, use_extrapolation_mask=True):
    dd = ddconfig
    ch, mult, nrb = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"]
    attn_res = list(dd["attn_resolutions"])
    nres = len(mult)
    p = OrderedDict()
    if use_extrapolation_mask:
        p["conv_in.weight"], p["conv_in.bias"] = (4, 5, 1, 1), (4,)
    p["encoder.conv_in.weight"], p["encoder.conv_in.bias"] = (ch, dd["in_channels"], 3, 3), (ch,)
    res = dd["resolution"]
    widths = [ch] + [ch * m for m in mult]
    for l in range(nres):
        cin, cout = widths[l], widths[l + 1]
        for b in range(nrb):
            _res(p, f"encoder.down.{l}.block.{b}", cin, cout)
            cin = cout
            if res in attn_res:
                _att(p, f"encoder.down.{l}.attn.{b}", cin)
        if l != nres - 1:
            p[f"encoder.down.{l}.downsample.conv.weight"], p[f"encoder.down.{l}.downsample.conv.bias"] = (cin, cin, 3, 3), (cin,)
            res //= 2
    _res(p, "encoder.mid.block_1", cin, cin)
    _att(p, "encoder.mid.attn_1", cin)
    _res(p, "encoder.mid.block_2", cin, cin)
    p["encoder.norm_out.weight"] = p["encoder.norm_out.bias"] = (cin,)
    zc = dd["z_channels"] * (2 if dd.get("double_z", True) else 1)
    p["encoder.conv_out.weight"], p["encoder.conv_out.bias"] = (zc, cin, 3, 3), (zc,)

    cin = ch * mult[-1]
    res = dd["resolution"] // 2 ** (nres - 1)
    p["decoder.conv_in.weight"], p["decoder.conv_in.bias"] = (cin, dd["z_channels"], 3, 3), (cin,)
    _res(p, "decoder.mid.block_1", cin, cin)
    _att(p, "decoder.mid.attn_1", cin)
    _res(p, "decoder.mid.block_2", cin, cin)
    for l in reversed(range(nres)):
        cout = ch * mult[l]
        for b in range(nrb + 1):
            _res(p, f"decoder.up.{l}.block.{b}", cin, cout)
            cin = cout
            if res in attn_res:
                _att(p, f"decoder.up.{l}.attn.{b}", cin)
        if l != 0:
            p[f"decoder.up.{l}.upsample.conv.weight"], p[f"decoder.up.{l}.upsample.conv.bias"] = (cin, cin, 3, 3), (cin,)
            res *= 2
    p["decoder.norm_out.weight"] = p["decoder.norm_out.bias"] = (cin,)
    p["decoder.conv_out.weight"], p["decoder.conv_out.bias"] = (dd["out_ch"], cin, 3, 3), (dd["out_ch"],)

    p["quantize.embedding.weight"] = (n_embed, embed_dim)
    p["quant_conv.weight"], p["quant_conv.bias"] = (embed_dim, dd["z_channels"], 1, 1), (embed_dim,)
    p["post_quant_conv.weight"], p["post_quant_conv.bias"] = (dd["z_channels"], embed_dim, 1, 1), (dd["z_channels"],)
    return p
