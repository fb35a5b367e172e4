"""Seeded synthetic inputs and weights.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Shared by tests/golden/make_golden.py (which feeds them to the unmodified
reference), the oracle tests and the GPU parity tests, so that every party sees
bit-identical inputs.  Everything is generated with numpy's PCG64
(`np.random.default_rng(seed)`), whose stream is stable across numpy versions,
and converted to fp32 once.

Constants cite the reference (paths relative to /root/reference):
  * intrinsics           sgam/inference_pipeline.py:61-65 (CLEVR), :83-89 (GoogleEarth)
  * start pose / steps   sgam/inference_pipeline.py:159-173
  * ddconfig             trained_models/*/config.yaml:18-29
"""
from collections import OrderedDict

import numpy as np

DDCONFIG = dict(double_z=False, z_channels=256, resolution=64, in_channels=4, out_ch=4, ch=128,
                ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[16], dropout=0.0)

DATASETS = {
    "clevr-infinite": dict(n_embed=16384, embed_dim=256, num_src=5, depth_lo=7.0, depth_hi=16.0),
    "google_earth": dict(n_embed=4096, embed_dim=256, num_src=3, depth_lo=1.4, depth_hi=3.8),
}


def intrinsics(dataset, res=256):
    """K of the inference pipeline, scaled to `res` (inference_pipeline.py:61-65, 83-89)."""
    if dataset == "clevr-infinite":
        K = np.array([[355.5555, 0, 128], [0, 355.5555, 128], [0, 0, 1]], dtype=np.float64)
        K[:2] *= res / 256.0
    elif dataset == "google_earth":
        K = np.array([[497.77774, 0, 256], [0, 497.77774, 256], [0, 0, 1]], dtype=np.float64)
        K[:2] *= res / 512.0
    else:
        raise NotImplementedError(dataset)
    return K


def grid_pose(dataset, i, j, step_size_denom=2):
    """World-to-camera (R, t) of grid cell (i, j): inference_pipeline.py:157-204."""
    if dataset == "google_earth":
        start = np.array([[1., 0., 0., -3.], [0., 0.86602527, -0.50000024, -6.],
                          [0., 0.50000024, 0.86602527, 2.], [0., 0., 0., 1.]])
        step_i = np.array([0., 0.11878788, 0.]) / step_size_denom
        step_j = np.array([0.12, 0, 0.]) / step_size_denom
    else:
        start = np.array([[1., 0., 0., -20.], [0., 0.95533651, -0.29552022, -20.],
                          [0., 0.29552022, 0.95533651, 0.], [0., 0., 0., 1.]])
        step_j = np.array([0.81632614, 0, 0.]) / step_size_denom
        step_i = np.array([0, 0.81632614, 0.]) / step_size_denom
    c2w = np.eye(4)
    c2w[:3, :3] = start[:3, :3]
    c2w[:3, 3] = start[:3, 3] + step_j * j + step_i * i
    c2w = c2w @ np.diag([1., -1., -1., 1.])
    w2c = np.linalg.inv(c2w)
    return w2c[:3, :3], w2c[:3, 3]


def relative_poses(dataset, tgt, srcs):
    """T_rel = T_tgt @ inv(T_src) in float64 -> (R_rels[N,3,3], t_rels[N,3]) (inference_pipeline.py:556-569)."""
    Rt, tt = grid_pose(dataset, *tgt)
    T_tgt = np.eye(4)
    T_tgt[:3, :3], T_tgt[:3, 3] = Rt, tt
    Rs, ts = [], []
    for s in srcs:
        R, t = grid_pose(dataset, *s)
        T_src = np.eye(4)
        T_src[:3, :3], T_src[:3, 3] = R, t
        T_rel = T_tgt @ np.linalg.inv(T_src)
        Rs.append(T_rel[:3, :3])
        ts.append(T_rel[:3, 3])
    return np.stack(Rs), np.stack(ts)


def smooth_field(rng, H, W, lo, hi, octaves=4):
    """Low-frequency random field in [lo, hi] (a plausible depth map: mostly smooth with a few steps)."""
    f = np.zeros((H, W))
    for o in range(octaves):
        n = 2 ** (o + 1) + 1
        coarse = rng.random((n, n))
        yi = np.linspace(0, n - 1, H)
        xi = np.linspace(0, n - 1, W)
        y0 = np.floor(yi).astype(int).clip(0, n - 2)
        x0 = np.floor(xi).astype(int).clip(0, n - 2)
        wy = (yi - y0)[:, None]
        wx = (xi - x0)[None, :]
        c = (coarse[y0][:, x0] * (1 - wy) * (1 - wx) + coarse[y0 + 1][:, x0] * wy * (1 - wx)
             + coarse[y0][:, x0 + 1] * (1 - wy) * wx + coarse[y0 + 1][:, x0 + 1] * wy * wx)
        f += c / (2 ** o)
    f = (f - f.min()) / (f.max() - f.min() + 1e-12)
    # a few box-shaped foreground objects (depth discontinuities -> occlusion / collisions)
    for _ in range(3):
        y, x = rng.integers(0, H), rng.integers(0, W)
        h, w = rng.integers(H // 16 + 1, H // 4 + 2), rng.integers(W // 16 + 1, W // 4 + 2)
        f[y:y + h, x:x + w] *= 0.6
    return (lo + (hi - lo) * f)


def scene_step_inputs(dataset, seed, res=256, batch=1, num_src=None, tgt=(1, 1), zero_frac=0.0):
    """One scene-generation step's `batch` dict (numpy, fp32) as prepare_batch_data builds it
    (inference_pipeline.py:533-609): src_imgs[B,N,H,W,3] in [-1,1] on the uint8 lattice,
    src_depths[B,N,H,W], Ks[B,N,3,3], K_invs, R_rels[B,N,3,3], t_rels[B,N,3]."""
    cfg = DATASETS[dataset]
    N = cfg["num_src"] if num_src is None else num_src
    rng = np.random.default_rng(seed)
    K = intrinsics(dataset, res)
    H = W = res
    # neighbours of the target in the pose grid (zig-zag predecessors)
    cand = [(tgt[0] - 1, tgt[1]), (tgt[0], tgt[1] - 1), (tgt[0] - 1, tgt[1] - 1), (tgt[0] - 1, tgt[1] + 1),
            (tgt[0] + 1, tgt[1] - 1), (tgt[0] - 2, tgt[1]), (tgt[0], tgt[1] - 2)]
    cand = [(max(a, 0), max(b, 0)) for a, b in cand]
    out = {k: [] for k in ("src_imgs", "src_depths", "Ks", "K_invs", "R_rels", "t_rels")}
    for b in range(batch):
        srcs = [cand[(k + b) % len(cand)] for k in range(N)]
        R_rels, t_rels = relative_poses(dataset, tgt, srcs)
        imgs = rng.integers(0, 256, size=(N, H, W, 3)).astype(np.float64) / 127.5 - 1.0
        # make the images smooth-ish but keep them on the uint8 lattice like PNG-loaded frames
        depths = np.stack([smooth_field(rng, H, W, cfg["depth_lo"], cfg["depth_hi"]) for _ in range(N)])
        if zero_frac > 0:
            depths = depths * (rng.random(depths.shape) >= zero_frac)
        out["src_imgs"].append(imgs)
        out["src_depths"].append(depths)
        out["Ks"].append(np.stack([K] * N))
        out["K_invs"].append(np.stack([np.linalg.inv(K)] * N))
        out["R_rels"].append(R_rels)
        out["t_rels"].append(t_rels)
    out = {k: np.stack(v).astype(np.float32) for k, v in out.items()}
    out["dst_img"] = np.zeros((batch, H, W, 3), np.float32)
    out["dst_depth"] = np.zeros((batch, H, W), np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# network parameters (checkpoint layout of the reference: SURVEY.md section 8b)
# ----------------------------------------------------------------------------------------------

def _resblock(p, name, cin, cout):
    p[f"{name}.norm1.weight"] = (cin,)
    p[f"{name}.norm1.bias"] = (cin,)
    p[f"{name}.conv1.weight"] = (cout, cin, 3, 3)
    p[f"{name}.conv1.bias"] = (cout,)
    p[f"{name}.norm2.weight"] = (cout,)
    p[f"{name}.norm2.bias"] = (cout,)
    p[f"{name}.conv2.weight"] = (cout, cout, 3, 3)
    p[f"{name}.conv2.bias"] = (cout,)
    if cin != cout:
        p[f"{name}.nin_shortcut.weight"] = (cout, cin, 1, 1)
        p[f"{name}.nin_shortcut.bias"] = (cout,)


def _attn(p, name, c):
    p[f"{name}.norm.weight"] = (c,)
    p[f"{name}.norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        p[f"{name}.{n}.weight"] = (c, c, 1, 1)
        p[f"{name}.{n}.bias"] = (c,)


def hot_path_param_shapes(n_embed, embed_dim=256, dd=DDCONFIG):
    """name -> shape for every tensor of the checkpoint the hot path reads, in module order
    (diffusionmodules/model.py:342-539, model.py:54-63)."""
    p = OrderedDict()
    ch, mult, nrb = dd["ch"], dd["ch_mult"], dd["num_res_blocks"]
    nres = len(mult)
    p["conv_in.weight"] = (4, 5, 1, 1)
    p["conv_in.bias"] = (4,)
    # encoder
    p["encoder.conv_in.weight"] = (ch, dd["in_channels"], 3, 3)
    p["encoder.conv_in.bias"] = (ch,)
    curr = dd["resolution"]
    in_mult = (1,) + tuple(mult)
    for l in range(nres):
        bi, bo = ch * in_mult[l], ch * mult[l]
        for b in range(nrb):
            _resblock(p, f"encoder.down.{l}.block.{b}", bi, bo)
            bi = bo
            if curr in dd["attn_resolutions"]:
                _attn(p, f"encoder.down.{l}.attn.{b}", bi)
        if l != nres - 1:
            p[f"encoder.down.{l}.downsample.conv.weight"] = (bi, bi, 3, 3)
            p[f"encoder.down.{l}.downsample.conv.bias"] = (bi,)
            curr //= 2
    _resblock(p, "encoder.mid.block_1", bi, bi)
    _attn(p, "encoder.mid.attn_1", bi)
    _resblock(p, "encoder.mid.block_2", bi, bi)
    p["encoder.norm_out.weight"] = (bi,)
    p["encoder.norm_out.bias"] = (bi,)
    zc = dd["z_channels"] * (2 if dd["double_z"] else 1)
    p["encoder.conv_out.weight"] = (zc, bi, 3, 3)
    p["encoder.conv_out.bias"] = (zc,)
    # decoder
    bi = ch * mult[-1]
    curr = dd["resolution"] // 2 ** (nres - 1)
    p["decoder.conv_in.weight"] = (bi, dd["z_channels"], 3, 3)
    p["decoder.conv_in.bias"] = (bi,)
    _resblock(p, "decoder.mid.block_1", bi, bi)
    _attn(p, "decoder.mid.attn_1", bi)
    _resblock(p, "decoder.mid.block_2", bi, bi)
    for l in reversed(range(nres)):
        bo = ch * mult[l]
        for b in range(nrb + 1):
            _resblock(p, f"decoder.up.{l}.block.{b}", bi, bo)
            bi = bo
            if curr in dd["attn_resolutions"]:
                _attn(p, f"decoder.up.{l}.attn.{b}", bi)
        if l != 0:
            p[f"decoder.up.{l}.upsample.conv.weight"] = (bi, bi, 3, 3)
            p[f"decoder.up.{l}.upsample.conv.bias"] = (bi,)
            curr *= 2
    p["decoder.norm_out.weight"] = (bi,)
    p["decoder.norm_out.bias"] = (bi,)
    p["decoder.conv_out.weight"] = (dd["out_ch"], bi, 3, 3)
    p["decoder.conv_out.bias"] = (dd["out_ch"],)
    # quantiser
    p["quantize.embedding.weight"] = (n_embed, embed_dim)
    p["quant_conv.weight"] = (embed_dim, dd["z_channels"], 1, 1)
    p["quant_conv.bias"] = (embed_dim,)
    p["post_quant_conv.weight"] = (dd["z_channels"], embed_dim, 1, 1)
    p["post_quant_conv.bias"] = (dd["z_channels"],)
    return p


def make_state_dict(n_embed, seed=0, embed_dim=256, dd=DDCONFIG, as_torch=True):
    """Random-init weights in the reference's checkpoint layout (no .ckpt ships with the reference:
    README.md:74 links Google Drive).  Conv weights U(+-1/sqrt(fan_in)) like torch's default,
    non-trivial GroupNorm affine, N(0,1) codebook (SURVEY.md section 7 'hard parts': the default
    U(+-1/n_e) init makes the arg-min ill-conditioned)."""
    rng = np.random.default_rng(1000 + seed)
    sd = OrderedDict()
    for name, shape in hot_path_param_shapes(n_embed, embed_dim, dd).items():
        if name == "quantize.embedding.weight":
            v = rng.standard_normal(shape)
        elif ".norm" in name and name.endswith(".weight"):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif ".norm" in name and name.endswith(".bias"):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith(".weight"):
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-b, b, shape)
        else:  # conv bias
            v = rng.uniform(-0.05, 0.05, shape)
        sd[name] = v.astype(np.float32)
    if as_torch:
        import torch
        sd = OrderedDict((k, torch.from_numpy(v)) for k, v in sd.items())
    return sd
