"""The fused (flash-style) attention kernel of the 256-channel AttnBlocks, through the C ABI.
Reference: sgam/generative_sensing_module/modules/diffusionmodules/model.py:168-192 -- w = softmax(q^T k * C^-0.5) over the
keys, h = v w^T -- evaluated in float64 by plain torch ops.  Tolerance: 5e-5 rel-L2 (3-term split-bf16 products with fp32
accumulation, fp32 online softmax), the same bar as the other tensor-core unit ops; the whole-network bars (tokens
identical, decoded RGB-D < 1e-3) are held in test_gpu_tc.py / test_gpu_bench_configs.py with this kernel in the path."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from sgam_neurips22_b200 import ops as _ops
    return _ops


def reference(q, k, v, scale):
    s = torch.einsum("btc,bsc->bts", q.double(), k.double()) * scale
    return torch.einsum("bts,bsc->btc", torch.softmax(s, dim=-1), v.double())


def run(ops, q, k, v, scale, kv_splits=None):
    qs, ks = ops.split_weight(q), ops.split_weight(k)
    vts = ops.split_weight(v.transpose(1, 2).contiguous())
    oh, ol = ops.attention_tc(qs, ks, vts, scale, kv_splits=kv_splits)
    torch.cuda.synchronize()
    return oh.float() + ol.float()


# (B, T): one pair tile; two key tiles; ragged cluster count; > 74 pair tiles (persistent loop, O hand-over); config-5 length
@pytest.mark.parametrize("B,T", [(1, 256), (2, 512), (3, 1024), (1, 4096), (8, 4096), (1, 16384)])
def test_fused_attention_matches_float64(ops, B, T):
    C = 256
    g = torch.Generator(device="cuda").manual_seed(B * 100003 + T)
    q = torch.randn(B, T, C, generator=g, device="cuda")
    k = torch.randn(B, T, C, generator=g, device="cuda")
    v = torch.randn(B, T, C, generator=g, device="cuda") * 2 + 0.5
    scale = C ** -0.5
    assert ops.attention_tc_supported(B, T, C)
    out = run(ops, q, k, v, scale)
    ref = reference(q, k, v, scale)
    # 16384 keys: the P.V accumulator takes 3072 dependent tensor-core additions per row; the tensor pipe truncates when it
    # adds into the fp32 accumulator (~2^-25 relative per step, one-sided), so a long COHERENT sum (v has mean 0.5 here)
    # drifts by ~5e-5 -- the three-pass GEMM path accumulates the same chain.  Zero-mean data (the network's) does not.
    assert tuple(out.shape) == (B, T, C) and rel(out, ref) < (5e-5 if T < 16384 else 1e-4)
    # per-row accuracy too (a wrong row-sum or a dropped key tile shows up in single rows, not in the global norm)
    row_err = ((out.double() - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item()
    assert row_err < 5e-4, row_err
    assert torch.equal(out, run(ops, q, k, v, scale)), "fused attention must be deterministic"


@pytest.mark.parametrize("B,T,splits", [(1, 4096, 4), (2, 4096, 2), (1, 1024, 4), (3, 2048, 8), (1, 4096, 1)])
def test_key_split_items_merge_to_the_same_attention(ops, B, T, splits):
    """Small batches: the keys of a query tile are split over several SM pairs and merged (un-normalised O, reference
    maximum and row sum per split).  Same bar as the unsplit kernel; the library's own choice is covered as well."""
    from sgam_neurips22_b200 import _lib
    C = 256
    g = torch.Generator(device="cuda").manual_seed(B * 7 + T + splits)
    q = torch.randn(B, T, C, generator=g, device="cuda") * 1.5
    k = torch.randn(B, T, C, generator=g, device="cuda")
    k = k * torch.linspace(0.3, 3.0, T, device="cuda")[None, :, None]          # later splits hold the larger scores
    v = torch.randn(B, T, C, generator=g, device="cuda")
    scale = C ** -0.5
    ref = reference(q, k, v, scale)
    out = run(ops, q, k, v, scale, kv_splits=splits)
    assert rel(out, ref) < 5e-5
    assert ((out.double() - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item() < 5e-4
    auto = _lib.load().sgam_attention_tc_splits(B, T)
    assert auto >= 1 and (T // 128) % auto == 0
    assert rel(run(ops, q, k, v, scale), ref) < 5e-5


def test_fused_attention_growing_maxima_exercise_the_o_correction(ops):
    """Keys ordered so that every key tile raises the row maxima by far more than the lazy-rescale threshold (2^8): O and
    the row sums are corrected in tensor memory at every tile.  Peaked rows (softmax ~ one-hot) and flat rows mixed."""
    B, T, C = 2, 1024, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    q = torch.randn(B, T, C, generator=g, device="cuda")
    k = torch.randn(B, T, C, generator=g, device="cuda")
    ramp = torch.linspace(0.2, 6.0, T, device="cuda")[None, :, None]                # later keys -> much larger scores
    k = k * ramp + q.mean(dim=1, keepdim=True) * ramp * 2
    q[:, ::7] *= 0.01                                                                # nearly flat rows
    v = torch.randn(B, T, C, generator=g, device="cuda")
    scale = C ** -0.5
    s = torch.einsum("btc,bsc->bts", q.double(), k.double()) * scale
    tile_max = s.view(B, T, T // 128, 128).amax(-1)
    assert ((tile_max[..., 1:] - tile_max[..., :-1].cummax(-1).values) * 1.4427 > 8).any(), "the case must trigger the correction"
    out = run(ops, q, k, v, scale)
    ref = reference(q, k, v, scale)
    assert rel(out, ref) < 5e-5
    assert ((out.double() - ref).norm(dim=-1) / ref.norm(dim=-1)).max().item() < 5e-4


def test_fused_attention_equals_three_pass_path_in_the_engine(ops):
    """AttnBlock through the engine with SGAM_ATTN=fused vs SGAM_ATTN=3pass (QK^T GEMM, softmax pass, PV GEMM)."""
    from oracle import recipes
    from sgam_neurips22_b200.vqgan import VQGANEngine
    sd = recipes.make_state_dict(recipes.DATASETS["google_earth"]["n_embed"], seed=0)
    eng = VQGANEngine(sd, recipes.DDCONFIG, "cuda:0", mode="tc")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 32, 256, generator=g).cuda()                              # 1024 tokens
    name = "encoder.down.2.attn.0"
    old = os.environ.get("SGAM_ATTN")
    try:
        os.environ["SGAM_ATTN"] = "fused"
        y_f = eng.attn_block(name, x)
        os.environ["SGAM_ATTN"] = "3pass"
        y_3 = eng.attn_block(name, x)
    finally:
        if old is None:
            os.environ.pop("SGAM_ATTN", None)
        else:
            os.environ["SGAM_ATTN"] = old
    assert rel(y_f, y_3) < 2e-5


def test_unsupported_shapes_are_refused(ops):
    assert not ops.attention_tc_supported(1, 256, 512) and not ops.attention_tc_supported(1, 384, 256)
    q = ops.split_weight(torch.randn(1, 384, 256, device="cuda"))
    vt = ops.split_weight(torch.randn(1, 256, 384, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.attention_tc(q, q, vt, 0.0625)


# ---- fused q / k / v projection (sgam_qkv_tc): one GEMM, q | k row-major with a shared pitch, V stored transposed -------------
# (B, H, W, C): pair kernel with 256-wide column tiles (8 x 64 x 64), 1-CTA kernel (1 x 64 x 64), 128-pixel rows (512^2 config),
# the 512-channel mid block's width (not fused in the network, must still be right), W = 32
@pytest.mark.parametrize("B,H,W,C", [(8, 64, 64, 256), (1, 64, 64, 256), (2, 128, 128, 256), (3, 32, 32, 512), (1, 8, 32, 128),
                                     (8, 16, 16, 512), (1, 16, 16, 512), (2, 8, 8, 256)])
def test_fused_qkv_projection_matches_separate_ops(ops, B, H, W, C):
    if not ops.qkv_tc_supported(B, H, W, C):
        pytest.skip("shape not supported by the fused projection")
    g = torch.Generator(device="cuda").manual_seed(C * 7 + H)
    x = torch.randn(B, H, W, C, generator=g, device="cuda")
    w = torch.randn(3 * C, C, generator=g, device="cuda") * C ** -0.5
    bias = torch.randn(3 * C, generator=g, device="cuda")
    xs = ops.split_bf16(x)
    T = H * W
    q, k, vT = ops.qkv_tc(xs, ops.split_weight(w), bias)
    torch.cuda.synchronize()
    assert q[0].shape == (B, T, C) and k[0].shape == (B, T, C) and vT[0].shape == (B, C, T)
    # float64 reference of the three 1x1 convs (diffusionmodules/model.py:158-175)
    ref = torch.einsum("btc,nc->btn", x.view(B, T, C).double(), w.double()) + bias.double()
    got_q, got_k, got_v = q[0].float() + q[1].float(), k[0].float() + k[1].float(), (vT[0].float() + vT[1].float()).transpose(1, 2)
    assert rel(got_q, ref[..., :C]) < 5e-5 and rel(got_k, ref[..., C:2 * C]) < 5e-5 and rel(got_v, ref[..., 2 * C:]) < 5e-5
    # q, k: bit-identical to the separate launches they replace (same k-block order, same epilogue arithmetic); V^T: the separate
    # launch computes W_v . h^T, i.e. with the operand roles -- and so the order of the hi*lo and lo*hi terms inside a k-block --
    # exchanged: equal to fp32 accumulation-order noise
    flat = (xs[0].view(B, T, C), xs[1].view(B, T, C))
    for i, got in enumerate((q, k)):
        wi = ops.split_weight(w[i * C:(i + 1) * C].contiguous())
        sep = ops.conv2d_tc(xs, wi, bias[i * C:(i + 1) * C].contiguous(), ksize=1, cout=C, out_f32=False, out_split=True, gn_stats=False)
        assert torch.equal(sep[0].view(B, T, C), got[0]) and torch.equal(sep[1].view(B, T, C), got[1])
    sep_v = ops.gemm_nt_tc(ops.split_weight(w[2 * C:].contiguous()), flat, bias_m=bias[2 * C:].contiguous(), out_f32=False, out_split=True)
    assert rel(vT[0].float() + vT[1].float(), sep_v[0].float() + sep_v[1].float()) < 1e-6


def test_attention_reads_strided_q_and_k(ops):
    """q and k as the column halves of one [B,T,2C] tensor (what qkv_tc returns) give the bits of the dense call."""
    B, T, C = 2, 512, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    qk = torch.randn(B, T, 2 * C, generator=g, device="cuda")
    v = torch.randn(B, T, C, generator=g, device="cuda")
    qs, ks = ops.split_weight(qk[..., :C].contiguous()), ops.split_weight(qk[..., C:].contiguous())
    vts = ops.split_weight(v.transpose(1, 2).contiguous())
    dense = ops.attention_tc(qs, ks, vts, C ** -0.5)
    both = ops.split_weight(qk)
    strided = ops.attention_tc((both[0][..., :C], both[1][..., :C]), (both[0][..., C:], both[1][..., C:]), vts, C ** -0.5)
    torch.cuda.synchronize()
    assert torch.equal(dense[0], strided[0]) and torch.equal(dense[1], strided[1])


def test_gemm_nt_reads_row_pitched_operands(ops):
    """The three-pass attention of the 512-channel blocks takes q and k straight from qkv_tc's [B,T,2C] tensor."""
    B, T, C = 3, 256, 512
    g = torch.Generator(device="cuda").manual_seed(11)
    qk = torch.randn(B, T, 2 * C, generator=g, device="cuda")
    both = ops.split_weight(qk)
    q, k = (both[0][..., :C], both[1][..., :C]), (both[0][..., C:], both[1][..., C:])
    dense = ops.gemm_nt_tc((q[0].contiguous(), q[1].contiguous()), (k[0].contiguous(), k[1].contiguous()), alpha=C ** -0.5)
    strided = ops.gemm_nt_tc(q, k, alpha=C ** -0.5)
    torch.cuda.synchronize()
    assert torch.equal(dense, strided)
    ref = torch.einsum("btc,bsc->bts", qk[..., :C].double(), qk[..., C:].double()) * C ** -0.5
    assert rel(strided, ref) < 5e-5
