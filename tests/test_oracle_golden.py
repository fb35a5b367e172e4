"""Pin the oracle (oracle/) against the outputs of the unmodified reference (tests/golden).
CPU only.  Bit-exact for splat / median / masks / indices / depth coding / inverse warp; 2e-5 for the
conv network (same ATen fp32 operators; thread count may differ from the generating run)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import model as omodel
from oracle import native, recipes


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SPLAT_CASES = [0, 1, 2, 3, 4]


@pytest.mark.parametrize("ci", SPLAT_CASES)
def test_splat_bit_exact_vs_reference(golden, ci):
    ds, seed, res, B, nsrc, zf = golden[f"splat{ci}.meta"]
    batch = recipes.scene_step_inputs(str(ds), int(seed), res=int(res), batch=int(B),
                                      num_src=None if int(nsrc) < 0 else int(nsrc), zero_frac=float(zf))
    s = omodel.splat(batch)
    assert np.array_equal(np.packbits(s["inbounds"]), golden[f"splat{ci}.inbounds"])
    assert np.array_equal(np.packbits(s["mask"]), golden[f"splat{ci}.mask"])
    assert sha(s["proj_rgb"]) == str(golden[f"splat{ci}.sha_proj_rgb"])
    assert sha(s["merge_depth"]) == str(golden[f"splat{ci}.sha_merge_depth"])
    assert sha(s["merge_rgb"]) == str(golden[f"splat{ci}.sha_merge_rgb"])
    if int(res) <= 64:
        assert np.array_equal(s["merge_depth"], golden[f"splat{ci}.merge_depth"])
        assert np.array_equal(s["merge_rgb"], golden[f"splat{ci}.merge_rgb"])


def test_median_blur_bit_exact(golden):
    rng = np.random.default_rng(21)
    xm = rng.standard_normal((2, 4, 37, 53)).astype(np.float32)
    xm[rng.random(xm.shape) < 0.4] = 0
    assert np.array_equal(native.median_blur3(xm), golden["median.out"])


@pytest.mark.parametrize("ds", ["clevr-infinite", "google_earth"])
def test_get_x_bit_exact(golden, ds):
    batch = recipes.scene_step_inputs(ds, 31, res=64, batch=2, zero_frac=0.03)
    x, mask, code = omodel.get_x(batch, ds)
    assert np.array_equal(mask, golden[f"getx.{ds}.mask"])
    assert np.array_equal(x, golden[f"getx.{ds}.x"])


@pytest.mark.parametrize("ds", ["clevr-infinite", "google_earth"])
def test_vq_indices_bit_exact(golden, state_dicts, ds):
    sd = state_dicts(ds)
    rng = np.random.default_rng(41)
    z = rng.standard_normal((1, 256, 16, 16)).astype(np.float32) * 0.9
    for canonical in (False, True):
        zq, idx = omodel.quantize(sd, torch.from_numpy(z), canonical=canonical)
        assert np.array_equal(idx.numpy()[0], golden[f"vq.{ds}.idx"]), f"canonical={canonical}"
        assert sha(zq.numpy()) == str(golden[f"vq.{ds}.sha_zq"])
    # canonical fp32 distances agree with the reference's to a small fraction of the top-2 gap
    zt = np.ascontiguousarray(z.transpose(0, 2, 3, 1).reshape(-1, 256))
    _, dmin, d2 = native.vq_nearest(zt, sd["quantize.embedding.weight"].numpy())
    assert np.allclose(d2 - dmin, golden[f"vq.{ds}.gap"], atol=2e-3)


@pytest.mark.parametrize("ds", ["clevr-infinite", "google_earth"])
def test_network_64(golden, state_dicts, ds):
    sd = state_dicts(ds)
    rng = np.random.default_rng(51)
    xin = rng.uniform(-1, 1, (1, 4, 64, 64)).astype(np.float32)
    mk = (rng.random((1, 1, 64, 64)) < 0.3)
    dec, pre, zq, idx = omodel.forward(sd, xin, mk)
    assert np.allclose(pre.numpy(), golden[f"net64.{ds}.pre_quant"], atol=2e-5, rtol=2e-5)
    assert np.array_equal(idx.numpy()[0], golden[f"net64.{ds}.idx"])
    assert np.allclose(dec.numpy(), golden[f"net64.{ds}.dec"], atol=5e-5, rtol=5e-5)


def test_config1_128_forward(golden, state_dicts):
    """BASELINE.json configs[0]: CLEVR 128x128 encode -> VQ -> decode on CPU."""
    sd = state_dicts("clevr-infinite")
    torch.manual_seed(0)
    x = torch.randn(1, 4, 128, 128)
    assert sha(x.numpy()) == str(golden["cfg1.x_sha"])
    dec, pre, zq, idx = omodel.forward(sd, x, None)
    ref = golden["cfg1.dec"]
    rel = np.linalg.norm(dec.numpy() - ref) / np.linalg.norm(ref)
    assert rel < 2e-5, rel


@pytest.mark.parametrize("ci", [0, 1])
def test_inverse_warp_bit_exact(golden, ci):
    ds, seed, res = golden[f"invwarp{ci}.meta"]
    batch = recipes.scene_step_inputs(str(ds), int(seed), res=int(res), batch=1)
    src = np.ascontiguousarray(batch["src_imgs"].transpose(0, 1, 4, 2, 3))
    Ks = torch.from_numpy(batch["Ks"])
    T = torch.from_numpy(golden[f"invwarp{ci}.T_tgt2srcs"])
    proj = (Ks.view(-1, 3, 3) @ T.view(-1, 4, 4)[:, :3]).numpy()          # inference_pipeline.py:696
    Kinv_tgt = Ks[:, 0].inverse().numpy()
    out, best = native.inverse_warp(src, batch["src_depths"], golden[f"invwarp{ci}.tgt_depth"][None], Kinv_tgt, proj)
    assert np.array_equal(out[0], golden[f"invwarp{ci}.out"])


@pytest.mark.parametrize("ds", ["clevr-infinite", "google_earth"])
def test_full_step_256(golden, state_dicts, ds):
    """configs[1] / configs[2]-shaped step at 256x256: splat -> encode -> VQ -> decode -> uint8 / metric depth."""
    sd = state_dicts(ds)
    batch = recipes.scene_step_inputs(ds, 61, res=256, batch=1)
    r = omodel.scene_step(sd, batch, ds)
    assert sha(r["x"]) == str(golden[f"step256.{ds}.sha_x"])
    assert np.array_equal(np.packbits(r["mask"].astype(np.uint8)), golden[f"step256.{ds}.mask"])
    assert np.array_equal(r["idx"][0], golden[f"step256.{ds}.idx"])
    assert np.allclose(r["pre_quant"], golden[f"step256.{ds}.pre_quant"], atol=5e-5, rtol=5e-5)
    assert np.allclose(r["dec"][:, :, ::4, ::4], golden[f"step256.{ds}.dec_sub"], atol=1e-4, rtol=1e-4)
    # uint8 truncation can flip by one code on a handful of pixels when dec differs in the last ulps
    d = np.abs(r["rgb_u8"][::4, ::4].astype(int) - golden[f"step256.{ds}.rgb_sub"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    ok = np.abs(golden[f"step256.{ds}.depth_sub"]) < 50
    assert np.allclose(r["depth"][::4, ::4][ok], golden[f"step256.{ds}.depth_sub"][ok], rtol=1e-3, atol=1e-3)


def test_prepare_pcd_oracle_is_bit_exact_to_the_reference():
    """inference_pipeline.py:1014-1036 (prepare_pcd) run by the unmodified reference (tests/golden/make_golden_pcd.py)."""
    import os
    from oracle import native
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prepare_pcd_vectors.npz"))
    for n in ("clevr", "ge"):
        xyz = native.unproject_world(g[n + "_depth"], np.linalg.inv(g[n + "_K"]), np.linalg.inv(g[n + "_Rt"]))
        assert np.array_equal(xyz, g[n + "_points"])
        assert np.array_equal(g[n + "_color"].reshape(-1, 3) / 255.0, g[n + "_colors"])
