"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference on the
seeded inputs of oracle/recipes.py.  Runs only in the build container (the GPU box has no
/root/reference); the outputs are committed and are what pins the oracle.

    python tests/golden/make_golden.py

Shims: pytorch_lightning, data.utils.utils, matplotlib are stubbed in
sys.modules; torchvision.models.vgg16 is patched to skip the ImageNet download; Tensor.cuda is the
identity so InfiniteSceneGeneration.inverse_warping runs on CPU.  None of the reference's hot-path
files are modified or copied.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import recipes  # noqa: E402


def install_shims():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        global_step = 0
        global_rank = 0

        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl

    import importlib

    def instantiate_from_config(config):
        module, cls = config["target"].rsplit(".", 1)
        return getattr(importlib.import_module(module), cls)(**config.get("params", dict()))

    for name in ("data", "data.utils", "data.utils.utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["data.utils.utils"].instantiate_from_config = instantiate_from_config

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.imshow = plt.show = plt.title = plt.imsave = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt

    import torchvision.models as tvm
    orig = tvm.vgg16
    tvm.vgg16 = lambda pretrained=False, **k: orig(weights=None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    os.chdir(REF)


class NoneDict(dict):
    def __missing__(self, k):
        return None


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_reference_model(dataset, sd):
    import yaml
    from sgam.generative_sensing_module.model import VQModel
    cfg = yaml.safe_load(open(f"{REF}/trained_models/{dataset}/config.yaml"))
    params = cfg["model"]["params"]
    params["data_config"] = cfg["data"]["params"]
    params["ckpt_path"] = None
    params["online_kmeans_config"] = NoneDict(params["online_kmeans_config"])
    torch.manual_seed(0)
    model = VQModel(**params).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)   # the reference's own loader contract (model.py:104)
    assert not unexpected, unexpected
    assert all(k.startswith(("loss.", "perceptual_loss.")) for k in missing), [k for k in missing if not k.startswith(("loss.", "perceptual_loss."))]
    return model


def tb(batch):
    return {k: torch.from_numpy(v) for k, v in batch.items()}


@torch.no_grad()
def main():
    install_shims()
    torch.set_num_threads(1)          # deterministic index_put_
    from sgam.point_rendering import warp as W
    out = {}

    # ---- 1. forward splat (warp.py:193-286), sequential (parallel=False) AND parallel=True/1 thread --------
    splat_cases = [("clevr-infinite", 11, 64, 1, None, 0.0), ("google_earth", 12, 64, 2, None, 0.05),
                   ("clevr-infinite", 13, 32, 1, 2, 0.0), ("clevr-infinite", 14, 256, 1, None, 0.0),
                   ("google_earth", 15, 256, 1, None, 0.02)]
    for ci, (ds, seed, res, B, nsrc, zf) in enumerate(splat_cases):
        batch = recipes.scene_step_inputs(ds, seed, res=res, batch=B, num_src=nsrc, zero_frac=zf)
        b = tb(batch)
        x_src = b["src_imgs"].permute(0, 1, 4, 2, 3).contiguous()
        T = torch.eye(4)[None].repeat(b["R_rels"].shape[0] * b["R_rels"].shape[1], 1, 1)
        T[:, :3, :3] = b["R_rels"].view(-1, 3, 3)
        T[:, :3, 3] = b["t_rels"].view(-1, 3)
        T = T.view(*b["R_rels"].shape[:2], 4, 4)
        res_par = W.render_projection_from_srcs_fast(x_src, b["src_depths"], b["Ks"][:, 0], b["Ks"], T,
                                                     src_num=x_src.shape[1], parallel=True)
        if res <= 64:
            res_seq = W.render_projection_from_srcs_fast(x_src, b["src_depths"], b["Ks"][:, 0], b["Ks"], T,
                                                         src_num=x_src.shape[1], parallel=False)
            for a, c in zip(res_par[:3], res_seq[:3]):
                assert torch.equal(a, c), "parallel(1 thread) != sequential"
        md, mf, em, inb, _, _, proj = res_par
        key = f"splat{ci}"
        out[f"{key}.meta"] = np.array([ds, seed, res, B, -1 if nsrc is None else nsrc, zf], dtype=object).astype(str)
        out[f"{key}.mask"] = np.packbits(em.numpy().astype(np.uint8))
        out[f"{key}.inbounds"] = np.packbits(inb.numpy().astype(np.uint8))
        out[f"{key}.sha_merge_depth"] = np.array(sha(md.numpy()))
        out[f"{key}.sha_merge_rgb"] = np.array(sha(mf.numpy()))
        out[f"{key}.sha_proj_rgb"] = np.array(sha(proj.numpy()))
        if res <= 64:
            out[f"{key}.merge_depth"] = md.numpy()
            out[f"{key}.merge_rgb"] = mf.numpy()
        print(key, ds, res, "holes", int(em.sum()), "inbounds", int(inb.sum()), "/", inb.numel())

    # ---- 2. median_blur (warp.py:289-347) -------------------------------------------------------------
    rng = np.random.default_rng(21)
    xm = rng.standard_normal((2, 4, 37, 53)).astype(np.float32)
    xm[rng.random(xm.shape) < 0.4] = 0
    out["median.out"] = W.median_blur(torch.from_numpy(xm), (3, 3)).numpy()

    # ---- 3. get_x depth coding (model.py:179-269) + 4. VQ + 5. network + 6. full step -----------------
    from sgam.generative_sensing_module.modules.vqvae.quantize import VectorQuantizer2
    for ds in ("clevr-infinite", "google_earth"):
        cfg = recipes.DATASETS[ds]
        sd = recipes.make_state_dict(cfg["n_embed"], seed=0)
        model = build_reference_model(ds, sd)
        model.use_rgbd_integration = False

        # get_x at 64^2, batch 2
        batch = recipes.scene_step_inputs(ds, 31, res=64, batch=2, zero_frac=0.03)
        b = tb(batch)
        b["src_depths"] = b["src_depths"][..., None]                      # inference_pipeline.py:870
        x, x_dst, em, wd = model.get_x(b, ds, return_extrapolation_mask=True, no_depth_range=True, parallel=True)
        out[f"getx.{ds}.x"] = x.numpy()
        out[f"getx.{ds}.mask"] = em.numpy()

        # VQ: reference get_multiple_codewords(topk=1) and forward() on a seeded latent
        rng = np.random.default_rng(41)
        z = rng.standard_normal((1, 256, 16, 16)).astype(np.float32) * 0.9
        m = (rng.random((1, 1, 256, 256)) < 0.5)
        zq, _, info = model.quantize.get_multiple_codewords(torch.from_numpy(z), 1, 1, torch.from_numpy(m))
        zq2, _, info2 = model.quantize(torch.from_numpy(z))
        assert torch.equal(info[2].reshape(-1), info2[2].reshape(-1))
        assert torch.allclose(zq[:, 0], zq2.detach(), atol=1e-5)       # forward() adds the straight-through z + (z_q - z)
        out[f"vq.{ds}.idx"] = info2[2].numpy().reshape(16, 16)
        out[f"vq.{ds}.sha_zq"] = np.array(sha(zq[:, 0].numpy()))
        d = torch.sum(torch.from_numpy(z).permute(0, 2, 3, 1).reshape(-1, 256) ** 2, 1, keepdim=True) + \
            torch.sum(sd["quantize.embedding.weight"] ** 2, 1) - 2 * torch.from_numpy(z).permute(0, 2, 3, 1).reshape(-1, 256) @ sd["quantize.embedding.weight"].t()
        top2 = torch.topk(d, 2, dim=1, largest=False).values
        out[f"vq.{ds}.gap"] = (top2[:, 1] - top2[:, 0]).numpy()

        # network at 64^2 (latent 4x4): encoder / decoder separately and VQModel.forward(topk=None)
        rng = np.random.default_rng(51)
        xin = rng.uniform(-1, 1, (1, 4, 64, 64)).astype(np.float32)
        mk = (rng.random((1, 1, 64, 64)) < 0.3)
        torch.set_num_threads(8)
        quant, _, info, pre = model.encode(torch.from_numpy(xin), extrapolation_mask=torch.from_numpy(mk).float())
        dec = model.decode(quant)
        out[f"net64.{ds}.pre_quant"] = pre.numpy()
        out[f"net64.{ds}.idx"] = info[2].numpy().reshape(4, 4)
        out[f"net64.{ds}.dec"] = dec.numpy()
        print("net64", ds, float(pre.abs().mean()), float(dec.abs().mean()))

        if ds == "clevr-infinite":
            # config 1: 128^2, model(x) (BASELINE.json configs[0])
            torch.manual_seed(0)
            x128 = torch.randn(1, 4, 128, 128)
            dec128, _ = model(x128)
            out["cfg1.x_sha"] = np.array(sha(x128.numpy()))
            out["cfg1.dec"] = dec128.numpy()
            print("cfg1", float(dec128.abs().mean()))

        # config 2/3-style full step at 256^2 through the pipeline call (inference_pipeline.py:872-883)
        batch = recipes.scene_step_inputs(ds, 61, res=256, batch=1)
        b = tb(batch)
        b["src_depths"] = b["src_depths"][..., None]
        torch.set_num_threads(1)
        x, x_dst, em, wd = model.get_x(b, ds, return_extrapolation_mask=True, no_depth_range=True, parallel=True)
        torch.set_num_threads(8)
        decs, _, pre, quants = model(x, topk=1, extrapolation_mask=em, get_pre_quantized_feature=True,
                                     get_quantized_feature=True, sample_number=1)
        dec = decs[0][0]
        rgb = np.clip(((dec[0][:3] + 1) / 2 * 255.).permute(1, 2, 0).numpy(), 0, 255).astype(np.uint8)
        if ds == "clevr-infinite":
            depth = (1 / ((dec[0][3] + 1) / 2 * (1 / 7 - 1 / 16) + 1 / 16)).numpy()
        else:
            depth = (1 / ((dec.squeeze()[3] + 1) / 2 * (1 / 10.099975586 - 1 / 14.765625) + 1 / 14.765625) - 10).numpy()
        zf = pre.permute(0, 2, 3, 1).reshape(-1, 256)
        dd_ = torch.sum(zf ** 2, 1, keepdim=True) + torch.sum(sd["quantize.embedding.weight"] ** 2, 1) - 2 * zf @ sd["quantize.embedding.weight"].t()
        top2 = torch.topk(dd_, 2, dim=1, largest=False)
        out[f"step256.{ds}.sha_x"] = np.array(sha(x.numpy()))
        out[f"step256.{ds}.mask"] = np.packbits(em.numpy().astype(np.uint8))
        out[f"step256.{ds}.idx"] = top2.indices[:, 0].numpy().reshape(16, 16)
        out[f"step256.{ds}.gap"] = (top2.values[:, 1] - top2.values[:, 0]).numpy()
        out[f"step256.{ds}.pre_quant"] = pre.numpy()
        out[f"step256.{ds}.dec_sub"] = dec.numpy()[:, :, ::4, ::4].copy()
        out[f"step256.{ds}.dec_norm"] = np.array([float(dec.norm()), float(dec.abs().mean())])
        out[f"step256.{ds}.rgb_sub"] = rgb[::4, ::4].copy()
        out[f"step256.{ds}.depth_sub"] = depth[::4, ::4].copy()
        assert torch.equal(quants[0, 0], torch.nn.functional.embedding(top2.indices[:, 0], sd["quantize.embedding.weight"]).view(1, 16, 16, 256).permute(0, 3, 1, 2)[0])
        print("step256", ds, "holes", int(em.sum()), "min gap", float(out[f"step256.{ds}.gap"].min()))
        torch.set_num_threads(1)

    # ---- 7. inverse_warping (inference_pipeline.py:662-743) -------------------------------------------
    from sgam.inference_pipeline import InfiniteSceneGeneration
    pipe = object.__new__(InfiniteSceneGeneration)
    for ci, (ds, seed, res) in enumerate([("google_earth", 71, 64), ("clevr-infinite", 72, 128)]):
        batch = recipes.scene_step_inputs(ds, seed, res=res, batch=1)
        b = tb(batch)
        src = b["src_imgs"].permute(0, 1, 4, 2, 3).contiguous()
        # target depth: the forward-splat depth with ~5% zeros
        T = torch.eye(4)[None].repeat(b["R_rels"].shape[1], 1, 1)
        T[:, :3, :3] = b["R_rels"][0]
        T[:, :3, 3] = b["t_rels"][0]
        md = W.render_projection_from_srcs_fast(src, b["src_depths"], b["Ks"][:, 0], b["Ks"], T[None],
                                                src_num=src.shape[1], parallel=True)[0]
        rng = np.random.default_rng(seed)
        tgt_depth = md[0, 0] * torch.from_numpy((rng.random((res, res)) >= 0.05).astype(np.float32))
        T_tgt2srcs = torch.linalg.inv(T.double()).float()[None]
        warped = pipe.inverse_warping(src, b["src_depths"], tgt_depth[None], b["Ks"], b["Ks"][:, 0], T_tgt2srcs)
        out[f"invwarp{ci}.meta"] = np.array([ds, seed, res]).astype(str)
        out[f"invwarp{ci}.tgt_depth"] = tgt_depth.numpy()
        out[f"invwarp{ci}.T_tgt2srcs"] = T_tgt2srcs.numpy()
        out[f"invwarp{ci}.out"] = warped
        print("invwarp", ds, res, "nonzero", int((warped != 0).any(0).sum()))

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")))


if __name__ == "__main__":
    main()
