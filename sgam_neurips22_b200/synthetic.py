"""Synthetic RGB-D scene-step inputs of the named resolution (there is no network for datasets): random colours
on the uint8 lattice, smooth random depth in the dataset's range, and the pose-grid neighbours the scene loop would
select (inference_pipeline.py:157-204, 507-531).  Used by bench.py and __graft_entry__.smoke()."""
import numpy as np

DATASETS = {
    "clevr-infinite": dict(n_embed=16384, num_src=5, depth=(7.0, 16.0)),
    "google_earth": dict(n_embed=4096, num_src=3, depth=(1.4, 3.8)),
}

DDCONFIG = dict(double_z=False, z_channels=256, resolution=64, in_channels=4, out_ch=4, ch=128,
                ch_mult=[1, 1, 2, 2, 4], num_res_blocks=2, attn_resolutions=[16], dropout=0.0)


def model_kwargs(dataset):
    return dict(ddconfig=dict(DDCONFIG), data_config=dict(dataset=dataset), lossconfig=None,
                n_embed=DATASETS[dataset]["n_embed"], embed_dim=256, phase="conditional_generation",
                online_kmeans_config=dict(do_online_kmeans_clustering=False))


def intrinsics(dataset, res):
    if dataset == "clevr-infinite":
        f, c = 355.5555 * res / 256.0, 128.0 * res / 256.0
    else:
        f, c = 497.77774 * res / 512.0, 256.0 * res / 512.0
    return np.array([[f, 0, c], [0, f, c], [0, 0, 1]], np.float64)


def _w2c(dataset, i, j):
    if dataset == "google_earth":
        R0 = np.array([[1., 0., 0.], [0., 0.86602527, -0.50000024], [0., 0.50000024, 0.86602527]])
        t0, si, sj = np.array([-3., -6., 2.]), np.array([0., 0.05939394, 0.]), np.array([0.06, 0., 0.])
    else:
        R0 = np.array([[1., 0., 0.], [0., 0.95533651, -0.29552022], [0., 0.29552022, 0.95533651]])
        t0, si, sj = np.array([-20., -20., 0.]), np.array([0., 0.40816307, 0.]), np.array([0.40816307, 0., 0.])
    c2w = np.eye(4)
    c2w[:3, :3] = R0
    c2w[:3, 3] = t0 + sj * j + si * i
    return np.linalg.inv(c2w @ np.diag([1., -1., -1., 1.]))


def scene_step_batch(dataset, res=256, batch=1, seed=0, num_src=None):
    """-> dict of numpy fp32 arrays shaped like prepare_batch_data's batch (B trajectories stacked on dim 0)."""
    cfg = DATASETS[dataset]
    N = cfg["num_src"] if num_src is None else num_src
    rng = np.random.default_rng(seed)
    K = intrinsics(dataset, res)
    tgt = (2, 2)
    neigh = [(1, 2), (2, 1), (1, 1), (1, 3), (3, 1), (0, 2), (2, 0), (0, 1)]
    out = dict(src_imgs=[], src_depths=[], Ks=[], R_rels=[], t_rels=[])
    lo, hi = cfg["depth"]
    yy, xx = np.meshgrid(np.linspace(0, 1, res), np.linspace(0, 1, res), indexing="ij")
    for b in range(batch):
        T_tgt = _w2c(dataset, *tgt)
        Rs, ts, deps = [], [], []
        for k in range(N):
            T_rel = T_tgt @ np.linalg.inv(_w2c(dataset, *neigh[(k + b) % len(neigh)]))
            Rs.append(T_rel[:3, :3])
            ts.append(T_rel[:3, 3])
            a = rng.uniform(-1, 1, 6)
            f = 0.5 + 0.2 * (a[0] * np.sin(3 * xx + a[1] * 3) + a[2] * np.cos(4 * yy + a[3] * 3) + a[4] * xx * yy + a[5] * (xx - yy))
            deps.append(lo + (hi - lo) * np.clip(f, 0, 1))
        out["src_imgs"].append(rng.integers(0, 256, (N, res, res, 3)).astype(np.float64) / 127.5 - 1.0)
        out["src_depths"].append(np.stack(deps))
        out["Ks"].append(np.stack([K] * N))
        out["R_rels"].append(np.stack(Rs))
        out["t_rels"].append(np.stack(ts))
    out = {k: np.stack(v).astype(np.float32) for k, v in out.items()}
    out["dst_img"] = np.zeros((batch, res, res, 3), np.float32)
    out["dst_depth"] = np.zeros((batch, res, res), np.float32)
    return out


def randomize_weights(model, seed=0):
    """Seeded random-init weights for measurements (no checkpoint ships with the reference): conv weights / biases
    U(+-1/sqrt(fan_in)) like torch's default, non-trivial GroupNorm affine, N(0,1) codebook (the
    default U(+-1/n_e) codebook makes the arg-min ill-conditioned).  Every tensor comes from the one generator, so
    all ranks / processes / both bench arms see identical weights."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    bound = 1.0
    for k, v in sd.items():
        if k == "quantize.embedding.weight":
            v.copy_(torch.randn(v.shape, generator=g))
        elif ".norm" in k:
            v.copy_((1.0 if k.endswith("weight") else 0.0) + 0.1 * torch.randn(v.shape, generator=g))
        else:
            if k.endswith(".weight"):
                bound = 1.0 / float(np.prod(v.shape[1:])) ** 0.5
            v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * bound)
    model.load_state_dict(sd)
    return model
