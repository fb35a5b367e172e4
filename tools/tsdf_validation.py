"""Depth-error statistics of the device TSDF volume on the reference's REAL seed frames (VERDICT r1
missing #2): Open3D is absent, so the ray-cast depth is validated against what is available without it --

  (a) self-consistency: integrate the seed frame, ray-cast its own pose, compare with the seed depth;
  (b) novel views: ray-cast the next poses of the trajectory and compare with the z-buffered forward splat of the same
      seed frame (stage (i) kernel, nearest-depth policy + 3x3 hole fill), the survey's stand-in for the Open3D render.

Needs the seed frames (`templates/`, staged by __graft_entry__.build() under baseline/_ref/) and a GPU.
    python tools/tsdf_validation.py [out.json]"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgam_neurips22_b200 import synthetic
from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration, forward_splat_depth
from sgam_neurips22_b200.model import VQModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEMPLATES = os.path.join(ROOT, "baseline", "_ref", "templates")


def stats(a, b, vox):
    ok = (a > 0) & (b > 0)
    e = np.abs(a - b)[ok]
    if e.size == 0:
        return {"pixels": 0}
    return {"pixels": int(ok.sum()), "coverage_tsdf": float((a > 0).mean()), "coverage_ref": float((b > 0).mean()),
            "median_m": float(np.median(e)), "mean_m": float(e.mean()), "p90_m": float(np.percentile(e, 90)),
            "p99_m": float(np.percentile(e, 99)), "median_voxels": float(np.median(e) / vox), "p90_voxels": float(np.percentile(e, 90) / vox),
            "median_rel": float(np.median(e / b[ok]))}


def run(ds, seed_index):
    model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="sgam_tsdfval_"))
    try:
        pipe = InfiniteSceneGeneration(model, ds, seed_index=seed_index, use_rgbd_integration=True, template_root=TEMPLATES,
                                       output_dim=(6, 6) if ds == "clevr-infinite" else (8, 1))
    finally:
        os.chdir(cwd)
    first = tuple(pipe._ordered_grid_coords[0])
    node = pipe.transform_grid[first[0]][first[1]]
    rgb, _ = pipe._frames[first]
    depth = pipe._seed_depth_single
    T0 = np.eye(4); T0[:3, :3], T0[:3, 3] = node["R"], node["t"]
    pipe.volume.integrate(depth.contiguous(), rgb, pipe.K, T0, depth_trunc=20.0)
    vox = pipe.volume.voxel_length

    def timed(fn, n=10):                                   # CUDA-event time of a call on REAL data, after one warm-up
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    H, W = pipe.image_resolution
    out = {"dataset": ds, "seed_index": seed_index, "voxel_length": vox, "sdf_trunc": pipe.volume.sdf_trunc,
           "seed_depth_range": [float(depth.min()), float(depth.max())],
           "units_in_use": pipe.volume.units_in_use(), "volume_bytes": pipe.volume.memory_bytes(), "dropped_units": pipe.volume.dropped_units()}
    d0 = pipe.volume.render_depth(pipe.K, T0, H, W, z_far=pipe._z_far).cpu().numpy()
    out["self_view_vs_seed_depth"] = stats(d0, depth.cpu().numpy(), vox)
    out["raycast_ms_real_seed_frame"] = timed(lambda: pipe.volume.render_depth(pipe.K, T0, H, W, z_far=pipe._z_far))
    # the splat path sees the seed depth the frame store holds (for CLEVR: after the second ray->z conversion, :582-590);
    # compare like with like: splat the depth that was integrated
    pipe._frames[first] = (rgb, depth)
    views = {}
    for k, c in enumerate(pipe._ordered_grid_coords[1:4], 1):
        n = pipe.transform_grid[c[0]][c[1]]
        T = np.eye(4); T[:3, :3], T[:3, 3] = n["R"], n["t"]
        d_t = pipe.volume.render_depth(pipe.K, T, H, W, z_far=pipe._z_far).cpu().numpy()
        d_s = forward_splat_depth(pipe, [node], T).cpu().numpy()
        views[f"pose{k}_{c[0]}_{c[1]}"] = stats(d_t, d_s, vox)
    out["novel_views_vs_forward_splat"] = views
    # re-integration of the same real frame (the reference fuses a selected source at every step): time of touch + integrate
    out["integrate_ms_real_seed_frame"] = timed(lambda: pipe.volume.integrate(depth.contiguous(), rgb, pipe.K, T0, depth_trunc=20.0), n=5)
    return out


if __name__ == "__main__":
    res = [run("google_earth", s) for s in range(5)] + [run("clevr-infinite", 0)]
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "tsdf_validation.json")
    json.dump(res, open(dst, "w"), indent=1)
    for r in res:
        sv = r["self_view_vs_seed_depth"]
        nv = list(r["novel_views_vs_forward_splat"].values())
        print(r["dataset"], r["seed_index"], "self: median %.4f m (%.2f vox) p90 %.4f" % (sv["median_m"], sv["median_voxels"], sv["p90_m"]),
              "| novel: " + ", ".join("median %.4f m (%.2f vox) p90 %.4f cov %.2f/%.2f" % (v["median_m"], v["median_voxels"], v["p90_m"], v["coverage_tsdf"], v["coverage_ref"]) for v in nv),
              "| units", r["units_in_use"], "bytes %.2f GB" % (r["volume_bytes"] / 2**30),
              "| raycast %.3f ms integrate %.3f ms" % (r["raycast_ms_real_seed_frame"], r["integrate_ms_real_seed_frame"]))
