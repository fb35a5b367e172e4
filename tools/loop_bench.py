"""Frames/s of the drop-in scene loop (InfiniteSceneGeneration.one_step_prediction, no disk writes) on synthetic seeds."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgam_neurips22_b200 import synthetic
from sgam_neurips22_b200.model import VQModel
from sgam_neurips22_b200.inference_pipeline import InfiniteSceneGeneration, forward_splat_depth

ds = sys.argv[1] if len(sys.argv) > 1 else "clevr-infinite"
rgbd = len(sys.argv) > 2 and sys.argv[2] in ("rgbd", "rgbd-splat")        # rgbd: device TSDF volume; rgbd-splat: forward-splat stand-in
stand_in = len(sys.argv) > 2 and sys.argv[2] == "rgbd-splat"
profile = "--stages" in sys.argv
os.chdir(tempfile.mkdtemp())
model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
rng = np.random.default_rng(0)
lo, hi = synthetic.DATASETS[ds]["depth"]
yy, xx = np.meshgrid(np.linspace(0, 1, 256), np.linspace(0, 1, 256), indexing="ij")
seed = (rng.integers(0, 256, (256, 256, 3)).astype(np.uint8), (lo + (hi - lo) * (0.5 + 0.3 * np.sin(3 * xx) * np.cos(2 * yy))).astype(np.float32))
dim = (6, 6) if ds == "clevr-infinite" else (36, 1)
pipe = InfiniteSceneGeneration(model, ds, seed_frame=seed, output_dim=dim, use_rgbd_integration=rgbd,
                               tsdf_depth_fn=forward_splat_depth if stand_in else None)
n = dim[0] * dim[1] - 1
stage_ms = {}
if profile:                                   # synchronising per-stage timers (perturbs the loop; for diagnosis only)
    def wrap(obj, name):
        fn = getattr(obj, name)
        def timed(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize(); stage_ms[name] = stage_ms.get(name, 0.0) + 1000 * (time.perf_counter() - t0)
            return r
        setattr(obj, name, timed)
    for nm in ("get_src_grid_coords", "prepare_batch_data", "rgbd_integration", "inverse_warping"):
        wrap(pipe, nm)
    if pipe.volume is not None:
        wrap(pipe.volume, "integrate"); wrap(pipe.volume, "render_depth")
    wrap(model, "get_x"); wrap(model, "forward")
t_first = None
for i in range(n):
    if i == 5:
        torch.cuda.synchronize(); t_first = time.perf_counter()
    pipe.one_step_prediction(pipe.next_pose(pipe.curr), save_res_to_disk=False)
    pipe.curr += 1
torch.cuda.synchronize()
dt = time.perf_counter() - t_first
print(f"{ds} rgbd={rgbd}: {(n - 5) / dt:.1f} frames/s ({1000 * dt / (n - 5):.2f} ms/frame) over {n - 5} sequential frames, graph={getattr(model, 'use_cuda_graph', False)}")
if profile:
    print({k: round(v / n, 3) for k, v in stage_ms.items()}, "ms/frame (prepare_batch_data includes rgbd_integration and inverse_warping)")
