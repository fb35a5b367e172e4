"""Per-stage device timings of one scene-generation step (development aid; the judged numbers come from bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import recipes, model as omodel
from sgam_neurips22_b200 import ops
from sgam_neurips22_b200.vqgan import VQGANEngine

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ds = sys.argv[1] if len(sys.argv) > 1 else "clevr-infinite"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = recipes.make_state_dict(recipes.DATASETS[ds]["n_embed"], 0)
eng = VQGANEngine(sd, recipes.DDCONFIG)
batch = recipes.scene_step_inputs(ds, 61, res=256, batch=B)
Ks = batch["Ks"]; Kinv = torch.from_numpy(Ks.reshape(-1,3,3)).inverse().numpy().reshape(Ks.shape)
T = omodel.src2tgt_transforms(batch["R_rels"], batch["t_rels"])
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rgb, dep, Kt, Ki, Tt = d(batch["src_imgs"]), d(batch["src_depths"]), d(Ks[:,0]), d(Kinv), d(T)
g = ops.splat_forward(rgb, dep, Kt, Ki, Tt, ds, channels_last=True)
print("splat ms", timeit(lambda: ops.splat_forward(rgb, dep, Kt, Ki, Tt, ds, channels_last=True), 20))
x, m = g["x"], g["mask"]
pre = eng.encode(x, m)
print("encode ms", timeit(lambda: eng.encode(x, m)))
print("vq ms", timeit(lambda: eng.quantize(pre), 20))
idx, zq = eng.quantize(pre)
print("decode ms", timeit(lambda: eng.decode(zq)))
t = timeit(lambda: eng.forward(x, m))
print("forward ms", t, "fps", 1000.0 * B / t)
