"""Per-stage device timings of one scene-generation step (development aid; the judged numbers come from bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgam_neurips22_b200 import ops, synthetic
from sgam_neurips22_b200.model import VQModel

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
model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
eng = model.engine
batch = synthetic.scene_step_batch(ds, res=256, batch=B, seed=61)
host = {k: torch.from_numpy(v) for k, v in batch.items()}
N = host["src_depths"].shape[1]
Kinv, K_tgt = model._kinv(host["Ks"]), host["Ks"][:, 0].contiguous().cuda()
T = torch.eye(4).repeat(B, N, 1, 1)
T[..., :3, :3], T[..., :3, 3] = host["R_rels"], host["t_rels"]
rgb, dep, Tt = host["src_imgs"].cuda(), host["src_depths"].cuda(), T.cuda()
g = ops.splat_forward(rgb, dep, K_tgt, Kinv, Tt, ds, channels_last=True)
print("splat ms", timeit(lambda: ops.splat_forward(rgb, dep, K_tgt, Kinv, Tt, ds, channels_last=True), 20))
x, m = g["x"], g["mask"]
pre = eng.encode(x, m)
print("encode ms", timeit(lambda: eng.encode(x, m)))
print("vq ms", timeit(lambda: eng.quantize(pre), 20))
idx, zq = eng.quantize(pre)
print("decode ms", timeit(lambda: eng.decode(zq)))
t = timeit(lambda: eng.forward(x, m))
print("forward ms", t, "fps", 1000.0 * B / t)
