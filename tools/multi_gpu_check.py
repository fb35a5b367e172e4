"""Multi-GPU check of the trajectory sharding + final map all-gather (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py

Every rank runs its shard of 2*world trajectories through TrajectoryBatch, the records are all-gathered over NCCL, and
each rank checks that (a) its own frames sit at its rank's offset of the gathered map, bit for bit, and (b) every rank
holds the same map (checksum all-reduce).  Rank 0 also re-runs trajectory 1 (owned by rank 1) locally and compares."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from sgam_neurips22_b200 import synthetic
from sgam_neurips22_b200.model import VQModel
from sgam_neurips22_b200.scene_batch import TrajectoryBatch

ds = "clevr-infinite"
model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to(f"cuda:{local}").eval()
seeds = []
for t in range(2 * world):
    rng = np.random.default_rng(100 + t)
    yy, xx = np.meshgrid(np.linspace(0, 1, 256), np.linspace(0, 1, 256), indexing="ij")
    seeds.append((rng.integers(0, 256, (256, 256, 3)).astype(np.uint8),
                  (8.0 + 7.0 * (0.5 + 0.3 * np.sin(3 * xx + t) * np.cos(2 * yy))).astype(np.float32)))
root = tempfile.mkdtemp(prefix=f"sgam_mg_{rank}_")
tb = TrajectoryBatch(model, ds, seeds, micro_batch=2, rank=rank, world_size=world, output_dim=(2, 2), output_root=root)
tb.scene_expansion()
rgb_l, depth_l, poses_l = tb.local_records()
xyz, col, (rgb, depth, poses) = tb.gather_map()
F = rgb_l.shape[0]
assert rgb.shape[0] == world * F, (rgb.shape, F)
assert torch.equal(rgb[rank * F:(rank + 1) * F], rgb_l) and torch.equal(depth[rank * F:(rank + 1) * F], depth_l)
assert torch.equal(poses[rank * F:(rank + 1) * F].cpu(), poses_l)
chk = torch.stack([rgb.double().sum(), depth.double().sum(), poses.double().sum().to(rgb.device)])
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi), "ranks hold different maps"
assert xyz.shape == (world * F * 65536, 3) and bool(torch.isfinite(xyz).all())
if rank == 0 and world > 1:
    # trajectory 1 lives on rank 1; with the same micro-batch size a local re-run must reproduce its frames exactly
    solo = TrajectoryBatch(model, ds, seeds, micro_batch=2, rank=1, world_size=world, output_dim=(2, 2), output_root=root + "_solo")
    solo.scene_expansion()
    r2, d2, _ = solo.local_records()
    assert torch.equal(r2, rgb[F:2 * F]) and torch.equal(d2, depth[F:2 * F]), "rank-1 trajectories differ when re-run on rank 0"
dist.barrier()
if rank == 0:
    print(f"multi-gpu check ok: world={world}, {world * F} frames gathered, {xyz.shape[0]} map points, maps identical on all ranks")
dist.destroy_process_group()
