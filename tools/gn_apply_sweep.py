"""Per-shape latency of the GroupNorm-apply (+ swish + bf16 split) kernel at the shapes of one 8-trajectory step.
Twenty stream-ordered launches in one CUDA graph, so the figure is the full launch-to-last-store latency that sits between
two GEMMs of the step.  SGAM_GN_APPLY_U=1|2|4 (read once per process) selects the groups-per-thread variant.

    SGAM_GN_APPLY_U=2 python tools/gn_apply_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sgam_neurips22_b200 import _lib, ops  # noqa: E402

SHAPES = [(8, 256, 256, 128, 10), (8, 128, 128, 128, 7), (8, 128, 128, 256, 1), (8, 64, 64, 256, 14), (8, 64, 64, 128, 1),
          (8, 32, 32, 256, 7), (8, 32, 32, 512, 2), (8, 16, 16, 512, 21), (8, 16, 16, 256, 2),
          (1, 256, 256, 128, 10), (1, 64, 64, 256, 14), (1, 16, 16, 512, 21)]


def time_graph(fn, reps=20, rounds=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rounds):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return 1000.0 * a.elapsed_time(b) / (reps * rounds)


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    total = 0.0
    print(f"SGAM_GN_APPLY_U={os.environ.get('SGAM_GN_APPLY_U', '(default)')}")
    for B, H, W, C, count in SHAPES:
        x = torch.randn(B, H, W, C, device=dev)
        gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
        n = int(lib.sgam_tc_gn_partial_floats(B, H, W))
        part = torch.zeros(n, device=dev)
        tiles = (n - B * 64) // (B * 64)
        pv = part[:B * tiles * 64].view(B, tiles, 32, 2)
        pv[..., 1] = float(H * W * (C // 32)) / tiles                # sum 0, sum of squares = count: mean 0, variance 1
        x.gn_partial = part
        us = time_graph(lambda: ops.groupnorm_split(x, gamma, beta, True))
        mb = x.numel() * 8 / 1e6
        print(f"gn_apply  [{B},{H},{W},{C}]  {mb:8.1f} MB  {us:7.2f} us  {mb / us * 1e-3 * 1e3:7.0f} GB/s   x{count} per step")
        if B == 8:
            total += us * count
        us2 = time_graph(lambda: ops.split_bf16(x))
        print(f"split     [{B},{H},{W},{C}]  {mb:8.1f} MB  {us2:7.2f} us  {mb / us2 * 1e-3 * 1e3:7.0f} GB/s")
        del x, part
    print(f"weighted gn_apply total at 8 trajectories: {total:.1f} us per step")


if __name__ == "__main__":
    main()
