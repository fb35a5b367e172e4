"""In-graph latency (20 stream-ordered launches in one CUDA graph) of the short-K tensor-core launches of the 8-trajectory step:
the AttnBlock output projection (1x1 conv + residual + GroupNorm statistics) and the fused q/k/v projection.
SGAM_TC_EPI8=0|1 selects four / eight epilogue warps in the pair kernel (read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sgam_neurips22_b200 import ops  # noqa: E402
from gn_apply_sweep import time_graph  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    print(f"SGAM_TC_EPI8={os.environ.get('SGAM_TC_EPI8', '(default)')}")
    for (B, H, W, C) in ((8, 64, 64, 256), (1, 64, 64, 256), (8, 16, 16, 512)):
        x = torch.randn(B, H, W, C, generator=g, device=dev)
        xs = ops.split_bf16(x)
        w1 = ops.split_weight((torch.randn(C, C, generator=g, device=dev) * C ** -0.5), pad_rows_to=32)
        b1 = torch.randn(C, generator=g, device=dev)
        w3 = ops.split_weight(torch.randn(3 * C, C, generator=g, device=dev) * C ** -0.5)
        b3 = torch.randn(3 * C, generator=g, device=dev)
        t_proj = time_graph(lambda: ops.conv2d_tc(xs, w1, b1, residual=x, ksize=1, cout=C, gn_stats=True), reps=10)
        t_q = time_graph(lambda: ops.conv2d_tc(xs, w1, b1, ksize=1, cout=C, out_f32=False, out_split=True, gn_stats=False), reps=10)
        t_qkv = time_graph(lambda: ops.qkv_tc(xs, w3, b3), reps=10) if ops.qkv_tc_supported(B, H, W, C) else float("nan")
        flops = 2.0 * B * H * W * C * C
        print(f"[{B},{H},{W},{C}]  proj_out (1x1 + residual + stats) {t_proj:6.2f} us ({flops / t_proj * 1e-6:6.1f} TFLOP/s)   "
              f"q alone {t_q:6.2f} us   fused qkv {t_qkv:6.2f} us ({3 * flops / t_qkv * 1e-6:6.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
