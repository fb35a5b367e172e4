"""In-graph latency of the fused GroupNorm + conv kernel against GroupNorm-apply + conv (128 channels, 8 x 256 x 256)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sgam_neurips22_b200 import ops
from gn_apply_sweep import time_graph
g = torch.Generator().manual_seed(0)
B, H, W, C = 8, 256, 256, 128
x0 = torch.randn(B, H, W, C, generator=g).cuda()
w0 = ops.split_weight((torch.randn(C, 9 * C, generator=g) * 0.03).cuda(), pad_rows_to=32)
x = ops.conv2d_tc(ops.split_bf16(x0), w0, torch.zeros(C, device="cuda"), ksize=3, gn_stats=True)
part = x.gn_partial
ga, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
bias = torch.zeros(C, device="cuda")
def fused():
    x.gn_partial = part
    return ops.gn_conv2d_tc(x, ga, be, w0, bias, residual=x0)
def unfused():
    x.gn_partial = part
    return ops.conv2d_tc(ops.groupnorm_split(x, ga, be, True), w0, bias, residual=x0, ksize=3, gn_stats=True)
xs = ops.split_bf16(x0)
def conv_only():
    return ops.conv2d_tc(xs, w0, bias, residual=x0, ksize=3, gn_stats=True)
print("probe", os.environ.get("SGAM_GNCONV_PROBE", "0"), "fused %.1f us   apply + conv %.1f us   conv alone %.1f us" % (time_graph(fused, reps=5), time_graph(unfused, reps=5), time_graph(conv_only, reps=5)))
