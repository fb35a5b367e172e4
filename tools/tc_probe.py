"""First-contact probe of the tcgen05 GEMM kernel (run under `timeout`): prints errors instead of asserting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sgam_neurips22_b200 import ops

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())

torch.manual_seed(0)
dev = "cuda"
print("== plain GEMM ==", flush=True)
for (batch, M, N, K, ns) in [(1, 128, 128, 64, 1), (1, 128, 128, 64, 3), (1, 256, 128, 256, 3), (2, 384, 256, 192, 3), (1, 16, 32, 16, 3), (3, 4096, 4096, 256, 3)]:
    A = torch.randn(batch, M, K, device=dev); B = torch.randn(batch, N, K, device=dev)
    a = ops.split_weight(A); b = ops.split_weight(B)
    C = ops.gemm_nt_tc(a, b, alpha=0.5, nsplit=ns)
    torch.cuda.synchronize()
    ref = 0.5 * (A.double() @ B.double().transpose(1, 2))
    print(f"batch={batch} M={M} N={N} K={K} nsplit={ns}: rel={rel(C, ref):.3e} max={float((C.double()-ref).abs().max()):.3e}", flush=True)
print("== conv ==", flush=True)
for (B_, H, W, Cin, Cout, ks) in [(1, 16, 16, 128, 128, 3), (2, 32, 32, 128, 256, 3), (1, 64, 64, 256, 256, 1), (1, 4, 4, 512, 512, 3), (1, 256, 256, 128, 128, 3), (1, 8, 8, 256, 32, 1)]:
    x = torch.randn(B_, Cin, H, W); w = torch.randn(Cout, Cin, ks, ks) / (Cin * ks * ks) ** 0.5; bias = torch.randn(Cout)
    r = torch.randn(B_, Cout, H, W)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=ks // 2) + r.double()
    xs = ops.split_bf16(x.permute(0, 2, 3, 1).contiguous().to(dev))
    ws = ops.split_weight(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(dev))
    y, (yh, yl) = ops.conv2d_tc(xs, ws, bias.to(dev), residual=r.permute(0, 2, 3, 1).contiguous().to(dev), ksize=ks, out_split=True)
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).cpu()
    got2 = (yh.float() + yl.float()).permute(0, 3, 1, 2).cpu()
    print(f"B={B_} H={H} W={W} Cin={Cin} Cout={Cout} k={ks}: rel={rel(got, ref):.3e} split-out rel={rel(got2, ref):.3e}", flush=True)
def bench_conv(B_, H, W, Cin, Cout, ks):
    x = torch.randn(B_, H, W, Cin, device=dev); w = torch.randn(Cout, ks * ks * Cin, device=dev) / 34; bias = torch.zeros(Cout, device=dev)
    xs = ops.split_bf16(x); ws = ops.split_weight(w)
    for _ in range(3): ops.conv2d_tc(xs, ws, bias, ksize=ks)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.conv2d_tc(xs, ws, bias, ksize=ks)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B_ * H * W * Cout * ks * ks * Cin
    print(f"conv B={B_} {H}x{W} {Cin}->{Cout} k{ks}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)

def bench_gemm(batch, M, N, K, split_out=False):
    A = torch.randn(batch, M, K, device=dev); Bm = torch.randn(batch, N, K, device=dev)
    a = ops.split_weight(A); b = ops.split_weight(Bm)
    kw = dict(out_f32=not split_out, out_split=split_out)
    for _ in range(3): ops.gemm_nt_tc(a, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm_nt_tc(a, b, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm batch={batch} M={M} N={N} K={K} split_out={split_out}: {ms:.3f} ms  {2.0*batch*M*N*K/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)

bench_conv(8, 64, 64, 256, 256, 3)
bench_conv(8, 32, 32, 256, 256, 3)
bench_conv(8, 16, 16, 512, 512, 3)
bench_conv(8, 64, 64, 256, 256, 1)
bench_gemm(8, 4096, 4096, 256)
bench_gemm(8, 4096, 256, 4096, split_out=True)
print("== timing 128ch 256x256 3x3, batch 8 ==", flush=True)
x = torch.randn(8, 256, 256, 128, device=dev); w = torch.randn(128, 1152, device=dev) / 34; bias = torch.zeros(128, device=dev)
xs = ops.split_bf16(x); ws = ops.split_weight(w)
for ns in (3, 1):
    for _ in range(3): ops.conv2d_tc(xs, ws, bias, ksize=3, nsplit=ns)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.conv2d_tc(xs, ws, bias, ksize=3, nsplit=ns)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * 8 * 65536 * 128 * 1152
    print(f"nsplit={ns}: {ms:.3f} ms  useful {fl/ms/1e9:.1f} TFLOP/s  tensor-equivalent {fl*ns/ms/1e9:.1f} TFLOP/s", flush=True)
