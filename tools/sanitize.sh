#!/bin/bash
# compute-sanitizer passes over a small end-to-end step (memcheck + racecheck + synccheck); development aid.
set -o pipefail
cat > /tmp/san_step.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from sgam_neurips22_b200 import ops, synthetic
from sgam_neurips22_b200.model import VQModel
ds, res = "google_earth", 64
model = synthetic.randomize_weights(VQModel(**synthetic.model_kwargs(ds)), seed=0).to("cuda:0").eval()
b = {k: torch.from_numpy(v) for k, v in synthetic.scene_step_batch(ds, res=res, batch=2, seed=3).items()}
x, _, mask, _ = model.get_x(b, ds, return_extrapolation_mask=True, no_depth_range=True)
decs, _, pre, quants = model(x, topk=1, extrapolation_mask=mask, get_pre_quantized_feature=True, get_quantized_feature=True)
rgb, depth = ops.frame_outputs(decs[0][0], ds)
torch.cuda.synchronize()
# swapped-operand GEMM (needs a full grid: 256 pixel boxes) with residual + fused GroupNorm statistics
g = torch.Generator().manual_seed(1)
xa = torch.randn(1, 256, 256, 128, generator=g).cuda()
wa = (torch.randn(128, 9 * 128, generator=g) * 0.03).cuda()
ya = ops.conv2d_tc(ops.split_bf16(xa), ops.split_weight(wa, pad_rows_to=32), torch.zeros(128, device="cuda"), residual=xa, ksize=3, gn_stats=True)
hi_a, lo_a = ops.groupnorm_split(ya, torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"), True)
torch.cuda.synchronize()
print("swap ok", float(ya.double().sum()), float(hi_a.float().abs().mean()))
# round-2 kernels: fused attention (2 pair tiles x 4 key tiles), sub-pixel Upsample (swap + pair kernels), fused stem / head
qa, ka, va = (torch.randn(2, 512, 256, generator=g).cuda() for _ in range(3))
oa = ops.attention_tc(ops.split_weight(qa), ops.split_weight(ka), ops.split_weight(va.transpose(1, 2).contiguous()), 256 ** -0.5)
torch.cuda.synchronize()
print("attn ok", float((oa[0].float() + oa[1].float()).double().sum()))
# fused q / k / v projection: vectorised transposed V store (W = 64), 2-byte fallback (ragged last tile), strided q / k consumers
for (Bq, Hq, Wq, Cq) in ((2, 64, 64, 256), (1, 8, 8, 128), (1, 24, 8, 128)):
    xq = torch.randn(Bq, Hq, Wq, Cq, generator=g).cuda()
    wq = (torch.randn(3 * Cq, Cq, generator=g) * Cq ** -0.5).cuda()
    q_, k_, vT_ = ops.qkv_tc(ops.split_bf16(xq), ops.split_weight(wq), torch.randn(3 * Cq, generator=g).cuda())
    s_ = ops.gemm_nt_tc(q_, k_, alpha=Cq ** -0.5)
    torch.cuda.synchronize()
    print("qkv ok", float(s_.double().sum()), float((vT_[0].float() + vT_[1].float()).double().sum()))
oq = ops.attention_tc(q_, k_, vT_, 128 ** -0.5) if ops.attention_tc_supported(1, 192, 128) else None
xq = torch.randn(2, 16, 16, 256, generator=g).cuda()
q_, k_, vT_ = ops.qkv_tc(ops.split_bf16(xq), ops.split_weight((torch.randn(768, 256, generator=g) / 16).cuda()), torch.zeros(768, device="cuda"))
oq = ops.attention_tc(q_, k_, vT_, 256 ** -0.5)
torch.cuda.synchronize()
print("qkv attn ok", float((oq[0].float() + oq[1].float()).double().sum()))
for (Bu, Hu, Cu, Co) in ((3, 128, 128, 128), (5, 64, 256, 256)):
    xu = torch.randn(Bu, Hu, Hu, Cu, generator=g).cuda()
    wu = (torch.randn(Co, 9 * Cu, generator=g) * 0.03).cuda()
    yu = ops.conv2d_tc_up2(ops.split_bf16(xu), ops.split_weight(ops.subpixel_weights(wu, Cu)), torch.zeros(Co, device="cuda"))
    hu, lu = ops.groupnorm_split(yu, torch.ones(Co, device="cuda"), torch.zeros(Co, device="cuda"), True)
    torch.cuda.synchronize()
    print("up2 ok", float(yu.double().sum()), float(hu.float().abs().mean()))
xs_ = torch.randn(1, 4, 64, 64, generator=g).cuda()
ys_ = ops.stem_conv_in(xs_, None, torch.randn(4, 5, generator=g).cuda(), torch.randn(4, generator=g).cuda(), (torch.randn(128, 36, generator=g) / 6).cuda(), torch.randn(128, generator=g).cuda())
yh_ = ops.gn_head_conv(ys_, torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"), (torch.randn(9, 128, 4, generator=g) * 0.03).cuda(), torch.zeros(4, device="cuda"))
torch.cuda.synchronize()
print("stem/head ok", float(ys_.double().sum()), float(yh_.double().sum()))
# RGB-D integration kernels on a small volume
from sgam_neurips22_b200.tsdf import TSDFVolume, frustum_box
K = np.array([[124.4, 0, 32.0], [0, 124.4, 32.0], [0, 0, 1.0]])
yy, xx = np.meshgrid(np.linspace(0, 1, 64), np.linspace(0, 1, 64), indexing="ij")
dm = torch.from_numpy((2.0 + 0.3 * np.sin(4 * xx) * np.cos(3 * yy)).astype(np.float32)).cuda()
im = torch.rand(64, 64, 3, device="cuda") * 2 - 1
lo, hi = frustum_box(K, [np.eye(4)], 64, 64, 2.8, pad=0.2)
vol = TSDFVolume(0.01, 0.03, lo, hi)
vol.integrate(dm, im, K, np.eye(4)); vol.integrate(dm, im, K, np.eye(4))
rd = vol.render_depth(K, np.eye(4), 64, 64, z_far=3.0)
pts, cols = vol.extract_point_cloud()
torch.cuda.synchronize()
print("tsdf ok", float(rd.mean()), int(pts.shape[0]))
import hashlib
print("step ok", float(decs[0][0].abs().mean()), hashlib.sha256(decs[0][0].cpu().numpy().tobytes()).hexdigest()[:16])
PY
echo "=== plain run (the sanitizer runs must reproduce this checksum) ==="
python /tmp/san_step.py 2>&1 | grep -E "step ok|tsdf ok|swap ok|attn ok|up2 ok|stem/head ok|qkv ok"
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout -s KILL 900 compute-sanitizer --tool $tool python /tmp/san_step.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|step ok|tsdf ok|swap ok|attn ok|up2 ok|stem/head ok|qkv ok|Error|hazard" | head -14
done
