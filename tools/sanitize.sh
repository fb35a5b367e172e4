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
print("step ok", float(decs[0][0].abs().mean()))
PY
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool ==="
  timeout -s KILL 900 compute-sanitizer --tool $tool --kernel-regex kns=sgam --kernel-regex kne=at:: python /tmp/san_step.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|step ok|Error|hazard" | head -8
done
