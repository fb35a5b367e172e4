#!/bin/bash
# Round-2 GPU pass X: fused q/k/v projection for every AttnBlock (row-pitched q / k in the three-pass path too).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -x -q ) > gpurun_out/x_pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest_attn.log; tail -6 gpurun_out/x_pytest_attn.log
( timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_attention.py ) > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log; tail -5 gpurun_out/x_pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
echo "bench rc=$?" >> gpurun_out/x_bench.err; tail -3 gpurun_out/x_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/x_bench.json"))
e = d["e2e"]
print("value", round(d["value"], 1), "e2e", round(e["value"], 1), "serial", round(e["serial"]["value"], 1), "prefetch", round(e["prefetch"]["value"], 1),
      "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1),
      "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"])
c = d["configs"]
print("cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1), c["configs[4]"].get("host_ms_per_call"), "resident", round(c["configs[4]"]["resident_step"]["value"], 1))
print("roof", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), round(d["roofline"]["frac_mma_issue"], 3), d["roofline"]["by_op"].keys())
PY
