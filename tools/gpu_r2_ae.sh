#!/bin/bash
# Round-2 GPU pass AE: GroupNorm + swish + split inside the operand path of the 128-channel row-reuse conv kernel.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "fused_groupnorm_conv or conv2d_tc" ) > gpurun_out/ae_pytest_unit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ae_pytest_unit.log; tail -12 gpurun_out/ae_pytest_unit.log
( timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_pipeline.py -m gpu -x -q ) > gpurun_out/ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ae_pytest.log; tail -5 gpurun_out/ae_pytest.log
for f in 1 0; do
( time SGAM_FUSED_GNCONV=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline ) > gpurun_out/ae_bench_gc$f.json 2> gpurun_out/ae_bench_gc$f.err
echo "bench rc=$?" >> gpurun_out/ae_bench_gc$f.err; tail -1 gpurun_out/ae_bench_gc$f.err
python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ae_bench_gc{sys.argv[1]}.json"))
e = d["e2e"]
print("gnconv", sys.argv[1], "value", round(d["value"], 1), "e2e", round(e["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
      "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"], "roof", round(d["roofline"]["frac"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["by_op"].items()})
PY
done
