#!/bin/bash
# Round-2 GPU pass D: paged TSDF, fused decoder head, sub-pixel upsample: targeted tests, full suite, bench, TSDF validation.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tsdf.py -x -q ) > gpurun_out/d_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/d_unit.log; tail -15 gpurun_out/d_unit.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log; tail -25 gpurun_out/d_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --dump-gemm gpurun_out/d_gemm.tsv ) > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?" >> gpurun_out/d_bench.err; tail -3 gpurun_out/d_bench.err
timeout 600 python tools/tsdf_validation.py gpurun_out/d_tsdf_validation.json > gpurun_out/d_tsdf_validation.log 2>&1
echo "tsdfval rc=$?" >> gpurun_out/d_tsdf_validation.log; tail -12 gpurun_out/d_tsdf_validation.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/d_bench.json"))
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1),
          "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"])
    c = d["configs"]
    print("cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1), {k: v for k, v in c["configs[2]"].items() if k.startswith("tsdf")})
    print("cfg4", round(c["configs[4]"]["value"], 1), "resident", round(c["configs[4]"]["resident_step"]["value"], 1), "roof", round(c["configs[4]"]["roofline"]["achieved"], 1))
    print("roof", round(d["roofline"]["achieved"], 1), d["roofline"]["by_op"])
except Exception as e:
    print("bench unreadable:", e)
PY
