#!/bin/bash
# Round-2 GPU pass U: packed FFMA2 in the fp32 stem / head kernels, GroupNorm apply with 2 groups per thread and hoisted
# affine parameters; parity tests, bench, launch list.
mkdir -p gpurun_out
( SGAM_GN_APPLY_U=2 timeout 300 python tools/gn_apply_sweep.py ) > gpurun_out/u_gn_sweep_u2.txt 2>&1; grep -h "weighted" gpurun_out/u_gn_sweep_u2.txt
( timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_pipeline.py -m gpu -x -q ) > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log; tail -3 gpurun_out/u_pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
echo "bench rc=$?" >> gpurun_out/u_bench.err; tail -3 gpurun_out/u_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/u_launches_step.csv python bench.py --profile-step > gpurun_out/u_ncu_step.log 2>&1
python tools/summarize_ncu.py gpurun_out/u_launches_step.csv | head -24
python - <<'PY'
import json
d = json.load(open("gpurun_out/u_bench.json"))
e = d["e2e"]
print("value", round(d["value"], 1), "e2e", round(e["value"], 1), "serial", round(e["serial"]["value"], 1), "prefetch", round(e["prefetch"]["value"], 1),
      "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1))
c = d["configs"]
print("cfg2", round(c["configs[2]"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1), c["configs[4]"].get("host_ms_per_call"), "resident", round(c["configs[4]"]["resident_step"]["value"], 1))
PY
