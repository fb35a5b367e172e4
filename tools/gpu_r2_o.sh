#!/bin/bash
# Round-2 GPU pass O: GroupNorm apply back-to-front (tests, bench, launch list).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_configs.py -x -q ) > gpurun_out/o_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/o_unit.log; tail -3 gpurun_out/o_unit.log
( time timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline ) > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/o_launches_step.csv python bench.py --profile-step > gpurun_out/o_ncu_step.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/o_bench.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1))
PY
python tools/summarize_ncu.py gpurun_out/o_launches_step.csv | head -8
