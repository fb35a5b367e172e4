#!/bin/bash
# Round-2 GPU pass N: bench with the double-buffered e2e leg; 512x512 launch list.
mkdir -p gpurun_out
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
echo "bench rc=$?" >> gpurun_out/n_bench.err; tail -5 gpurun_out/n_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/n_launches_ge512_b1.csv python bench.py --dataset google_earth --res 512 --batch 1 --profile-step > gpurun_out/n_ncu_512.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/n_bench.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "serial", round(d["e2e"]["serial"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1))
PY
