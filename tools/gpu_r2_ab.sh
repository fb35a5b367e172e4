#!/bin/bash
# Round-2 GPU pass AB: end-to-end prefetch leg with the read-back on its own stream.
mkdir -p gpurun_out
( time timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline ) > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
echo "bench rc=$?" >> gpurun_out/ab_bench.err; tail -2 gpurun_out/ab_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/ab_bench.json"))
e = d["e2e"]
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(e["value"], 1), "serial", round(e["serial"]["value"], 1), "prefetch", round(e["prefetch"]["value"], 1), round(e["prefetch"]["ms_per_step"], 3))
PY
