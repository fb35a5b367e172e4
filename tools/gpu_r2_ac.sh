#!/bin/bash
# Final default bench line (what the driver runs) + the reference arm, after the last bench.py change.
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; echo "bench rc=$?" >> gpurun_out/ac_bench.err; tail -2 gpurun_out/ac_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/ac_bench_ref.json 2> gpurun_out/ac_bench_ref.err; echo "ref rc=$?" >> gpurun_out/ac_bench_ref.err; tail -2 gpurun_out/ac_bench_ref.err; cat gpurun_out/ac_bench_ref.json | cut -c1-400
python - <<'PY'
import json
d = json.load(open("gpurun_out/ac_bench.json"))
e = d["e2e"]
print("value", round(d["value"], 1), "e2e", round(e["value"], 1), "serial", round(e["serial"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1),
      "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"])
c = d["configs"]
print("cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1), "resident", round(c["configs[4]"]["resident_step"]["value"], 1))
print("roof", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), round(d["roofline"]["frac_mma_issue"], 3), "cpu", d["cpu_baseline"]["value"])
PY
