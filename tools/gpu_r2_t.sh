#!/bin/bash
# Round-2 GPU pass T: GroupNorm-apply variants (groups per thread, loads hoisted above the statistics prologue),
# end-to-end leg with the next step's upload on a copy stream, host-time diagnostics of the 512x512 loop.
mkdir -p gpurun_out
for u in 4 2 1; do
  ( SGAM_GN_APPLY_U=$u timeout 300 python tools/gn_apply_sweep.py ) > gpurun_out/t_gn_sweep_u$u.txt 2>&1; echo "rc=$?" >> gpurun_out/t_gn_sweep_u$u.txt
done
grep -h "SGAM_GN\|gn_apply\|weighted" gpurun_out/t_gn_sweep_u*.txt
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -m gpu -x -q ) > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log; tail -3 gpurun_out/t_pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench rc=$?" >> gpurun_out/t_bench.err; tail -3 gpurun_out/t_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/t_bench.json"))
e = d["e2e"]
print("value", round(d["value"], 1), "e2e", round(e["value"], 1), "serial", round(e["serial"]["value"], 1), "prefetch", round(e["prefetch"]["value"], 1),
      "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1))
c = d["configs"]
print("cfg2", round(c["configs[2]"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1), c["configs[4]"].get("host_ms_per_call"), "resident", round(c["configs[4]"]["resident_step"]["value"], 1))
PY
