#!/bin/bash
# Round-2 GPU pass S (final evidence): full GPU suite + smoke + bench after the GroupNorm finalize folding; launch lists (B = 8, B = 1).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s_pytest.log; tail -14 gpurun_out/s_pytest.log
( timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/s_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s_smoke.log; tail -2 gpurun_out/s_smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 3 --dump-gemm gpurun_out/s_gemm.tsv ) > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$?" >> gpurun_out/s_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/s_launches_step.csv python bench.py --profile-step > gpurun_out/s_ncu_step.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/s_launches_b1.csv python bench.py --batch 1 --profile-step > gpurun_out/s_ncu_b1.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/s_bench.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1),
      "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"])
c = d["configs"]
print("cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1), "resident", round(c["configs[4]"]["resident_step"]["value"], 1))
print("roof", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), round(d["roofline"]["frac_mma_issue"], 3), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:40])
PY
timeout 600 python tools/tsdf_validation.py gpurun_out/s_tsdf_validation.json > gpurun_out/s_tsdf_validation.log 2>&1
echo "tsdfval rc=$?" >> gpurun_out/s_tsdf_validation.log; tail -8 gpurun_out/s_tsdf_validation.log
