#!/bin/bash
# Round-2 GPU pass Y (final evidence): full GPU suite + smoke + bench + launch lists (8 trajectories, 1 trajectory) + per-launch
# tensor-core metrics + sanitizer passes over the new projection / strided paths.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/y_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/y_pytest.log; tail -12 gpurun_out/y_pytest.log
( timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/y_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/y_smoke.log; tail -2 gpurun_out/y_smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 3 --dump-gemm gpurun_out/y_gemm.tsv ) > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
echo "bench rc=$?" >> gpurun_out/y_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/y_launches_step.csv python bench.py --profile-step > gpurun_out/y_ncu_step.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/y_launches_b1.csv python bench.py --batch 1 --profile-step > gpurun_out/y_ncu_b1.log 2>&1
M2=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M2 --clock-control none -k regex:'tc_gemm|attn_fwd' --csv --log-file gpurun_out/y_gemm_metrics.csv python bench.py --profile-step > gpurun_out/y_ncu_gemm.log 2>&1
python tools/gemm_metrics_table.py gpurun_out/y_gemm_metrics.csv > gpurun_out/y_gemm_metrics.txt 2>&1; tail -8 gpurun_out/y_gemm_metrics.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/y_bench.json"))
e = d["e2e"]
print("value", round(d["value"], 1), "e2e", round(e["value"], 1), "serial", round(e["serial"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1),
      "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"])
c = d["configs"]
print("cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1), "resident", round(c["configs[4]"]["resident_step"]["value"], 1))
print("roof", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), round(d["roofline"]["frac_mma_issue"], 3), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:40])
PY
( SAN_TOOLS="memcheck synccheck racecheck" bash tools/sanitize.sh ) > gpurun_out/y_sanitizer.txt 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|===" gpurun_out/y_sanitizer.txt | head -12
