#!/bin/bash
# Copies the outputs of tools/gpu_r2_y.sh (gpurun_out/y_*) into the tracked evidence files under profiles/.
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/y_bench.json profiles/r2_bench_final.json
python tools/summarize_ncu.py gpurun_out/y_launches_step.csv > profiles/r2_launches_final_summary.txt
python tools/summarize_ncu.py gpurun_out/y_launches_b1.csv > profiles/r2_launches_final_b1_summary.txt
gzip -c gpurun_out/y_launches_step.csv > profiles/r2_launches_final_raw.csv.gz
python tools/traffic_from_launches.py gpurun_out/y_launches_step.csv profiles/r2_traffic.json "profiles/r2_launches_final_raw.csv.gz: ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py --profile-step (one eager CLEVR 256x256 step, 8 trajectories, final round-2 kernels)" > /dev/null
cp gpurun_out/y_gemm.tsv profiles/r2_gemm_calls_final.tsv
{ echo "ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,lts__throughput...,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_*.sum --clock-control none -k regex:'tc_gemm|attn_fwd' python bench.py --profile-step"
  echo "(every tensor-core launch of one eager CLEVR 256x256 step, 8 trajectories, final round-2 kernels, launch order; tools/gemm_metrics_table.py; cold-cache, serialised)"
  cat gpurun_out/y_gemm_metrics.txt; } > profiles/r2_gemm_metrics.txt
{ grep "passed" gpurun_out/y_pytest.log; cat gpurun_out/y_smoke.log; } > profiles/r2_gpu_tests_final.txt
{ sed -n 1,8p profiles/r2_compute_sanitizer.txt; echo; grep -v "and Read access\|and Write access" gpurun_out/y_sanitizer.txt | cut -c1-220 | head -60; } > /tmp/san.txt
mv /tmp/san.txt profiles/r2_compute_sanitizer.txt
