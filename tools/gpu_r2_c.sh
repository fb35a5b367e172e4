#!/bin/bash
# Round-2 GPU pass C: fixed tests, ncu --set full of the fused attention kernel, launch list of the step, PDL probes.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_pipeline.py -q ) > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log; tail -4 gpurun_out/c_pytest.log
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd -c 2 -f -o gpurun_out/c_full_attn \
    python bench.py --profile-step > gpurun_out/c_ncu_full_attn.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c_launches_step.csv \
    python bench.py --profile-step > gpurun_out/c_ncu_step.log 2>&1
# programmatic dependent launch: cluster kernels (family 2) never launched early, everything else early
for mask in 0 29 13; do
  SGAM_PDL=$mask timeout -s KILL 120 python tools/pdl_probe.py 512 1 > gpurun_out/c_pdl_512_$mask.txt 2>&1; echo "rc=$?" >> gpurun_out/c_pdl_512_$mask.txt
  SGAM_PDL=$mask timeout -s KILL 120 python tools/pdl_probe.py 256 1 > gpurun_out/c_pdl_256_$mask.txt 2>&1; echo "rc=$?" >> gpurun_out/c_pdl_256_$mask.txt
done
tail -2 gpurun_out/c_pdl_*.txt
for mask in 29; do
  SGAM_PDL=$mask timeout -s KILL 400 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/c_bench_pdl$mask.json 2> gpurun_out/c_bench_pdl$mask.err
  echo "bench pdl$mask rc=$?"
done
python - <<'PY'
import json
for f in ("gpurun_out/c_bench_pdl29.json",):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
