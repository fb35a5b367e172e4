#!/bin/bash
# Round-2 GPU pass I: VQ refinement at 32-code granularity (tests, bench), ncu of vq kernels, compute-sanitizer over every kernel.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -x -q ) > gpurun_out/i_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/i_unit.log; tail -6 gpurun_out/i_unit.log
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench rc=$?" >> gpurun_out/i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/i_bench.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "vq", d["kernels"]["vq"]["ms"], "splat", d["kernels"]["splat"]["ms"])
PY
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"vq_refine" -c 1 -f -o gpurun_out/i_full_vq \
    python bench.py --profile-step > gpurun_out/i_ncu_full_vq.log 2>&1
( time bash tools/sanitize.sh ) > gpurun_out/i_sanitize.txt 2>&1
tail -40 gpurun_out/i_sanitize.txt
