#!/bin/bash
# Round-2 GPU pass AD: swapped-operand 3x3 kernel with one pixel-row load per filter row (three taps from shifted descriptors).
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_pipeline.py -m gpu -x -q ) > gpurun_out/ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ad_pytest.log; tail -8 gpurun_out/ad_pytest.log
for f in 1 0; do
( time SGAM_TC_SWAPROW=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --dump-gemm gpurun_out/ad_gemm_sr$f.tsv ) > gpurun_out/ad_bench_sr$f.json 2> gpurun_out/ad_bench_sr$f.err
echo "bench rc=$?" >> gpurun_out/ad_bench_sr$f.err; tail -1 gpurun_out/ad_bench_sr$f.err
python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ad_bench_sr{sys.argv[1]}.json"))
e = d["e2e"]
print("swaprow", sys.argv[1], "value", round(d["value"], 1), "e2e", round(e["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
      "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["frac_mma_issue"], 3))
PY
grep -P "256, 256, 128, 128, 1152|128, 128, 128, 128, 1152" gpurun_out/ad_gemm_sr$f.tsv | awk -F'\t' '{print $2, $5, $6}' | sort | uniq -c | head -6
done
