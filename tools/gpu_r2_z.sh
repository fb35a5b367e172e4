#!/bin/bash
# Round-2 GPU pass Z: eight epilogue warps in the pair kernel for short-K launches (1x1 convs, fused q/k/v projection).
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_attention.py tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_pipeline.py -m gpu -x -q ) > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log; tail -5 gpurun_out/z_pytest.log
for f in 1 0; do
( time SGAM_TC_EPI8=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --dump-gemm gpurun_out/z_gemm_epi$f.tsv ) > gpurun_out/z_bench_epi$f.json 2> gpurun_out/z_bench_epi$f.err
echo "bench rc=$?" >> gpurun_out/z_bench_epi$f.err
python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/z_bench_epi{sys.argv[1]}.json"))
e = d["e2e"]
print("epi8", sys.argv[1], "value", round(d["value"], 1), "e2e", round(e["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
      "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["by_op"].items()})
PY
done
grep -P "^qkv_tc|ksize': 1" gpurun_out/z_gemm_epi1.tsv | awk -F'\t' '{print $1, $2, $5}' | sort | uniq -c | sort -k1,1nr | head -8
