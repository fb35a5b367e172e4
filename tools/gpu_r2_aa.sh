#!/bin/bash
# Round-2 GPU pass AA: producers of Downsample / Upsample inputs emit split-bf16 planes (no separate split pass).
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_pipeline.py tests/test_gpu_attention.py -m gpu -x -q ) > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aa_pytest.log; tail -5 gpurun_out/aa_pytest.log
for f in 1 0; do
( time SGAM_EMIT_SPLIT=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline ) > gpurun_out/aa_bench_es$f.json 2> gpurun_out/aa_bench_es$f.err
echo "bench rc=$?" >> gpurun_out/aa_bench_es$f.err
python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/aa_bench_es{sys.argv[1]}.json"))
e = d["e2e"]
print("emit_split", sys.argv[1], "value", round(d["value"], 1), "e2e", round(e["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
      "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"], "roof", round(d["roofline"]["frac"], 3))
PY
done
