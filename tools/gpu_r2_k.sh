#!/bin/bash
# Round-2 GPU pass K: new TSDF / get_x tests, wide split-K experiment (on vs off).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tsdf.py tests/test_gpu_pipeline.py tests/test_gpu_tc.py tests/test_gpu_bench_configs.py -x -q ) > gpurun_out/k_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/k_unit.log; tail -8 gpurun_out/k_unit.log
for w in 1 0; do
  SGAM_TC_WIDE_SPLITK=$w timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --dump-gemm gpurun_out/k_gemm_wide$w.tsv > gpurun_out/k_bench_wide$w.json 2> gpurun_out/k_bench_wide$w.err
done
python - <<'PY'
import json
for w in (1, 0):
    d = json.load(open(f"gpurun_out/k_bench_wide{w}.json"))
    print("wide", w, "value", round(d["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "roof", round(d["roofline"]["achieved"], 1))
    rows = [l.rstrip("\n").split("\t") for l in open(f"gpurun_out/k_gemm_wide{w}.tsv")][1:]
    for shape in ("(8, 16, 16, 512, 512, 4608)", "(8, 32, 32, 256, 256, 2304)", "(8, 32, 32, 512, 512, 4608)"):
        t = [float(r[4]) for r in rows if r[1] == shape]
        if t: print("   ", shape, len(t), "calls", round(sum(t), 3), "ms")
PY
