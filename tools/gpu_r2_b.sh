#!/bin/bash
# Round-2 GPU pass B: the fused attention kernel (unit tests first, under a timeout), then the full GPU suite and a bench line.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_attention.py -x -q ) > gpurun_out/b_attn.log 2>&1
echo "attn rc=$?" >> gpurun_out/b_attn.log
tail -25 gpurun_out/b_attn.log
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -30 gpurun_out/b_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --dump-gemm gpurun_out/b_gemm.tsv ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?" >> gpurun_out/b_bench.err
tail -3 gpurun_out/b_bench.err
SGAM_ATTN=3pass timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/b_bench_3pass.json 2> gpurun_out/b_bench_3pass.err
python - <<'PY'
import json
for f in ("gpurun_out/b_bench.json", "gpurun_out/b_bench_3pass.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
              "roof", round(d["roofline"]["achieved"], 1), {k: (round(v["ms"], 3), round(v["tflops"], 1)) for k, v in d["roofline"]["by_op"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
