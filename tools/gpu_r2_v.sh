#!/bin/bash
# Round-2 GPU pass V: fused q/k/v projection GEMM (transposed-V epilogue) + strided q/k in the attention kernel.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -x -q ) > gpurun_out/v_pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest_attn.log; tail -15 gpurun_out/v_pytest_attn.log
( timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_pipeline.py -m gpu -x -q ) > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log; tail -5 gpurun_out/v_pytest.log
for f in 1 0; do
( time SGAM_FUSED_QKV=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline ) > gpurun_out/v_bench_qkv$f.json 2> gpurun_out/v_bench_qkv$f.err
echo "bench rc=$?" >> gpurun_out/v_bench_qkv$f.err
python - $f <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/v_bench_qkv{sys.argv[1]}.json"))
e = d["e2e"]
print("fused_qkv", sys.argv[1], "value", round(d["value"], 1), "e2e", round(e["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
      "loop", round(d["scene_loop"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"])
PY
done
