#!/bin/bash
# Round-2 GPU pass G: fused stem, head kernel with batched staging loads: tests, bench, ncu of both, batch-1 launch list.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_configs.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q ) > gpurun_out/g_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/g_unit.log; tail -8 gpurun_out/g_unit.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --dump-gemm gpurun_out/g_gemm.tsv ) > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?" >> gpurun_out/g_bench.err; tail -3 gpurun_out/g_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/g_launches_step.csv \
    python bench.py --profile-step > gpurun_out/g_ncu_step.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/g_launches_b1.csv \
    python bench.py --batch 1 --profile-step > gpurun_out/g_ncu_b1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gn_head_conv" -c 1 -f -o gpurun_out/g_full_head_stem \
    python bench.py --profile-step > gpurun_out/g_ncu_full.log 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/g_bench.json"))
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1),
          "tb", round(d["trajectory_batch"]["value"], 1), "launches", d["gpu_launches_per_step"])
    c = d["configs"]
    print("cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1))
    print("cfg4", round(c["configs[4]"]["value"], 1), "resident", round(c["configs[4]"]["resident_step"]["value"], 1), "roof", round(c["configs[4]"]["roofline"]["achieved"], 1))
    print("roof", round(d["roofline"]["achieved"], 1), {k: (round(v["ms"], 3), round(v["tflops"], 1)) for k, v in d["roofline"]["by_op"].items()})
except Exception as e:
    print("bench unreadable:", e)
PY
