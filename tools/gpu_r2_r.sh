#!/bin/bash
# Round-2 GPU pass R: deferred split-K (fused finish): unit test first, then the full suite, bench, batch-1 launch list.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "deferred or split_k or statistics" ) > gpurun_out/r_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/r_unit.log; tail -12 gpurun_out/r_unit.log
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log; tail -8 gpurun_out/r_pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench rc=$?" >> gpurun_out/r_bench.err
SGAM_DEFER_SPLITK=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r_bench_nodefer.json 2> gpurun_out/r_bench_nodefer.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r_launches_b1.csv python bench.py --batch 1 --profile-step > gpurun_out/r_ncu_b1.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r_bench.json", "gpurun_out/r_bench_nodefer.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "loop", round(d["scene_loop"]["value"], 1),
              "launches", d["gpu_launches_per_step"], d["single_trajectory"]["gpu_launches_per_frame"])
        if "configs" in d:
            c = d["configs"]
            print("   cfg2", round(c["configs[2]"]["value"], 1), "once", round(c["configs[2]"]["integrate_once"]["value"], 1), "cfg4", round(c["configs[4]"]["value"], 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
