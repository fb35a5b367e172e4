#!/bin/bash
# Round-2 GPU pass H (2 GPUs): GroupNorm-apply changes (tests), two-rank NCCL test, two-rank bench (rank identity, real-size gather).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/h_gpus.txt
( timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_configs.py tests/test_gpu_multi.py tests/test_gpu_scene_batch.py -x -q ) > gpurun_out/h_unit.log 2>&1
echo "unit rc=$?" >> gpurun_out/h_unit.log; tail -6 gpurun_out/h_unit.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/h_bench_n2.err; grep -E "rank identity|bit-identical|rc=|Error|error" gpurun_out/h_bench_n2.err | tail -8
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err
echo "bench n1 rc=$?" >> gpurun_out/h_bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/h_bench_n1.json", "gpurun_out/h_bench_n2.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1),
              "tb", round(d["trajectory_batch"]["value"], 1), d.get("rank_identity", {}).get("bit_identical_to_1_rank"), d["trajectory_batch"].get("bit_identical_to_1_rank"),
              "allgather", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.get("configs", {}).get("configs[3]", {}).get("allgather", {}).items() if k != "note"})
        print("   cfg4", round(d["configs"]["configs[4]"]["value"], 1), "cfg2", round(d["configs"].get("configs[2]", {}).get("value", 0), 1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
