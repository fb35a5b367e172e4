#!/bin/bash
# Round-2 GPU pass M (8 GPUs): the bench line at N = 4 (rank identity, trajectory batches on every rank, real-size gather).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/p_gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/p_bench_n8.json 2> gpurun_out/p_bench_n8.err
echo "bench n8 rc=$?" >> gpurun_out/p_bench_n8.err; grep -E "rank identity|bit-identical|rc=|Error|error|real" gpurun_out/p_bench_n8.err | tail -8
python - <<'PY'
import json
d = json.load(open("gpurun_out/p_bench_n8.json"))
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single", round(d["single_trajectory"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1),
      d["rank_identity"]["bit_identical_to_1_rank"], d["trajectory_batch"]["bit_identical_to_1_rank"])
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["configs"]["configs[3]"]["allgather"].items() if k != "note"})
print("cfg4", round(d["configs"]["configs[4]"]["value"], 1))
PY
