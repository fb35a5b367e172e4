#!/bin/bash
# 2-GPU check of the final bench (prefetch end-to-end leg on every rank, rank identity, real-size gather) + the 2-GPU test.
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "bench rc=$?" >> gpurun_out/n2_bench.err; tail -4 gpurun_out/n2_bench.err
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/n2_pytest.log 2>&1; tail -2 gpurun_out/n2_pytest.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n2_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["mode"][:30], "identity", d.get("rank_identity"), "tb", round(d["trajectory_batch"]["value"], 1),
      d["trajectory_batch"].get("bit_identical_to_1_rank"), "allgather", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d.get("allgather", {}).items() if k in ("ms", "busbw_GBps", "delivered")})
PY
