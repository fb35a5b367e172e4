#!/bin/bash
# 8-GPU run of the final bench: every rank runs the headline step, the prefetch end-to-end leg, the 512x512 100-step loop and the
# lock-step trajectory batch; rank identity asserted; the final-map all-gather at configs[3]'s real size.
mkdir -p gpurun_out
N=${1:-8}
( timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "bench rc=$?" >> gpurun_out/n${N}_bench.err; grep -h "rank identity\|trajectory batch\|bench rc" gpurun_out/n${N}_bench.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/n{n}_bench.json").read().strip().splitlines()[-1])
ag = d.get("allgather") or d.get("trajectory_batch", {}).get("allgather") or {}
print("n", d["n_gpus"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "tb", round(d["trajectory_batch"]["value"], 1),
      "cfg4", round(d["configs"]["configs[4]"]["value"], 1), "gather", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in ag.items() if k in ("ms", "busbw_GBps", "bytes_per_rank", "delivered")})
print([k for k in d.keys()])
PY
