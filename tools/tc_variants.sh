#!/bin/bash
# compare tc_gemm kernel variants (development aid)
for v in ${VARIANTS:-0 1 2 3}; do
  echo "=== SGAM_TC_VARIANT=$v ==="
  SGAM_TC_VARIANT=$v timeout -s KILL 150 python tools/tc_probe.py 2>&1 | grep -E "rel=|nsplit=|algorithmic" | tail -24
  SGAM_TC_VARIANT=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], 'gemm ms', d['kernels']['tc_gemm_ms_per_step'], 'single', d['single_trajectory']['value'])"
done
