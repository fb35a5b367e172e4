#!/bin/bash
# Round-2 GPU pass W: two / four half-batches on parallel graph branches against one full batch.
mkdir -p gpurun_out
( timeout 600 python tools/dual_stream_experiment.py 8 ) > gpurun_out/w_dual_stream.txt 2>&1; echo "rc=$?" >> gpurun_out/w_dual_stream.txt; tail -12 gpurun_out/w_dual_stream.txt
( timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -x -q ) > gpurun_out/w_pytest_attn.log 2>&1; tail -3 gpurun_out/w_pytest_attn.log
