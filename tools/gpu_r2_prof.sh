#!/bin/bash
# ncu --set full (+ source) of one 512->512 3x3 @16x16 split-K launch (1-CTA kernel) and of the fused q/k/v projection + the
# attention output projection (pair kernel, eight epilogue warps); tests of the extended GroupNorm shapes.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q ) > gpurun_out/p_pytest.log 2>&1; tail -2 gpurun_out/p_pytest.log
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'tc_gemm_kernel' -s 8 -c 1 -f -o gpurun_out/p_tc1_512 python bench.py --profile-step > gpurun_out/p_ncu1.log 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'tc_gemm2_kernel' -s 12 -c 2 -f -o gpurun_out/p_tc2_short python bench.py --profile-step > gpurun_out/p_ncu2.log 2>&1
ls -la gpurun_out/p_*.ncu-rep
