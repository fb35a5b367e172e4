#!/bin/bash
# ncu --set full of the row-reuse kernel (one 128 -> 128 3x3 conv at 256 x 256, 8 trajectories) for profiles/.
mkdir -p gpurun_out
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'tc_gemm_swaprow_kernel' -s 2 -c 1 -f -o gpurun_out/ag_swaprow python bench.py --profile-step > gpurun_out/ag_ncu.log 2>&1
ncu -i gpurun_out/ag_swaprow.ncu-rep --page raw --csv > gpurun_out/ag_swaprow_raw.csv 2>/dev/null
python tools/ncu_full_summary.py gpurun_out/ag_swaprow.ncu-rep | cut -c1-260
