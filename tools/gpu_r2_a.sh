#!/bin/bash
# Round-2 GPU pass A: full GPU test suite, bench line, launch list + DRAM bytes (step and configs[2] loop), --set full captures
# of the non-GEMM kernels the north_star names.  Run under gpurun from the repo root.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/a_launches_step.csv \
    python bench.py --profile-step > gpurun_out/a_ncu_step.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/a_launches_loop.csv \
    python bench.py --profile-step --profile-loop > gpurun_out/a_ncu_loop.log 2>&1
for k in splat_scatter splat_resolve vq_refine gn_apply_split; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 2 -f -o gpurun_out/a_full_$k \
      python bench.py --profile-step > gpurun_out/a_ncu_full_$k.log 2>&1
done
for k in inverse_warp tsdf_integrate tsdf_raycast; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 2 -f -o gpurun_out/a_full_$k \
      python bench.py --profile-step --profile-loop > gpurun_out/a_ncu_full_$k.log 2>&1
done
ls -la gpurun_out | tail -30
tail -5 gpurun_out/a_pytest.log
