#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do ( cd tools && SGAM_TC_EPI8=$f timeout 300 python short_k_sweep.py ) > gpurun_out/z2_shortk_epi$f.txt 2>&1; cat gpurun_out/z2_shortk_epi$f.txt | tail -5; done
