#!/bin/bash
mkdir -p gpurun_out
for pr in 0 1 2; do ( cd tools && SGAM_GNCONV_PROBE=$pr timeout 300 python gnconv_probe.py ) 2>&1 | tail -1; done | tee gpurun_out/af_probe.txt
