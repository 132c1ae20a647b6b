#!/bin/bash
# ncu --set full capture of the norm kernels on the largest PatchGAN activation (scripts/exp_norm.py)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:norm_ -s 12 -c 8 \
  -o gpurun_out/ncu_norm -f python scripts/exp_norm.py > gpurun_out/ncu_norm.log 2>&1
tail -3 gpurun_out/ncu_norm.log
