#!/bin/bash
# ncu --set full capture of the norm kernels on the largest PatchGAN activation (scripts/exp_norm.py):
# skips the warm-up launches of norm_apply, captures its last timed launches and the first bwd reduce/apply pairs
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:norm_(apply|bwd)" -s 20 -c 9 \
  -o gpurun_out/ncu_norm -f python scripts/exp_norm.py > gpurun_out/ncu_norm.log 2>&1
ncu -i gpurun_out/ncu_norm.ncu-rep --page raw --csv > gpurun_out/ncu_norm_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_norm.log
