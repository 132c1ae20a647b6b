#!/bin/bash
# one `ncu --set full` capture of the dominant kernels (1 GPU, few launches); reports come back in gpurun_out/
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm -c 12 -f -o gpurun_out/prof_gemm \
    python scripts/probe_one_conv.py 32 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
