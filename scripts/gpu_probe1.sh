#!/bin/bash
# first GPU probe: tcgen05 conv + wgrad kernels against direct kernels / torch
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/gpu.txt
for f in decode gemm1x1 conv wgrad gram act; do
  echo "=== $f ===" | tee -a gpurun_out/probe1.log
  timeout 300 python scripts/probe_conv.py $f 2>&1 | tee -a gpurun_out/probe1.log
done
