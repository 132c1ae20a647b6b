#!/bin/bash
# run the GPU test-suite on the B200 box, keep logs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s "$@" 2>&1 | tail -150 | tee gpurun_out/pytest_gpu.log
