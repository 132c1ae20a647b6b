#!/bin/bash
# round 2, call H (2 GPUs): where the data-parallel milliseconds go
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
   scripts/exp_dp_overhead.py > gpurun_out/dp_overhead_2gpu.txt 2> gpurun_out/dp_overhead_2gpu.err
echo "exit $?"; cat gpurun_out/dp_overhead_2gpu.txt; grep -v "^$" gpurun_out/dp_overhead_2gpu.err | tail -5 | cut -c1-300
