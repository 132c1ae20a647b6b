#!/bin/bash
# round 2, call K (1 GPU): per-shape GEMM traces of the SRGAN and CycleGAN iterations
mkdir -p gpurun_out
for cfg in srgan cyclegan; do
  timeout 300 python bench.py --config $cfg --trace 1 --graph 0 > /dev/null 2> gpurun_out/trace_$cfg.log
  python scripts/summarize_trace.py gpurun_out/trace_$cfg.log > gpurun_out/trace_summary_$cfg.txt
  echo "== $cfg"; head -28 gpurun_out/trace_summary_$cfg.txt | cut -c1-170
done
