#!/bin/bash
# one development cycle on the GPU box: tests ($1 = pytest targets), bench line, per-launch GEMM trace
mkdir -p gpurun_out
T=${1:-"tests/test_kernels_gpu.py tests/test_step_parity_gpu.py tests/test_model_surface_gpu.py"}
timeout 900 python -m pytest $T -q -m gpu > gpurun_out/cycle_tests.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|assert|Error" gpurun_out/cycle_tests.log | grep -v Warning | tail -15
timeout 300 python bench.py --skip_cpu_baseline --steps 10 2>gpurun_out/cycle_bench.err | tail -1 > gpurun_out/cycle_bench.json
cut -c1-200 gpurun_out/cycle_bench.json
if [ "$2" != "notrace" ]; then
  timeout 200 python bench.py --trace 1 --graph 0 2> gpurun_out/trace.log
  python scripts/summarize_trace.py gpurun_out/trace.log > gpurun_out/trace_summary.txt
  head -3 gpurun_out/trace_summary.txt
fi
