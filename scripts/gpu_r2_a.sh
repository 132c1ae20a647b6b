#!/bin/bash
# round 2, call A: full GPU test suite + first bench lines of every config
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_dp_nccl_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; tail -c 1500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
for cfg in c2_pruned c2_resnet cyclegan srgan sagan; do
  timeout 420 python bench.py --config $cfg --steps 5 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit $?"; tail -c 600 gpurun_out/bench_$cfg.json; tail -3 gpurun_out/bench_$cfg.err
done
