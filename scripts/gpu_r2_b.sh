#!/bin/bash
# round 2, call B: full GPU test suite (no -x) + ncu launch lists of one step of every config
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_dp_nccl_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|dW rel|Error" gpurun_out/pytest_gpu.log | tail -40
for cfg in c2 c2_resnet cyclegan srgan sagan; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$cfg.csv \
    python bench.py --config $cfg --steps 1 --warmup 1 --graph 0 --skip_cpu_baseline --skip_e2e --skip_roofline > gpurun_out/ncu_$cfg.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$cfg.csv > gpurun_out/launches_summary_$cfg.txt
  echo "== $cfg"; head -24 gpurun_out/launches_summary_$cfg.txt
  rm -f gpurun_out/launches_$cfg.csv
done
