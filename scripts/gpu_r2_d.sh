#!/bin/bash
# round 2, call D (1 GPU): full GPU suite with per-test timeouts, then one bench line per config
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q -s --timeout=240 --deselect tests/test_dp_nccl_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|condition|dW rel|Timeout" gpurun_out/pytest_gpu.log | tail -40
for cfg in c2 c2_resnet cyclegan srgan sagan; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit $?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$cfg.json"))
    print("$cfg value %.1f img/s  ms %.2f  e2e %.1f  roofline %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("no line", e)
PY
  tail -3 gpurun_out/bench_$cfg.err
done
