#!/bin/bash
# round 2, call L (1 GPU): full suite + one bench line per config after the channel-padding change (multiples of 16)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=200 --deselect tests/test_dp_nccl_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -30
for cfg in c2 c2_pruned c2_resnet cyclegan srgan sagan; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_$cfg.json") if l.startswith("{")][-1])
    print("$cfg value %.1f img/s  ms %.2f  e2e %.1f  launches/step %d  roofline %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"], d["roofline"]["frac"]))
except Exception as e:
    print("$cfg no line", e)
PY
  tail -2 gpurun_out/bench_$cfg.err
done
