#!/bin/bash
# round 2, call G (1 GPU): slab kernels -- unit tests, the MobileResNet / CycleGAN parity and replay tests, bench lines
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_kernels_gpu.py tests/test_cyclegan_parity_gpu.py tests/test_step_parity_gpu.py tests/test_graph_replay_gpu.py -q --timeout=200 -k "slab or cyclegan or resnet or dw" > gpurun_out/pytest_g.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_g.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  " gpurun_out/pytest_g.log | cut -c1-300 | tail -30
for cfg in cyclegan c2_resnet; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "bench $cfg exit $?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_$cfg.json") if l.startswith("{")][-1])
    print("$cfg value %.1f img/s  ms %.2f  e2e %.1f  launches/step %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print("no line", e)
PY
  tail -3 gpurun_out/bench_$cfg.err
done
timeout 200 python bench.py --config cyclegan --profile gpurun_out/kernels_cyclegan.txt > /dev/null 2> gpurun_out/profile_cyclegan.err
head -16 gpurun_out/kernels_cyclegan.txt | cut -c1-150
