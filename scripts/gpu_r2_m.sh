#!/bin/bash
# round 2, call M (1 GPU): norm grid-cap experiment, slab sliding-window kernels (tests + cyclegan / c2_resnet lines)
mkdir -p gpurun_out
for w in 16 4 2; do GCC_B200_NORM_UNROLL=4 GCC_B200_NORM_WAVES=$w python scripts/exp_norm_unroll.py | grep -v reduce; done > gpurun_out/norm_waves.txt 2>&1
cat gpurun_out/norm_waves.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_cyclegan_parity_gpu.py tests/test_step_parity_gpu.py -q --timeout=200 -k "slab or cyclegan or resnet" > gpurun_out/pytest_m.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_m.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  " gpurun_out/pytest_m.log | cut -c1-300 | tail -20
for cfg in cyclegan c2_resnet; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_$cfg.json") if l.startswith("{")][-1])
    print("$cfg value %.1f img/s  ms %.2f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$cfg no line", e)
PY
  tail -2 gpurun_out/bench_$cfg.err
done
timeout 200 python bench.py --config cyclegan --profile gpurun_out/kernels_cyclegan.txt > /dev/null 2> gpurun_out/profile_cyclegan.err
head -10 gpurun_out/kernels_cyclegan.txt | cut -c1-150
