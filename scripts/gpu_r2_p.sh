#!/bin/bash
# round 2, call P (1 GPU): instruction clean-up of the GEMM issue loops, parallel tail finalize, pooled norm scratch: tests + A/B + bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_p.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_p.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  |exit" gpurun_out/pytest_p.log | cut -c1-300 | tail -25
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 200 python scripts/exp_conv_bound.py > gpurun_out/conv_bound.txt 2>&1; cat gpurun_out/conv_bound.txt | cut -c1-150
timeout 400 python scripts/exp_ab_c2.py c2 15 > gpurun_out/ab_c2.txt 2>&1; cat gpurun_out/ab_c2.txt | cut -c1-150
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_c2.json") if l.startswith("{")][-1])
    print("c2 value %.1f  ms %.2f  e2e %.1f  frac %.3f  cpu %s  launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"), d["gpu_launches"]))
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/bench_c2.err | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 1 --graph 0 --skip_cpu_baseline --skip_e2e --skip_roofline > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_c2.csv > gpurun_out/launches_c2_summary.txt 2>&1; head -30 gpurun_out/launches_c2_summary.txt | cut -c1-130
