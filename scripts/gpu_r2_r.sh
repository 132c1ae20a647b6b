#!/bin/bash
# round 2, call R (1 GPU): validation at HEAD -- GPU suite, smoke, bench lines of all configs, reference arm, ncu launch
# list of one c2 step, per-shape GEMM trace, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_r.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  |exit" gpurun_out/pytest_r.log | cut -c1-300 | tail -25
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(sys.argv[1], "value %.1f %s  ms %.2f  e2e %.1f  frac %.3f  cpu %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; line c2 gpurun_out/bench_c2.json
timeout 120 python bench.py --steps 20 --warmup 5 --skip_cpu_baseline > gpurun_out/bench_c2_20steps.json 2> /dev/null; line c2_20steps gpurun_out/bench_c2_20steps.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 1 --graph 0 --skip_cpu_baseline --skip_e2e --skip_roofline > gpurun_out/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_c2.csv > gpurun_out/launches_c2_summary.txt 2>&1; head -24 gpurun_out/launches_c2_summary.txt | cut -c1-130
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_persistent\|wgrad_gemm -c 10 -f -o gpurun_out/prof_gemm python scripts/probe_one_conv.py 32 2 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
timeout 200 python bench.py --trace 1 --graph 0 > /dev/null 2> gpurun_out/trace_c2.err
python scripts/summarize_trace.py gpurun_out/trace_c2.err > gpurun_out/gemm_trace_c2.txt 2>&1; head -8 gpurun_out/gemm_trace_c2.txt | cut -c1-150
timeout 200 python scripts/exp_ab_c2.py c2 15 > gpurun_out/ab_c2.txt 2>&1; cat gpurun_out/ab_c2.txt | cut -c1-150
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-300
for cfg in c2_pruned c2_resnet cyclegan srgan sagan; do
  timeout 200 python bench.py --config $cfg --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  line $cfg gpurun_out/bench_$cfg.json
done
