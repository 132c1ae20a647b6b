#!/bin/bash
# round 2, call N (1 GPU): reduce-kernel variants, then the full validation at HEAD: GPU suite, smoke, bench lines
mkdir -p gpurun_out
for f in 0 4 8; do GCC_B200_NORM_REDUCE_FLAT=$f timeout 120 python scripts/exp_norm_unroll.py; done > gpurun_out/norm_reduce_variants.txt 2>&1
grep -E "reduce \(|bwd|apply" gpurun_out/norm_reduce_variants.txt | grep -v "fwd" | cut -c1-120
GCC_B200_NORM_REDUCE_FLAT=4 timeout 300 python -m pytest tests/test_kernels_gpu.py -q --timeout=100 -k "norm" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout=200 > gpurun_out/pytest_n.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_n.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  |exit" gpurun_out/pytest_n.log | cut -c1-300 | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(sys.argv[1], "value %.1f %s  ms %.2f  e2e %.1f  frac %.3f  cpu %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; line default gpurun_out/bench_default.json
GCC_B200_NORM_REDUCE_FLAT=4 timeout 300 python bench.py --skip_cpu_baseline > gpurun_out/bench_c2_flat4.json 2> /dev/null; line c2_flat4 gpurun_out/bench_c2_flat4.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-400
for cfg in c2_pruned c2_resnet cyclegan srgan sagan; do
  timeout 300 python bench.py --config $cfg --skip_cpu_baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  line $cfg gpurun_out/bench_$cfg.json
done
