#!/bin/bash
# round 2, call Q (1 GPU): norm-backward reduction fused into the data-gradient epilogue, 64x64 weight-pack tiles,
# 32-bit col2im indices, tail split from 128 k-blocks on: tests + A/B + bench + launch list + per-shape trace + ncu full
# (historical: the fused reduction measured slower and was removed in the following commit, see
# profiles/r02_fused_norm_bwd_negative.txt; GCC_B200_FUSE_NORM_BWD no longer exists)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_q.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_q.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  |exit" gpurun_out/pytest_q.log | cut -c1-300 | tail -25
if grep -q failed gpurun_out/pytest_q.log; then
  GCC_B200_FUSE_NORM_BWD=0 timeout 600 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/pytest_q_nofuse.log 2>&1
  echo "without fusion:"; grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_q_nofuse.log | cut -c1-300 | tail -8
fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
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
python scripts/summarize_launches.py gpurun_out/launches_c2.csv > gpurun_out/launches_c2_summary.txt 2>&1; head -34 gpurun_out/launches_c2_summary.txt | cut -c1-130
timeout 200 python bench.py --trace 1 --graph 0 > /dev/null 2> gpurun_out/trace_c2.err
python scripts/summarize_trace.py gpurun_out/trace_c2.err > gpurun_out/gemm_trace_c2.txt 2>&1; head -14 gpurun_out/gemm_trace_c2.txt | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_persistent\|wgrad_gemm -c 10 -f -o gpurun_out/prof_gemm python scripts/probe_one_conv.py 32 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
