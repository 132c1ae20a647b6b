#!/bin/bash
# round 2, call F (1 GPU): recalibrated tests, bench c2 (e2e check), kernel tables of four configs (torch.profiler),
# ncu --set full of the dominant kernels, ncu launch list of one c2 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cyclegan_parity_gpu.py tests/test_pipeline_prune_gpu.py tests/test_graph_replay_gpu.py -q -s --timeout=200 > gpurun_out/pytest_f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_f.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  " gpurun_out/pytest_f.log | cut -c1-300 | tail -20
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_c2.json') if l.startswith('{')][-1]); print('c2 value %.1f ms %.3f e2e %.1f cpu %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value']))"
for cfg in c2 c2_resnet cyclegan srgan; do
  timeout 200 python bench.py --config $cfg --profile gpurun_out/kernels_$cfg.txt > /dev/null 2> gpurun_out/profile_$cfg.err
  echo "== $cfg"; head -14 gpurun_out/kernels_$cfg.txt | cut -c1-150
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -c 10 -f -o gpurun_out/prof_gemm \
    python scripts/probe_one_conv.py 32 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 905 -c 1200 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --steps 1 --warmup 1 --graph 0 --skip_cpu_baseline --skip_e2e --skip_roofline > gpurun_out/ncu_c2.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_c2.csv > gpurun_out/launches_summary_c2.txt
head -30 gpurun_out/launches_summary_c2.txt
