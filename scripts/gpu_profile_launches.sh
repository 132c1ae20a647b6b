#!/bin/bash
# ncu launch list (device time per launch; cold-cache, serialised: compare SHARES) of one bench step
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --graph 0 --skip_cpu_baseline --skip_e2e --skip_roofline > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt
tail -40 gpurun_out/launches_summary.txt
