#!/bin/bash
# round 2, call I (2 GPUs): data-parallel tests with the deferred teacher-generator step, overhead breakdown, bench N = 1, 2
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dp_nccl_gpu.py tests/test_graph_replay_gpu.py -q --timeout=200 > gpurun_out/pytest_i.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_i.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  " gpurun_out/pytest_i.log | cut -c1-300 | tail -20
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
   scripts/exp_dp_overhead.py > gpurun_out/dp_overhead_2gpu.txt 2> gpurun_out/dp_overhead_2gpu.err
echo "exit $?"; cat gpurun_out/dp_overhead_2gpu.txt; grep -v "^$" gpurun_out/dp_overhead_2gpu.err | grep -v OMP | tail -5 | cut -c1-300
timeout 240 python bench.py --steps 10 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_c2.json') if l.startswith('{')][-1]); print('N=1 value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1]); print('N=2 value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
grep -v "^$" gpurun_out/bench_2gpu.err | grep -v OMP | tail -3 | cut -c1-300
