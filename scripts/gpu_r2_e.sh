#!/bin/bash
# round 2, call E (2 GPUs): whole GPU suite incl. the NCCL tests (per-test timeouts), bench at N = 1 / 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout=200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|condition|Timeout|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -40
timeout 240 python bench.py --steps 10 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print('c2 value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "2gpu segments exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('2gpu value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
grep -v "^$" gpurun_out/bench_2gpu.err | tail -3 | cut -c1-300
GCC_B200_CAPTURE_NCCL=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
   bench.py --gpus 2 --steps 10 --warmup 3 --watchdog_s 120 > gpurun_out/bench_2gpu_cap.json 2> gpurun_out/bench_2gpu_cap.err
echo "2gpu captured-nccl exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_cap.json')); print('2gpu-cap value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
grep -v "^$" gpurun_out/bench_2gpu_cap.err | tail -3 | cut -c1-300
