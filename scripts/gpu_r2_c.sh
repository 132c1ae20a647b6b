#!/bin/bash
# round 2, call C (2 GPUs): data-parallel tests, NCCL-in-graph vs graph segments, PDL on/off at N = 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dp_nccl_gpu.py tests/test_grad_terms_gpu.py tests/test_sagan_parity_gpu.py tests/test_graph_replay_gpu.py tests/test_model_surface_gpu.py tests/test_cyclegan_parity_gpu.py tests/test_step_parity_gpu.py -q -s > gpurun_out/pytest_dp.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dp.log
grep -E "passed|failed|FAILED|Error|condition|both_vs|dW rel" gpurun_out/pytest_dp.log | tail -50
for pdl in 1 0; do
  GCC_B200_PDL=$pdl timeout 400 python bench.py --steps 10 --warmup 3 --skip_cpu_baseline --skip_roofline > gpurun_out/bench_c2_pdl$pdl.json 2> gpurun_out/bench_c2_pdl$pdl.err
  echo "pdl=$pdl exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_pdl$pdl.json')); print('value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
  tail -3 gpurun_out/bench_c2_pdl$pdl.err
done
for cap in 1 0; do
  GCC_B200_CAPTURE_NCCL=$cap timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_cap$cap.json 2> gpurun_out/bench_2gpu_cap$cap.err
  echo "2gpu capture=$cap exit $?"; tail -c 900 gpurun_out/bench_2gpu_cap$cap.json; grep -v "^$" gpurun_out/bench_2gpu_cap$cap.err | tail -4
done
