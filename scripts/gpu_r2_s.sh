#!/bin/bash
# round 2, call S (2 GPUs): data-parallel tests and the 2-GPU bench line at HEAD
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_nccl_gpu.py -q --timeout=300 > gpurun_out/pytest_s.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_s.log
grep -E "passed|failed|FAILED|Error|Timeout|^E  |exit" gpurun_out/pytest_s.log | cut -c1-300 | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_2gpu.json 2> gpurun_out/bench_c2_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_c2_2gpu.json") if l.startswith("{")][-1])
    print("2 GPUs: value %.1f  ms %.2f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("no 2-GPU line", e)
PY
tail -3 gpurun_out/bench_c2_2gpu.err | cut -c1-300
timeout 120 python bench.py --steps 10 --warmup 3 --skip_cpu_baseline > gpurun_out/bench_c2_1gpu_same_box.json 2> /dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_c2_1gpu_same_box.json') if l.startswith('{')][-1]); print('1 GPU same box: value %.1f ms %.2f' % (d['value'], d['ms_per_step']))"
