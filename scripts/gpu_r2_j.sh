#!/bin/bash
# round 2, call J (8 GPUs): one data-parallel bench line at N = 8 (and N = 4 on the same box)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n \
     bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "N=$n exit $?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${n}gpu.json') if l.startswith('{')][-1]); print('N=$n value %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
  grep -v "^$" gpurun_out/bench_${n}gpu.err | grep -v "OMP\|\*\*\*" | tail -3 | cut -c1-300
done
timeout 120 python bench.py --steps 10 --warmup 3 --skip_cpu_baseline --skip_roofline > gpurun_out/bench_c2_box8.json 2> /dev/null
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_c2_box8.json') if l.startswith('{')][-1]); print('N=1 (same box) value %.1f ms %.3f' % (d['value'], d['ms_per_step']))"
