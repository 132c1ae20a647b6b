#!/bin/bash
# effect of pacing the host on the N-GPU bench line ($1 = N)
N=${1:-4}
mkdir -p gpurun_out
P=29520
for w in 0 1; do
  P=$((P+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --skip_cpu_baseline --skip_roofline --warmup 3 --pace $w --steps 10 --watchdog_s 180 \
    > gpurun_out/warm_$w.json 2> gpurun_out/warm_$w.err
  tail -1 gpurun_out/warm_$w.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pace $w', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
