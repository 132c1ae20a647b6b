#!/bin/bash
# round 2, calls T (N GPUs, N = $1): the N-GPU bench line at HEAD and the single-GPU line of the same box
N=${1:-4}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_${N}gpu.json 2> gpurun_out/bench_c2_${N}gpu.err
timeout 100 python bench.py --steps 10 --warmup 3 --skip_cpu_baseline --skip_roofline > gpurun_out/bench_c2_1gpu_same_box_as_${N}gpu.json 2> /dev/null
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in ("gpurun_out/bench_c2_%sgpu.json" % n, "gpurun_out/bench_c2_1gpu_same_box_as_%sgpu.json" % n):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value %.1f  ms %.2f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    except Exception as e:
        print(f, "no line", e)
PY
tail -2 gpurun_out/bench_c2_${N}gpu.err | cut -c1-200
