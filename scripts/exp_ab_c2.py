"""A/B harness on the c2 bench workload in ONE process: the model is built once, every variant re-captures the
iteration as a CUDA graph with its switches set (they are read on the host at launch = capture time) and times N replays.
Variants: tail-wave split of the conv kernel (debug bit 11 = off).  Usage: python scripts/exp_ab_c2.py [config] [replays]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from gcc_b200 import _lib, factory
from gcc_b200.graph import GraphedIteration

config = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 15
c = bench.CONFIGS[config]
torch.cuda.set_device(0)
opt = factory.make_opt(c["model"], 0, [], batch_size=c["batch"], **c["opt"])
torch.manual_seed(1234)
model, teacher = factory.build_pair(opt, c["cfgs"])
g = torch.Generator().manual_seed(99)
B = c["batch"]
devb = []
for _ in range(2):
    pair = (factory.synthetic_batch(c["model"], B, c["size"], g, pin=False), factory.synthetic_batch(c["model"], B, c["size"], g, pin=False))
    devb.append(tuple({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for d in pair))
L = _lib.lib()
for i in range(2):
    factory.run_iteration(model, devb[i][0], devb[i][1])
torch.cuda.synchronize()

VARIANTS = [("baseline (no tail split)", 2048), ("tail split", 0), ("baseline again", 2048), ("tail split again", 0)]
for name, flags in VARIANTS:
    L.gcc_debug_set_flags(flags)
    graphed = GraphedIteration(model).capture(devb[0][0], devb[0][1], warmup=1)
    for i in range(3):
        graphed.run(devb[i % 2][0], devb[i % 2][1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        graphed.run(devb[i % 2][0], devb[i % 2][1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    losses = model.get_current_losses()
    print("%-52s %7.3f ms / iteration  %7.1f images/s   G_L1 %.4f" % (name, ms, B / ms * 1e3, losses.get("G_L1", float("nan"))), flush=True)
    del graphed
    torch.cuda.empty_cache()
L.gcc_debug_set_flags(0)
