"""Where do the data-parallel milliseconds go?  (torchrun --nproc-per-node N scripts/exp_dp_overhead.py)
Times the captured c2 iteration (graph segments) in four variants: full; without the all-reduces; without host
pacing; all-reduces alone.  Rank 0 prints one line per variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from gcc_b200 import factory, pix2pix
from gcc_b200.graph import GraphedIteration


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    opt = factory.make_opt("pix2pix", local, [], ngf=32, teacher_ngf=64, ndf=128, teacher_ndf=128, batch_size=32)
    model, teacher = factory.build_pair(opt)
    g = torch.Generator().manual_seed(5 + rank)
    tr = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in factory.synthetic_batch("pix2pix", 32, 256, g).items()}
    va = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in factory.synthetic_batch("pix2pix", 32, 256, g).items()}
    gi = GraphedIteration(model).capture(tr, va, warmup=2)

    def timed(n=12):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            gi.run(tr, va)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for _ in range(3):
        gi.run(tr, va)
    out = {"full (segments + all-reduce + pacing)": timed()}
    gi.pace = False
    out["no host pacing"] = timed()
    gi.pace = True
    real = pix2pix._allreduce_grads
    pix2pix._allreduce_grads = lambda arena, asynchronous=False: (pix2pix._DoneWork() if asynchronous else None)
    out["no all-reduce (segments + pacing only)"] = timed()
    gi.pace = False
    out["no all-reduce, no pacing"] = timed()
    pix2pix._allreduce_grads = real
    arenas = [b[0] for _, b in gi.segments if b is not None and b[1] != 'wait']
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for a in arenas:
            real(a)
    e1.record()
    torch.cuda.synchronize()
    out["the %d all-reduces alone (%.0f MB fp32)" % (len(arenas), sum(a.numel for a in arenas) * 4 / 1e6)] = e0.elapsed_time(e1) / 10
    if rank == 0:
        for k, v in out.items():
            print("N=%d  %-48s %.3f ms / iteration" % (world, k, v), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
