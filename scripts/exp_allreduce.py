"""All-reduce cost of the step's five gradient arenas (fp32 AVG), device-timed, max over ranks."""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = {"D_T": 11.04e6, "G_T": 54.4e6, "D_S": 11.04e6, "G_S": 13.7e6, "alpha": 1920}
bufs = {k: torch.randn(int(v), device="cuda") for k, v in sizes.items()}
for dt in (torch.float32, torch.bfloat16):
    tot = 0.0
    for k, b in bufs.items():
        x = b.to(dt)
        for _ in range(3): dist.all_reduce(x, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dist.all_reduce(x, op=dist.ReduceOp.AVG)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot += float(t)
        if dist.get_rank() == 0:
            print("%s %-6s %8.1f MB  %7.3f ms  algbw %6.1f GB/s" % (str(dt)[6:], k, x.numel() * x.element_size() / 1e6, float(t),
                                                                     x.numel() * x.element_size() / float(t) / 1e6), flush=True)
    if dist.get_rank() == 0:
        print("%s total per iteration %.3f ms" % (str(dt)[6:], tot), flush=True)
dist.destroy_process_group()
