#!/usr/bin/env python
"""Benchmark of the GCC cooperative-compression training step (BASELINE.json metric:
"GCC train images/sec (pix2pix 256^2, 1/2/4/8 B200); conv tensor-pipe %").

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle port on the host cores

One "step" = one GCC iteration of the reference's train loop (train.py:144-151):
set_input(train batch) -> optimize_parameters() -> set_input(val batch) -> clipping_mask_alpha()
-> optimizer_netD_arch(), on the configuration BASELINE.json quotes the metric on (configs[1]):
pix2pix GCC student (U-Net ngf 32, gated PatchGAN ndf 128) distilled online from the ngf-64 teacher,
256x256 synthetic cityscapes-shaped pairs, batch 32 per GPU, bf16 tensor-core math, dropout on
(scripts/pix2pix/train.sh).  Data parallel over N GPUs = N ranks x batch 32 (weak scaling), flat
gradient arenas all-reduced with NCCL before each optimizer step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    p.add_argument("--trace", type=int, default=0, help="debug: after warm-up run one eager step with per-GEMM-launch "
                   "event timing logged to stderr (GCCTRACE lines) and exit without a bench line")
    p.add_argument("--pace", type=int, default=1, help="N > 1: host waits for each iteration before queueing the next")
    p.add_argument("--clock_ms", type=int, default=200, help="nvidia-smi sampling interval during the timed region "
                   "(the profiling recipe's 200 ms; every sample briefly stalls the sampled GPU)")
    p.add_argument("--watchdog_s", type=int, default=600, help="abort the whole process after this many seconds")
    p.add_argument("--ngf", type=int, default=32)
    p.add_argument("--teacher_ngf", type=int, default=64)
    p.add_argument("--ndf", type=int, default=128)
    p.add_argument("--backbone", default="unet")
    p.add_argument("--no_dropout", action="store_true")
    p.add_argument("--cpu_iters", type=int, default=20, help="timed CPU-baseline iterations (batch 1), ~0.5 s each on 16 cores")
    p.add_argument("--skip_cpu_baseline", action="store_true")
    p.add_argument("--skip_e2e", action="store_true", help="profiling runs only")
    p.add_argument("--skip_roofline", action="store_true", help="profiling runs only")
    p.add_argument("--graph", type=int, default=1, help="1 (default): capture the whole iteration in a CUDA graph and replay it; 0: eager launches")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# fwd GMACs per sample (probe with a conv hook): used for the step-level FLOP accounting
def step_gmacs(ngf, teacher_ngf, ndf, backbone):
    unet = {64: 6.05, 32: 1.55}
    resnet = {64: 11.07, 32: 3.12}
    tab = resnet if backbone == "resnet" else unet
    g_t = tab.get(teacher_ngf, tab[64] * (teacher_ngf / 64.0) ** 2)
    g_s = tab.get(ngf, tab[64] * (ngf / 64.0) ** 2)
    d = 12.57 * (ndf / 128.0) ** 2
    gram = 8.0 if backbone != "resnet" else 9.4
    return 4 * g_t + 4 * g_s + 24 * d + gram + 0.6


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, interval_ms=200):
        self.idx, self.rows, self.proc, self.ms = gpu_index, [], None, int(interval_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.ms)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(args, iters, warmup=1):
    """images/s of the CPU oracle port (oracle/gcc_oracle.py, pinned to the reference by tests/golden) on
    the host cores: a bounded sample = `iters` GCC iterations at batch 1 of the same networks."""
    import torch
    from oracle import gcc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = O.Opt(ngf=args.ngf, ndf=args.ndf, teacher_ngf=args.teacher_ngf, teacher_ndf=128, backbone=args.backbone,
                no_dropout=args.no_dropout, direction="BtoA")
    S, T = O.build_pair(opt)
    times = []
    for it in range(warmup + iters):
        A, B = O.det_image("bench.A.%d" % it, 1, 3, 256, 256), O.det_image("bench.B.%d" % it, 1, 3, 256, 256)
        vA, vB = O.det_image("bench.vA.%d" % it, 1, 3, 256, 256), O.det_image("bench.vB.%d" % it, 1, 3, 256, 256)
        t0 = time.perf_counter()
        S.set_input(A, B)
        S.optimize_parameters()
        S.set_input(vA, vB)
        S.clipping_mask_alpha()
        S.optimizer_netD_arch()
        float(S.loss_D_arch)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"value": 1.0 / mean, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d GCC iterations at batch 1 (same nets: %s ngf %d / teacher %d / ndf %d, fp32, torch CPU "
                      "%d threads), %.2f s/iter" % (iters, args.backbone, args.ngf, args.teacher_ngf, args.ndf, cores, mean),
            "s_per_iter": mean}


def workload_name(args):
    return ("pix2pix GCC student (%s ngf %d, gated PatchGAN ndf %d) distilled online from ngf-%d teacher, 256x256, "
            "batch %d/GPU, dropout %s" % (args.backbone, args.ngf, args.ndf, args.teacher_ngf, args.batch,
                                          "off" if args.no_dropout else "on"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_oracle_rate(args, iters=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    line = {"metric": "GCC train images/sec (pix2pix 256^2)", "value": base["value"], "unit": "images/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * base["s_per_iter"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "note": "CPU oracle port of the reference step, batch 1 per step"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ B200 arm
def dominant_kernel_roofline(batch, pk):
    """CUDA-event timing of the dominant kernel alone: PatchGAN 512->1024 k4 s1 fprop (conv_gemm_persistent_kernel<256,4>),
    L2 flushed between launches.  Algorithmic FLOPs = 2 * (batch*31*31) * 1024 * (512*16)."""
    import torch
    from gcc_b200 import _lib
    n, h, w, cin, cout, k = batch, 32, 32, 512, 1024, 4
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, k * k, cin, device="cuda") * 0.02).to(torch.bfloat16)
    y = torch.empty(n, 31, 31, cout, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        _lib.call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cin, wt.data_ptr(), cout, k * k, cin, None, y.data_ptr(),
                  31, 31, cout, 0, 0, k, k, 1, 1, 0, 0.0, 0, None, 0, None, 0, st)

    for _ in range(3):
        launch()
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    flops = 2.0 * n * 31 * 31 * cout * cin * k * k
    ach = flops / (ms * 1e-3) / 1e12
    # DRAM traffic of this exact launch from the committed `ncu --set full` capture
    # (profiles/r01_ncu_full_gemm_v2_persistent.csv: dram__bytes_read.sum 55.9 MB + dram__bytes_write.sum 35.1 MB at
    # batch 32; algorithmic unique bytes = 33.5 MB input + 16.8 MB weights + 63.0 MB output = 113 MB)
    traffic = 91.1e6 if n == 32 else None
    return {"bound": "tensor", "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_burst"], "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
            "ncu_tensor_pipe_active_pct": 89.6,
            "kernel": "conv_gemm_persistent_kernel<256,4>: PatchGAN 512->1024 k4 s1 fprop, M=%d N=1024 K=8192" % (n * 961),
            "avg_launch_ms": ms, "peak_source": pk["source"] + ", burst (kernel timed alone)"}


def _watchdog(seconds):
    """A hung collective or kernel must not hold the GPU box: give up loudly after `seconds`."""
    import threading

    def bark():
        sys.stderr.write("bench.py: watchdog expired after %d s, aborting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, bark)
    t.daemon = True
    t.start()
    return t


def run_b200(args):
    import torch
    _watchdog(args.watchdog_s)
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gcc_b200 import _lib, options
    from gcc_b200.pix2pix import Pix2PixModel, build_teacher
    argv = ["--dataroot", "./database/cityscapes/", "--model", "pix2pix", "--ngf", str(args.ngf), "--ndf", str(args.ndf),
            "--teacher_ngf", str(args.teacher_ngf), "--darts_discriminator", "--online_distillation",
            "--lambda_content", "50", "--lambda_gram", "1e4", "--gpu_ids", str(local), "--backbone", args.backbone,
            "--batch_size", str(args.batch)]
    if args.no_dropout:
        argv.append("--no_dropout")
    opt = options.parse(argv)
    torch.manual_seed(1234 + rank)
    model = Pix2PixModel(opt)
    build_teacher(model, opt)
    model.model_train()
    if world > 1:  # replicate the initial weights of rank 0
        for m in (model, model.teacher_model):
            for a in (m.arena_G, m.arena_D, m.arena_A):
                if a is not None:
                    dist.broadcast(a.P, 0)
                    a.mark_dirty()

    B = args.batch
    g = torch.Generator().manual_seed(99 + rank)
    nbatch = 2
    host = [{k: torch.rand(B, 3, 256, 256, generator=g).mul_(2).sub_(1).pin_memory() for k in ("A", "B", "vA", "vB")}
            for _ in range(nbatch)]
    devb = [{k: v.cuda() for k, v in h.items()} for h in host]

    graphed = None
    pace_ev = torch.cuda.Event()

    def step(d, read_losses):
        if graphed is not None:
            graphed.run(d)
            if read_losses:
                return model.get_current_losses()
            if world > 1 and args.pace:
                # data parallel: keep the host at most one iteration ahead of the device.  Measured on 4 and 8 GPUs:
                # an unpaced host (all K iterations' graph segments and collectives queued at once) runs 1-4 ms per
                # iteration SLOWER than the end-to-end loop, whose loss read-back paces it.
                pace_ev.record()
                pace_ev.synchronize()
            return None
        model.set_input({"A": d["A"], "B": d["B"], "A_paths": "", "B_paths": ""})
        model.optimize_parameters()
        model.set_input({"A": d["vA"], "B": d["vB"], "A_paths": "", "B_paths": ""})
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
        if read_losses:
            return model.get_current_losses()  # float() of every loss: device -> host reads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, data, read_losses):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.lib().gcc_launch_count()
        torch.arange(5, device="cuda").cumsum(0)  # marker kernel: scripts/summarize_launches.py cuts the ncu list here
        e0.record()
        for i in range(nsteps):
            step(data[i % nbatch], read_losses)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _lib.lib().gcc_launch_count() - l0

    if args.trace:
        for i in range(3):
            step(devb[i % nbatch], False)
        torch.cuda.synchronize()
        _lib.lib().gcc_debug_set_flags(32)
        step(devb[0], False)
        torch.cuda.synchronize()
        _lib.lib().gcc_debug_set_flags(0)
        return
    sampler = ClockSampler(local, args.clock_ms)
    if rank == 0:
        sampler.start()  # sampled from the warm-up on: the GPU is under the same load as in the timed region
    launches_per_step = None
    if args.graph:
        from gcc_b200.graph import GraphedIteration
        step(devb[0], False)
        l0 = _lib.lib().gcc_launch_count()
        step(devb[1], False)
        launches_per_step = _lib.lib().gcc_launch_count() - l0  # the captured graph replays exactly these launches
        graphed = GraphedIteration(model, B).capture(devb[0], warmup=1)
    for i in range(args.warmup):
        step(devb[i % nbatch], False)
    ms, launches = timed(args.steps, devb, False)
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if args.skip_e2e:
        ms_e2e = float("nan")
    else:
        step(host[0], True)  # warm the host-input path once
        ms_e2e, _ = timed(args.steps, host, True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    imgs = world * B * args.steps
    value = imgs / (ms * 1e-3)
    gmacs = step_gmacs(args.ngf, args.teacher_ngf, args.ndf, args.backbone)
    step_tflops = value / world * gmacs * 2e9 / 1e12
    line = {
        "metric": "GCC train images/sec (pix2pix 256^2)", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": world * B, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (activations of 15 net passes at batch %d, several GB) >> 126 MB L2; "
                         "two alternating input batches" % B,
                   "cuda_graph": bool(args.graph),
                   "algorithmic_gmac_per_image": gmacs,
                   "step_tensor_tflops_per_gpu": step_tflops,
                   "step_frac_of_sustained_bf16_peak": step_tflops / pk["bf16_sustained"]},
        "clocks": clocks,
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": 4 * B * 3 * 256 * 256 * 4, "d2h_bytes_per_step": 4 * len(model.loss_names),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
    }
    if world == 1:
        if not args.skip_roofline:
            line["roofline"] = dominant_kernel_roofline(B, pk)
        if not args.skip_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_rate(args, iters=args.cpu_iters)
    else:
        dist.destroy_process_group()
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
