#!/usr/bin/env python
"""Benchmark of the GCC cooperative-compression training step (BASELINE.json metric:
"GCC train images/sec (pix2pix 256^2, 1/2/4/8 B200); conv tensor-pipe %").

    python bench.py --gpus N --steps K --warmup W [--config NAME]           # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ... [--config NAME] # the CPU oracle port on the host cores

One "step" = one GCC iteration of the reference's train loop (train.py:144-151):
set_input(train batch) -> optimize_parameters() -> set_input(val batch) -> clipping_mask_alpha()
-> optimizer_netD_arch().  Workloads (--config):

  c2 (default)  BASELINE configs[1]: pix2pix GCC student (U-Net ngf 32, gated PatchGAN ndf 128) distilled online from
                the ngf-64 teacher, 256x256, batch 32 per GPU, bf16, dropout on (scripts/pix2pix/train.sh)
  c2_pruned     the same with the literal pruned student of SURVEY.md 7/8d (odd widths 37, 65, 143 ...)
  c2_resnet     the same with the MobileResNet-9 backbone (student ngf 32, teacher ngf 64)
  cyclegan      configs[2]: CycleGAN 256x256, student ngf 24 with the reference's shipped channel lists
                (utils/prune_util.py:120-121), teacher ngf 64, ndf 64, batch 8 per GPU
  srgan         configs[3]: SRGAN 4x, 96 -> 384 (HR crops of 384), student ngf 24 / teacher 64, batch 16 per GPU
  sagan         configs[4]: SAGAN 64x64 (the only size the reference's forward supports), ngf 48 / teacher 64, batch 64

Data parallel over N GPUs = N ranks x the per-GPU batch (weak scaling), flat gradient arenas all-reduced with NCCL
before each optimizer step (captured inside the iteration's CUDA graph).
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

PRUNED_F = [32, 37, 65, 143, 144, 136, 134, 256, 120, 127, 128, 138, 62, 33, 13]
PRUNED_C = [32, 37, 65, 143, 144, 136, 134, 256, 254, 263, 272, 281, 127, 70, 45]
CYC_A = [24, 48, 86, 72, 86, 47, 86, 44, 86, 43, 86, 43, 86, 29, 86, 30, 86, 37, 86, 36, 86, 48, 24]
CYC_B = [24, 48, 96, 91, 96, 73, 96, 62, 96, 61, 96, 74, 96, 54, 96, 51, 96, 58, 96, 81, 96, 48, 24]

CONFIGS = {
    "c2": dict(model="pix2pix", opt=dict(ngf=32, teacher_ngf=64, ndf=128, teacher_ndf=128, backbone="unet"),
               cfgs=(None, None), batch=32, size=256),
    "c2_pruned": dict(model="pix2pix", opt=dict(ngf=32, teacher_ngf=64, ndf=128, teacher_ndf=128, backbone="unet"),
                      cfgs=(PRUNED_F, PRUNED_C), batch=32, size=256),
    "c2_resnet": dict(model="pix2pix", opt=dict(ngf=32, teacher_ngf=64, ndf=128, teacher_ndf=128, backbone="resnet"),
                      cfgs=(None, None), batch=32, size=256),
    "cyclegan": dict(model="cyclegan", opt=dict(ngf=24, teacher_ngf=64, ndf=64, teacher_ndf=64), cfgs=(CYC_A, CYC_B),
                     batch=8, size=256),
    "srgan": dict(model="srgan", opt=dict(ngf=24, teacher_ngf=64, ndf=128, teacher_ndf=64, image_size=384),
                  cfgs=(None, None), batch=16, size=384),
    "sagan": dict(model="sagan", opt=dict(ngf=48, teacher_ngf=64, ndf=64, teacher_ndf=64), cfgs=(None, None), batch=64,
                  size=64),
}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    p.add_argument("--batch", type=int, default=0, help="images per GPU per step (0 = the config's)")
    p.add_argument("--trace", type=int, default=0, help="debug: after warm-up run one eager step with per-GEMM-launch "
                   "event timing logged to stderr (GCCTRACE lines) and exit without a bench line")
    p.add_argument("--profile", default="", help="debug: after warm-up run one eager step under torch.profiler (CUPTI kernel "
                   "times, no replay overhead), write the per-kernel table to this file and exit without a bench line")
    p.add_argument("--clock_ms", type=int, default=200, help="nvidia-smi sampling interval during the timed region")
    p.add_argument("--watchdog_s", type=int, default=900, help="abort the whole process after this many seconds")
    p.add_argument("--no_dropout", action="store_true")
    p.add_argument("--sync_bn", action="store_true", help="N > 1: global-batch parity (synchronised BatchNorm + loss sums)")
    p.add_argument("--cpu_iters", type=int, default=0, help="timed CPU-baseline iterations at batch 1 (0 = sized to ~20 s)")
    p.add_argument("--skip_cpu_baseline", action="store_true")
    p.add_argument("--skip_e2e", action="store_true", help="profiling runs only")
    p.add_argument("--skip_roofline", action="store_true", help="profiling runs only")
    p.add_argument("--graph", type=int, default=1, help="1 (default): capture the whole iteration in a CUDA graph and replay it; 0: eager launches")
    a = p.parse_args()
    a.cfg = CONFIGS[a.config]
    if a.batch <= 0:
        a.batch = a.cfg["batch"]
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def step_gmacs(args):
    """Algorithmic GMAC per image per iteration (fwd = 1x, full bwd = 2x, dgrad-only = 1x of a net's
    forward MACs; true multiply-accumulates from gcc_b200.macs), pix2pix: 4 g_T + 4 g_S + 24 d + Gram + transform."""
    from gcc_b200 import macs, nets
    c, o = args.cfg, args.cfg["opt"]
    if c["model"] != "pix2pix":
        return None
    if o["backbone"] == "resnet":
        mk = lambda ngf, f: nets.MobileResnetGenerator(ngf=ngf, cfg=f, device="cpu")
    else:
        mk = lambda ngf, f: nets.UnetGenertor(ngf=ngf, filter_cfgs=f, channel_cfgs=c["cfgs"][1] if f is not None else None,
                                              device="cpu")
    g_t = macs.count_macs(mk(o["teacher_ngf"], None)) / 1e9
    g_s = macs.count_macs(mk(o["ngf"], c["cfgs"][0])) / 1e9
    d = macs.count_macs(nets.NLayerDiscriminator(input_nc=6, ndf=o["ndf"], device="cpu")) / 1e9
    gram = 8.0 if o["backbone"] != "resnet" else 9.4
    return {"g_T": g_t, "g_S": g_s, "d": d, "total": 4 * g_t + 4 * g_s + 24 * d + gram + 0.6}


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


def workload_name(args):
    c, o = args.cfg, args.cfg["opt"]
    if c["model"] == "pix2pix":
        stu = "%s ngf %d" % (o["backbone"], o["ngf"])
        if c["cfgs"][0] is not None:
            stu += " pruned to filter_cfgs %s" % (c["cfgs"][0],)
        return ("%s: pix2pix GCC student (%s, gated PatchGAN ndf %d) distilled online from ngf-%d teacher, 256x256, "
                "batch %d/GPU, dropout %s" % (args.config, stu, o["ndf"], o["teacher_ngf"], args.batch,
                                              "off" if args.no_dropout else "on"))
    if c["model"] == "cyclegan":
        return ("cyclegan: CycleGAN GCC student (MobileResNet ngf %d, cfgs of utils/prune_util.py:120-121, gated ndf %d) "
                "distilled online from ngf-%d teacher, 256x256, batch %d/GPU" % (o["ngf"], o["ndf"], o["teacher_ngf"], args.batch))
    if c["model"] == "srgan":
        return ("srgan: SRGAN 4x GCC student (SRResNet ngf %d, gated D ndf %d) distilled online from ngf-%d teacher, "
                "%d -> %d, truncated-VGG19 perceptual loss (random weights), batch %d/GPU" % (
                    o["ngf"], o["ndf"], o["teacher_ngf"], c["size"] // 4, c["size"], args.batch))
    return ("sagan: SAGAN GCC student (ngf %d, gated spectral-norm D ndf %d, self-attention) distilled online from ngf-%d "
            "teacher, 64x64, batch %d/GPU" % (o["ngf"], o["ndf"], o["teacher_ngf"], args.batch))


# ------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(args, iters=0, warmup=1, budget_s=20.0):
    """images/s of the CPU oracle port (oracle/*.py, pinned to the reference by tests/golden) on the host cores: a
    bounded sample = `iters` GCC iterations at batch 1 of the same networks (0: as many as fit ~budget_s), with the two
    phases (optimize_parameters / optimizer_netD_arch) timed separately."""
    import torch
    from oracle import gcc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c, o = args.cfg, args.cfg["opt"]
    model = c["model"]
    if model == "pix2pix":
        opt = O.Opt(no_dropout=args.no_dropout, direction="BtoA", **o)
        S, _ = O.build_pair(opt, c["cfgs"][0], c["cfgs"][1])
        mk = lambda t: (O.det_image(t + ".A", 1, 3, 256, 256), O.det_image(t + ".B", 1, 3, 256, 256))
    elif model == "cyclegan":
        S, _ = O.build_cycle_pair(O.CycleOpt(**o), c["cfgs"][0], c["cfgs"][1])
        mk = lambda t: (O.det_image(t + ".A", 1, 3, 256, 256), O.det_image(t + ".B", 1, 3, 256, 256))
    elif model == "srgan":
        from oracle import srgan_oracle as SR
        oo = {k: v for k, v in o.items() if k != "image_size"}
        S, _ = SR.build_sr_pair(SR.SROpt(**oo))
        hr = c["size"]
        mk = lambda t: (SR.convert_to_imagenet(O.det_image(t + ".lr", 1, 3, hr // 4, hr // 4)), O.det_image(t + ".hr", 1, 3, hr, hr))
    else:
        from oracle import sagan_oracle as SA
        S, _ = SA.build_sa_pair(SA.SAOpt(**o))
        mk = lambda t: (O.det_normal(t + ".z", (1, 128)), O.det_image(t + ".real", 1, 3, 64, 64))
    t_opt, t_arch = [], []
    it, t_start = 0, time.perf_counter()
    while True:
        a, b = mk("bench.%d" % it)
        va, vb = mk("bench.v%d" % it)
        t0 = time.perf_counter()
        S.set_input(a, b)
        S.optimize_parameters()
        t1 = time.perf_counter()
        S.set_input(va, vb)
        S.clipping_mask_alpha()
        S.optimizer_netD_arch()
        t2 = time.perf_counter()
        if it >= warmup:
            t_opt.append(t1 - t0)
            t_arch.append(t2 - t1)
        it += 1
        done = len(t_opt)
        if (iters and done >= iters) or (not iters and done >= 3 and time.perf_counter() - t_start > budget_s):
            break
    med = lambda v: sorted(v)[len(v) // 2]
    mean = (sum(t_opt) + sum(t_arch)) / len(t_opt)
    return {"value": 1.0 / mean, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d GCC iterations at batch 1 after %d warm-up (oracle port of the reference step, same nets as the "
                      "GPU arm, fp32, torch CPU %d threads), %.3f s/iter" % (len(t_opt), warmup, cores, mean),
            "s_per_iter": mean, "optimize_parameters_s_median": med(t_opt), "optimizer_netD_arch_s_median": med(t_arch)}


def cpu_c1_rate(iters=3):
    """BASELINE configs[0] / BASELINE.md section 4 (C1): pix2pix MobileResNet-9 ngf 64 + PatchGAN ndf 128, batch 1, on the
    CPU oracle port -- reported beside the headline as the reference's own CPU-runnable case."""
    a = argparse.Namespace(cfg=dict(model="pix2pix", opt=dict(ngf=64, teacher_ngf=64, ndf=128, teacher_ndf=128,
                                                              backbone="resnet"), cfgs=(None, None)), no_dropout=False)
    r = cpu_oracle_rate(a, iters=iters)
    return {k: r[k] for k in ("value", "unit", "cores", "s_per_iter", "optimize_parameters_s_median",
                              "optimizer_netD_arch_s_median")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_oracle_rate(args, iters=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    line = {"metric": "GCC train images/sec (pix2pix 256^2)" if args.cfg["model"] == "pix2pix" else
            "GCC train images/sec (%s)" % args.config, "value": base["value"], "unit": "images/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * base["s_per_iter"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "note": "CPU oracle port of the reference step, batch 1 per step"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.config == "c2":
        line["c1_resnet_ngf64_b1"] = cpu_c1_rate(iters=2)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ B200 arm
def ncu_facts(key):
    """DRAM traffic / tensor-pipe utilisation of a kernel from the committed `ncu --set full` capture, as extracted by
    scripts/ncu_extract.py into profiles/roofline_kernels.json (names the capture file it came from)."""
    path = os.path.join(ROOT, "profiles", "roofline_kernels.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get(key)
    return None


def dominant_kernel_roofline(args, pk):
    """CUDA-event timing of the dominant kernel of the workload, alone, L2 flushed between launches.
    pix2pix: PatchGAN 512->1024 k4 s1 fprop (conv_gemm_persistent_kernel<256,4>), algorithmic FLOPs =
    2 * (batch*31*31) * 1024 * (512*16), against the measured burst bf16 peak.
    CycleGAN / SRGAN / SAGAN (HBM-bound steps): the norm-apply kernel on the workload's largest normalised activation,
    algorithmic bytes = read x + write y = 4 B/element, against the measured HBM copy bandwidth."""
    import torch
    from gcc_b200 import _lib
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    model, batch = args.cfg["model"], args.batch

    def time_it(launch, reps=10):
        for _ in range(3):
            launch()
        times = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        return sum(times) / len(times)

    if model == "pix2pix":
        n, h, w, cin, cout, k = batch, 32, 32, 512, 1024, 4
        x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
        wt = (torch.randn(cout, k * k, cin, device="cuda") * 0.02).to(torch.bfloat16)
        y = torch.empty(n, 31, 31, cout, device="cuda", dtype=torch.bfloat16)
        ms = time_it(lambda: _lib.call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cin, wt.data_ptr(), cout, k * k, cin,
                                       None, y.data_ptr(), 31, 31, cout, 0, 0, k, k, 1, 1, 0, 0.0, 0, None, 0, None, 0, st))
        flops = 2.0 * n * 31 * 31 * cout * cin * k * k
        ach = flops / (ms * 1e-3) / 1e12
        facts = ncu_facts("conv_gemm_512_1024_k4s1_b%d" % n) or {}
        return {"bound": "tensor", "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_burst"], "traffic": facts.get("dram_bytes"),
                "traffic_unit": "bytes/launch (ncu dram read+write)", "ncu_tensor_pipe_active_pct": facts.get("tensor_pipe_pct"),
                "ncu_capture": facts.get("capture"),
                "kernel": "conv_gemm_persistent_kernel<256,4>: PatchGAN 512->1024 k4 s1 fprop, M=%d N=1024 K=8192" % (n * 961),
                "algorithmic_flop_per_launch": flops, "avg_launch_ms": ms,
                "peak_source": pk["source"] + ", burst (kernel timed alone)"}
    # HBM-bound workloads: instance/batch-norm apply on the largest normalised activation
    if model == "cyclegan":
        n, hw, c, per_sample, what = batch, 256 * 256, 64, 1, "InstanceNorm apply + ReLU, teacher stem 64 ch @ 256x256"
    elif model == "srgan":
        n, hw, c, per_sample, what = batch, 96 * 96, 64, 0, "BatchNorm apply, teacher SRResNet block 64 ch @ 96x96"
    else:
        n, hw, c, per_sample, what = batch, 32 * 32, 64, 0, "BatchNorm apply + ReLU, teacher l4 64 ch @ 32x32"
    x = torch.randn(n, hw, c, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x)
    sums = torch.zeros((n if per_sample else 1) * 2 * c, device="cuda")
    _lib.call("gcc_norm_stats_bf16", x.data_ptr(), n, hw, c, per_sample, sums.data_ptr(), st)
    ms = time_it(lambda: _lib.call("gcc_norm_apply_bf16", x.data_ptr(), y.data_ptr(), n, hw, c, c, per_sample, sums.data_ptr(),
                                   None, None, None, 0.5, 1e-5, None, None, 0.1, 2, 0.2, 0, None, 0, 0, 0, 0, st))
    nbytes = 4.0 * n * hw * c
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None,
            "kernel": "norm_apply_kernel: " + what, "algorithmic_bytes_per_launch": nbytes, "avg_launch_ms": ms,
            "peak_source": pk["source"] + ", copy bandwidth"}


def _watchdog(seconds):
    """A hung collective or kernel must not hold the GPU box: give up loudly after `seconds`."""
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
    os.environ.setdefault("NCCL_DEBUG", "WARN")     # never overrides the caller's setting (e.g. NCCL_DEBUG=INFO)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / logs go to stderr: stdout = one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gcc_b200 import _lib, factory
    from gcc_b200.graph import GraphedIteration
    from gcc_b200.prefetch import Prefetcher
    c = args.cfg
    name = c["model"]
    argv = []
    if args.no_dropout and name == "pix2pix":
        argv.append("--no_dropout")
    if args.sync_bn:
        argv.append("--sync_bn")
    opt = factory.make_opt(name, local, argv, batch_size=args.batch, **c["opt"])
    torch.manual_seed(1234 + rank)          # different seeds per rank: the constructors broadcast rank 0's parameters
    model, teacher = factory.build_pair(opt, c["cfgs"])

    B = args.batch
    g = torch.Generator().manual_seed(99 + rank)
    nbatch = 2
    host = [(factory.synthetic_batch(name, B, c["size"], g, pin=True), factory.synthetic_batch(name, B, c["size"], g, pin=True))
            for _ in range(nbatch)]
    devb = [tuple({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for d in pair) for pair in host]
    h2d_fp32 = sum(v.numel() * 4 for pair in host[:1] for d in pair for v in d.values() if torch.is_tensor(v))

    graphed = None

    def step(pair, read_losses):
        if graphed is not None:
            graphed.run(pair[0], pair[1])
        else:
            factory.run_iteration(model, pair[0], pair[1])
        if read_losses:
            return model.get_current_losses()  # float() of every loss: device -> host reads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, source, read_losses):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.lib().gcc_launch_count()
        torch.arange(5, device="cuda").cumsum(0)  # marker kernel: scripts/summarize_launches.py cuts the ncu list here
        torch.cuda.profiler.start()               # `ncu --profile-from-start off` captures exactly the timed region
        e0.record()
        for i in range(nsteps):
            step(source(i), read_losses)
        e1.record()
        torch.cuda.profiler.stop()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _lib.lib().gcc_launch_count() - l0

    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        for i in range(2):
            step(devb[i % nbatch], False)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(devb[0], False)
            torch.cuda.synchronize()
        rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages()), key=lambda r: -r[2])
        tot = sum(r[2] for r in rows)
        with open(args.profile, "w") as f:
            f.write("%s: one eager step, %d kernels launches, %.3f ms of device time (torch.profiler / CUPTI)\n" % (
                args.config, sum(r[1] for r in rows), tot / 1e3))
            f.write("%-90s %7s %10s %7s %9s\n" % ("kernel", "count", "ms", "share", "avg_us"))
            for k, c, t in rows:
                f.write("%-90s %7d %10.3f %6.1f%% %9.1f\n" % (k[:90], c, t / 1e3, 100.0 * t / max(tot, 1e-9), t / max(c, 1)))
        return
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
        step(devb[0], False)
        l0 = _lib.lib().gcc_launch_count()
        step(devb[1], False)
        launches_per_step = _lib.lib().gcc_launch_count() - l0  # the captured graph replays exactly these launches
        graphed = GraphedIteration(model).capture(devb[0][0], devb[0][1], warmup=1)
        if world > 1:
            # data parallel: the first replays of the graph segments and the first eager collectives between them run
            # slower than the steady state (graph upload, NCCL channel set-up; at N = 8 the e2e loop, measured later in
            # the same process, was 3-9 % faster than the device-resident loop measured right after capture): settle
            # before the W warm-up steps the contract asks for
            for i in range(10):
                step(devb[i % nbatch], False)
    for i in range(args.warmup):
        step(devb[i % nbatch], False)
    ms, launches = timed(args.steps, lambda i: devb[i % nbatch], False)
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    e2e = None
    if not args.skip_e2e:
        # end to end through the public API with HOST buffers: every step's inputs go pinned host memory -> device
        # (gcc_b200.prefetch.Prefetcher: copy stream, double-buffered device slots) inside the timed region, every
        # step's losses are read back to the host.
        def stream_of(n):
            for i in range(n):
                yield host[i % nbatch]
        pf = Prefetcher(stream_of(args.steps + 3), torch.device("cuda", local))
        for _ in range(3):       # warm the host-input path: every prefetch slot is allocated and used once
            step(next(pf), True)
        b0 = pf.h2d_bytes
        ms_e2e, _ = timed(args.steps, lambda i: next(pf), True)
        h2d = (pf.h2d_bytes - b0) / max(1, args.steps)
        pf.close()
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
               "h2d_note": "fp32 host batch from pinned memory on a copy stream, overlapped with the previous step",
               "d2h_bytes_per_step": 4 * len(model.loss_names), "ms_per_step": ms_e2e / args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    imgs = world * B * args.steps
    value = imgs / (ms * 1e-3)
    cfgd = {"workload": workload_name(args), "name": args.config, "global_batch": world * B, "parallelism": "dp%d" % world,
            "l2": "per-step working set (saved activations of all net passes at batch %d) >> 126 MB L2; two alternating "
                  "input batches" % B, "cuda_graph": bool(args.graph), "sync_bn": bool(args.sync_bn)}
    gm = step_gmacs(args)
    if gm is not None:
        step_tflops = value / world * gm["total"] * 2e9 / 1e12
        cfgd.update({"algorithmic_gmac_per_image": gm["total"], "fwd_gmac": {k: gm[k] for k in ("g_T", "g_S", "d")},
                     "gmac_formula": "4 g_T + 4 g_S + 24 d + Gram + transform (SURVEY.md 8d; true MACs, gcc_b200.macs)",
                     "step_tensor_tflops_per_gpu": step_tflops,
                     "step_frac_of_sustained_bf16_peak": step_tflops / pk["bf16_sustained"],
                     "step_frac_of_burst_bf16_peak": step_tflops / pk["bf16_burst"]})
    line = {
        "metric": "GCC train images/sec (pix2pix 256^2)" if name == "pix2pix" else "GCC train images/sec (%s)" % args.config,
        "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfgd, "clocks": clocks,
        "e2e": e2e, "gpu_launches": int(launches),
    }
    if world == 1:
        if not args.skip_roofline:
            line["roofline"] = dominant_kernel_roofline(args, pk)
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
