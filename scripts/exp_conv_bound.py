"""Single-launch timings of the persistent conv kernel on the layers that were furthest from their bound: full, without
the epilogue stores (debug bit 0; garbage results), and the tail-wave split on / off (bit 11) on the layers it targets.
(The revision that produced profiles/r02_conv_pipeline_bound.txt also had switches for the A boxes / B boxes / MMAs
inside the producer and MMA loops; they were removed with the instruction clean-up of those loops.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gcc_b200 import _lib, ops

L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(label, n, h, w, cin, cout, k, s, p, tr=0, flags=0, stats=False, reps=8):
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, k * k, cin, device="cuda") * 0.02).to(torch.bfloat16)
    if not tr:
        oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    else:
        oh, ow = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
    cop = (cout + 15) // 16 * 16
    y = torch.empty(n, oh, ow, cop, device="cuda", dtype=torch.bfloat16)
    ws = ops._tail_ws(x.device)
    sums = torch.zeros(2 * cop, device="cuda") if stats else None
    L.gcc_debug_set_flags(flags)

    def go():
        _lib.call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cin, wt.data_ptr(), cout, k * k, cin, None, y.data_ptr(),
                  oh, ow, cop, 0, tr, k, k, s, p, 0, 0.0, 0, ws.data_ptr(), ws.numel(),
                  None if sums is None else sums.data_ptr(), cop, st)
    for _ in range(3):
        go()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in evs:
        flush.zero_()          # L2 flush between launches
        e0.record()
        go()
        e1.record()
    torch.cuda.synchronize()
    L.gcc_debug_set_flags(0)
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    ms = ts[len(ts) // 2]
    taps = k * k // (4 if (tr and s == 2) else 1)
    fl = 2.0 * n * oh * ow * cout * cin * taps
    print("%-34s %-22s %8.1f us  %7.0f TFLOP/s" % (label, VARIANTS.get(flags, str(flags)), ms * 1e3, fl / ms / 1e9), flush=True)


VARIANTS = {0: "full", 1: "no stores", 256: "no A boxes", 512: "no B boxes", 768: "no TMA loads", 1024: "no MMA",
            1025: "no MMA, no stores", 769: "no loads, no stores", 1792: "no loads, no MMA", 2048: "tail split off"}
B = 32
LAYERS = [
    ("D dgrad 256->128 k4s2 (BN128)", (B, 64, 64, 256, 128, 4, 2, 1, 1)),
    ("U-Net convT 256->64 k4s2 (BN64)", (B, 64, 64, 256, 64, 4, 2, 1, 1)),
    ("D fprop 512->1024 k4s1 (BN256)", (B, 32, 32, 512, 1024, 4, 1, 1, 0)),
]
for label, shp in LAYERS:
    for fl in (0, 1):
        run(label, *shp, flags=fl)
print()
for label, shp, stats in (("D dgrad 1024->512 k4s1 (512 tiles)", (B, 31, 31, 1024, 512, 4, 1, 1, 1), False),
                          ("D fprop 256->512 k4s2 (512 tiles)", (B, 64, 64, 256, 512, 4, 2, 1, 0), True),
                          ("G fprop 128->512 k4s2 (512 tiles)", (B, 64, 64, 128, 512, 4, 2, 1, 0), True)):
    for fl in (0, 2048):
        run(label, *shp, flags=fl, stats=stats)
