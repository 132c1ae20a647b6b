"""Timing experiments on the short-K 1x1 GEMM shape (col-path layer 0): where does the time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream

def run(n, h, w, cin, cout, k=1, s=1, p=0, bias=False, bn=0, flags=0, reps=10, label=""):
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, k * k, cin, device="cuda") * 0.02).to(torch.bfloat16)
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    cop = (cout + 7) // 8 * 8
    y = torch.empty(n, oh, ow, cop, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(cout, device="cuda") if bias else None
    L.gcc_debug_force_block_n(bn)
    L.gcc_debug_set_flags(flags)
    def go():
        _lib.call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cin, wt.data_ptr(), cout, k * k, cin, _lib.ptr(b),
                  y.data_ptr(), oh, ow, cop, 0, 0, k, k, s, p, 0, 0.0, 0, None, 0, None, 0, st)
    for _ in range(3): go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = x.numel() * 2 + y.numel() * 2
    fl = 2.0 * n * oh * ow * cout * cin * k * k
    print("%-44s %8.1f us  %7.1f GB/s  %7.1f TFLOP/s" % (label, ms * 1e3, byts / ms / 1e6, fl / ms / 1e9), flush=True)
    L.gcc_debug_force_block_n(0); L.gcc_debug_set_flags(0)


def run_wgrad(n, h, w, cin, cout, k, s, p, label, reps=10):
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, oh, ow, cout, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(cout, k * k, cin, device="cuda")
    def go():
        _lib.call("gcc_wgrad_gemm_bf16", dy.data_ptr(), n, oh, ow, cout, x.data_ptr(), h, w, cin, dw.data_ptr(), cout,
                  cin, k, k, s, p, 0, 1, 1.0, st)
    for _ in range(3): go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * oh * ow * cout * cin * k * k
    print("%-44s %8.1f us  %7.1f TFLOP/s" % (label, ms * 1e3, fl / ms / 1e9), flush=True)

B = 32
for fl, nm in ((16, "MT=1 (128-row tiles)"), (0, "MT=2 (256-row tiles)")):
    L.gcc_debug_set_flags(fl)
    run_wgrad(B, 32, 32, 512, 1024, 4, 1, 1, "wgrad D L3 512->1024 k4s1 " + nm)
    run_wgrad(B, 64, 64, 256, 512, 4, 2, 1, "wgrad D L2 256->512 k4s2 " + nm)
    run_wgrad(B, 128, 128, 128, 256, 4, 2, 1, "wgrad D L1 128->256 k4s2 " + nm)
    run_wgrad(B, 16, 16, 512, 512, 4, 2, 1, "wgrad U-Net 512->512 k4s2 16->8 " + nm)
    L.gcc_debug_set_flags(0)
