"""Launch a few representative conv_gemm / wgrad shapes a handful of times (for `ncu --set full -k regex:...`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gcc_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
st = torch.cuda.current_stream().cuda_stream


def conv(n, h, w, cin, cout, k, s, p, tr=0):
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, k * k, cin, device="cuda") * 0.02).to(torch.bfloat16)
    if not tr:
        oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    else:
        oh, ow = (h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k
    cop = (cout + 7) // 8 * 8
    y = torch.empty(n, oh, ow, cop, device="cuda", dtype=torch.bfloat16)
    for _ in range(reps):
        _lib.call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cin, wt.data_ptr(), cout, k * k, cin, None, y.data_ptr(),
                  oh, ow, cop, 0, tr, k, k, s, p, 0, 0.0, 0, None, 0, None, 0, st)
    torch.cuda.synchronize()


def wgrad(n, h, w, cin, cout, k, s, p):
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, oh, ow, cout, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(cout, k * k, cin, device="cuda")
    for _ in range(reps):
        _lib.call("gcc_wgrad_gemm_bf16", dy.data_ptr(), n, oh, ow, cout, x.data_ptr(), h, w, cin, dw.data_ptr(), cout,
                  cin, k, k, s, p, 0, 1, 1.0, st)
    torch.cuda.synchronize()


conv(B, 32, 32, 512, 1024, 4, 1, 1)      # PatchGAN 512->1024 k4 s1 fprop (dominant)
conv(B, 128, 128, 128, 128, 1, 1, 0)     # short-K 1x1 (col-path layer-0 GEMM shape): M=524288 N=128 K=128
conv(B, 128, 128, 128, 256, 4, 2, 1)     # PatchGAN 128->256 k4 s2 fprop
conv(B, 1, 1, 512, 512, 4, 2, 1, 1)      # U-Net innermost ConvT (tiny M, K=512 per tap)
wgrad(B, 32, 32, 512, 1024, 4, 1, 1)     # dominant wgrad
print("done")
