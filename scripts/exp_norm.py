"""HBM roofline check of the norm kernels on the largest activation (PatchGAN layer-0 output, batch 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import _lib
st = torch.cuda.current_stream().cuda_stream
N, H, W, C = 32, 128, 128, 128
x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
dy = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
y = torch.empty_like(x); dx = torch.empty_like(x)
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
sums = torch.zeros(2 * C, device="cuda"); red = torch.zeros(2 * C, device="cuda")
rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
nbytes = x.numel() * 2

def timeit(fn, label, byts, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-36s %8.1f us  %7.1f GB/s" % (label, ms * 1e3, byts / ms / 1e6), flush=True)

timeit(lambda: torch.mul(x, 2.0, out=y), "torch mul bf16 (r+w)", 2 * nbytes)
timeit(lambda: y.copy_(x), "torch copy bf16 (r+w)", 2 * nbytes)
timeit(lambda: x.float().sum(), "torch sum (read)", nbytes)
timeit(lambda: _lib.call("gcc_norm_stats_bf16", x.data_ptr(), N, H * W, C, 0, sums.data_ptr(), st), "norm_stats (read)", nbytes)
timeit(lambda: _lib.call("gcc_norm_apply_bf16", x.data_ptr(), y.data_ptr(), N, H * W, C, C, 0, sums.data_ptr(), gamma.data_ptr(),
                         beta.data_ptr(), None, 0.5, 1e-5, rm.data_ptr(), rv.data_ptr(), 0.1, 1, 0.2, 0, None, 0, 0, 0, 0, st),
       "norm_apply bn+lrelu (r+w)", 2 * nbytes)
timeit(lambda: _lib.call("gcc_norm_bwd_bf16", x.data_ptr(), N, H * W, C, C, 0, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                         None, 0.5, 1e-5, 1, 0.2, 0, dy.data_ptr(), C, 0, None, 0, 0, 0, red.data_ptr(), dx.data_ptr(), None, None,
                         None, 0, 0, None, st), "norm_bwd reduce+apply (2r+2r+w)", 5 * nbytes)
timeit(lambda: _lib.call("gcc_act_fwd_bf16", x.data_ptr(), y.data_ptr(), x.numel(), 1, 0.2, st), "act_fwd lrelu (r+w)", 2 * nbytes)
