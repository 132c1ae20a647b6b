"""norm kernels (forward apply / backward reduce / backward apply separately, L2 flushed before every launch) on the
PatchGAN activations at batch 32.  GCC_B200_NORM_REDUCE_FLAT=4|8 selects the un-pipelined reduce variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import _lib
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, H, W, C) in ((32, 128, 128, 128), (32, 64, 64, 256), (32, 32, 32, 512), (32, 31, 31, 1024)):
    x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
    dx = torch.empty_like(x)
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    sums = torch.zeros(2 * C, device="cuda"); red = torch.zeros(2 * C, device="cuda")
    _lib.call("gcc_norm_stats_bf16", x.data_ptr(), N, H * W, C, 0, sums.data_ptr(), st)
    nbytes = x.numel() * 2
    y = torch.empty_like(x)
    rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
    for phase, label, mult in ((0, "fwd apply (1 read + 1 write)", 2), (1, "reduce (2 reads)", 2), (2, "apply (2 reads + 1 write)", 3)):
        def fn():
            if phase == 0:
                _lib.call("gcc_norm_apply_bf16", x.data_ptr(), y.data_ptr(), N, H * W, C, C, 0, sums.data_ptr(), gamma.data_ptr(),
                          beta.data_ptr(), None, 0.5, 1e-5, rm.data_ptr(), rv.data_ptr(), 0.1, 1, 0.2, 0, None, 0, 0, 0, 0, st)
                return
            _lib.call("gcc_norm_bwd_bf16", x.data_ptr(), N, H * W, C, C, 0, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                      None, 0.5, 1e-5, 1, 0.2, 0, dy.data_ptr(), C, 0, None, 0, 0, 0, red.data_ptr(), dx.data_ptr(), None, None,
                      None, 0, phase, None, st)
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print("reduce_flat %s  [%d,%d,%d,%d]  %-26s %7.1f us  %6.0f GB/s" % (os.environ.get("GCC_B200_NORM_REDUCE_FLAT", "0"), N, H, W, C, label,
                                                                      ms * 1e3, mult * nbytes / ms / 1e6), flush=True)
