"""Image-mode (8-channel input, K = taps x channels) conv / wgrad against the CUDA-core direct kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import _lib
import _check
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
def conv(fn, x, w, b, R, flags):
    n, h, wd, _ = x.shape
    rp = (R + 7) // 8 * 8
    y = torch.zeros(n, h // 2, wd // 2, rp, device="cuda", dtype=torch.bfloat16)
    L.gcc_debug_set_flags(flags)
    _check.call(fn, x.data_ptr(), n, h, wd, 8, w.data_ptr(), R, 16, 8, b.data_ptr(), y.data_ptr(), h // 2, wd // 2, rp, 0, 0,
              4, 4, 2, 1, 1, 0.2, 0, None, 0, None, 0, st)
    torch.cuda.synchronize(); L.gcc_debug_set_flags(0)
    return y.float()
def wgrad(fn, dy, x, R, flags):
    n, h, wd, _ = x.shape
    dw = torch.zeros(R, 16, 8, device="cuda")
    L.gcc_debug_set_flags(flags)
    _check.call(fn, dy.data_ptr(), n, h // 2, wd // 2, dy.shape[3], x.data_ptr(), h, wd, 8, dw.data_ptr(), R, 8, 4, 4, 2, 1,
              0, 0, 1.0, st)
    torch.cuda.synchronize(); L.gcc_debug_set_flags(0)
    return dw
for (n, h, R) in ((2, 32, 64), (3, 64, 128), (2, 16, 40), (4, 128, 256)):
    x = torch.randn(n, h, h, 8, device="cuda").to(torch.bfloat16); x[..., 6:] = 0
    w = (torch.randn(R, 16, 8, device="cuda") * 0.1).to(torch.bfloat16); w[..., 6:] = 0
    b = torch.randn(R, device="cuda")
    ref = conv("gcc_conv_direct_bf16", x, w, b, R, 0)
    for fl in (0, 128):  # 128 = generic path (image mode off)
        try:
            y = conv("gcc_conv_gemm_bf16", x, w, b, R, fl)
            print("conv n%d h%d R%d flags %d: max err %.3e (ref max %.2f)" % (n, h, R, fl, (y - ref).abs().max().item(), ref.abs().max().item()), flush=True)
        except Exception as e:
            print("conv flags", fl, "error", e)
    dy = torch.randn(n, h // 2, h // 2, (R + 7) // 8 * 8, device="cuda").to(torch.bfloat16)
    refw = wgrad("gcc_wgrad_direct_bf16", dy, x, R, 0)
    for fl in (0, 128):  # 128 = generic path (image mode off)
        try:
            dw = wgrad("gcc_wgrad_gemm_bf16", dy, x, R, fl)
            print("wgrad n%d h%d R%d flags %d: max err %.3e (ref max %.2f)" % (n, h, R, fl, (dw - refw).abs().max().item(), refw.abs().max().item()), flush=True)
        except Exception as e:
            print("wgrad flags", fl, "error", e)
