"""GPU probe for the tcgen05 conv / wgrad kernels: compares against the CUDA-core direct kernels and
torch fp32 on identical bf16-rounded inputs and prints diagnostics.  Run on the B200 box:
    python scripts/probe_conv.py [case-filter]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from gcc_b200 import _lib

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _check  # noqa: E402

dev = "cuda"


def rp8(c):
    return (c + 7) // 8 * 8


def pack_act(x_nchw, cp=None):
    """NCHW fp32 -> NHWC bf16 with channels padded to a multiple of 8."""
    n, c, h, w = x_nchw.shape
    cp = cp or rp8(c)
    out = torch.zeros(n, h, w, cp, dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def pack_w(w_rtc, cp=None):
    """[R, T, C] fp32 -> bf16 with C padded."""
    r, t, c = w_rtc.shape
    cp = cp or rp8(c)
    out = torch.zeros(r, t, cp, dtype=torch.bfloat16, device=w_rtc.device)
    out[..., :c] = w_rtc.to(torch.bfloat16)
    return out


def run_conv(fn, x, w, bias, OH, OW, R, transposed, KH, KW, stride, pad, act=0, slope=0.2, cy=None, coff=0):
    N, H, W, Cx = x.shape
    Rr, T, Cw = w.shape
    cy = cy or rp8(R)
    y = torch.full((N, OH, OW, cy), 7.0, dtype=torch.bfloat16, device=dev)
    _check.call(fn, x.data_ptr(), N, H, W, Cx, w.data_ptr(), R, T, Cw, _lib.ptr(bias), y.data_ptr(), OH, OW, cy, coff,
              transposed, KH, KW, stride, pad, act, slope, 0, None, 0, None, 0, _lib.current_stream())
    torch.cuda.synchronize()
    return y


def conv_case(name, N, H, W, Cin, Cout, k, stride, pad, transposed=0, outpad=0, bias=True, act=0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    if not transposed:
        wt = (torch.randn(Cout, Cin, k, k, generator=g) * 0.1).to(dev)
        OH = (H + 2 * pad - k) // stride + 1
        OW = (W + 2 * pad - k) // stride + 1
        w_rtc = wt.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin)
    else:
        wt = (torch.randn(Cin, Cout, k, k, generator=g) * 0.1).to(dev)
        OH = (H - 1) * stride - 2 * pad + k + outpad
        OW = (W - 1) * stride - 2 * pad + k + outpad
        w_rtc = wt.permute(1, 2, 3, 0).reshape(Cout, k * k, Cin)
    b = (torch.randn(Cout, generator=g)).to(dev) if bias else None
    xb = pack_act(x)
    wb = pack_w(w_rtc)
    xr = xb[..., :Cin].permute(0, 3, 1, 2).float()
    wr = wt.to(torch.bfloat16).float()
    if not transposed:
        ref = F.conv2d(xr, wr, b, stride=stride, padding=pad)
    else:
        ref = F.conv_transpose2d(xr, wr, b, stride=stride, padding=pad, output_padding=outpad)
    if act == 1:
        ref = F.leaky_relu(ref, 0.2)
    elif act == 2:
        ref = torch.tanh(ref)
    ref = ref.permute(0, 2, 3, 1)
    t0 = time.time()
    y_d = run_conv("gcc_conv_direct_bf16", xb, wb, b, OH, OW, Cout, transposed, k, k, stride, pad, act)
    y_t = run_conv("gcc_conv_gemm_bf16", xb, wb, b, OH, OW, Cout, transposed, k, k, stride, pad, act)
    yd = y_d[..., :Cout].float()
    yt = y_t[..., :Cout].float()
    scale = ref.abs().max().item() + 1e-6
    e_d = (yd - ref).abs().max().item() / scale
    e_t = (yt - ref).abs().max().item() / scale
    padok = bool((y_t[..., Cout:].float() == 0).all().item())
    ok = e_t < 2e-2 and padok
    print("%-34s direct_err %.2e  tcgen05_err %.2e  pad_zero %s  %s" % (name, e_d, e_t, padok, "OK" if ok else "FAIL"),
          flush=True)
    if not ok:
        bad = ((yt - ref).abs() / scale > 2e-2)
        print("   bad fraction %.4f; per-col-block(8) bad:" % bad.float().mean().item(),
              [round(v, 2) for v in bad.float().mean(dim=(0, 1, 2)).reshape(-1, min(8, Cout)).mean(1).tolist()][:32])
        rows = bad.reshape(-1, Cout).float().mean(1)
        print("   first 64 pixel rows bad:", [int(v > 0) for v in rows[:64].tolist()])
        print("   sample ref ", ref.reshape(-1, Cout)[0, :8].tolist())
        print("   sample got ", yt.reshape(-1, Cout)[0, :8].tolist())
    return ok


def wgrad_case(name, N, H, W, Cin, Cout, k, stride, pad, batched=0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    OH = (H + 2 * pad - k) // stride + 1
    OW = (W + 2 * pad - k) // stride + 1
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    dy = torch.randn(N, Cout, OH, OW, generator=g).to(dev)
    xb, dyb = pack_act(x), pack_act(dy)
    xr = xb[..., :Cin].permute(0, 3, 1, 2).float()
    dyr = dyb[..., :Cout].permute(0, 3, 1, 2).float()
    if not batched:
        wz = torch.zeros(Cout, Cin, k, k, device=dev, requires_grad=True)
        F.conv2d(xr, wz, None, stride=stride, padding=pad).backward(dyr)
        ref = wz.grad.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin)
        shape = (Cout, k * k, Cin)
    else:
        fm = dyr.reshape(N, Cout, OH * OW)
        fq = xr.reshape(N, Cin, H * W)
        ref = torch.bmm(fm, fq.transpose(1, 2)).reshape(N, Cout, 1, Cin)
        shape = (N, Cout, 1, Cin)
    outs = {}
    for fn in ("gcc_wgrad_direct_bf16", "gcc_wgrad_gemm_bf16"):
        dw = torch.full(shape, 3.0, dtype=torch.float32, device=dev)
        _check.call(fn, dyb.data_ptr(), N, OH, OW, dyb.shape[-1], xb.data_ptr(), H, W, xb.shape[-1], dw.data_ptr(),
                  Cout, Cin, k, k, stride, pad, batched, 0, 1.0, _lib.current_stream())
        torch.cuda.synchronize()
        outs[fn] = dw
    scale = ref.abs().max().item() + 1e-6
    e_d = (outs["gcc_wgrad_direct_bf16"] - ref).abs().max().item() / scale
    e_t = (outs["gcc_wgrad_gemm_bf16"] - ref).abs().max().item() / scale
    ok = e_t < 1e-3
    print("%-34s direct_err %.2e  tcgen05_err %.2e  %s" % (name, e_d, e_t, "OK" if ok else "FAIL"), flush=True)
    if not ok:
        got = outs["gcc_wgrad_gemm_bf16"].reshape(-1, Cin)
        rf = ref.reshape(-1, Cin)
        bad = ((got - rf).abs() / scale > 1e-3)
        print("   bad fraction %.4f" % bad.float().mean().item())
        print("   sample ref ", rf[0, :8].tolist())
        print("   sample got ", got[0, :8].tolist())
        print("   ratio got/ref", (got[0, :8] / rf[0, :8]).tolist())
    return ok


def decode_case():
    """One 128x64x64 tile: X = one-hot rows, W encodes (r, c) -> shows which K index each lane really reads."""
    N, H, W, C, R = 1, 8, 16, 64, 64
    x = torch.zeros(N, H, W, C, dtype=torch.bfloat16, device=dev)
    pix = torch.arange(H * W, device=dev)
    x.view(-1, C)[pix, pix % C] = 1.0
    w = torch.zeros(R, 1, C, dtype=torch.bfloat16, device=dev)
    rr = torch.arange(R, device=dev).view(R, 1)
    cc = torch.arange(C, device=dev).view(1, C)
    w[:, 0, :] = (rr * 64 + cc).to(torch.bfloat16)  # exact in bf16 up to 256; fine for a coarse decode
    y = run_conv("gcc_conv_gemm_bf16", x, w, None, H, W, R, 0, 1, 1, 1, 0)
    yv = y.view(-1, R).float()
    exp = (torch.arange(R, device=dev).view(1, R) * 64 + (pix % C).view(-1, 1)).to(torch.bfloat16).float()
    ok = bool((yv == exp).all().item())
    print("decode_case exact:", ok, flush=True)
    if not ok:
        for r in (0, 1, 2, 7, 8, 9, 17, 33, 64, 65, 127):
            print("   row %3d got %s   exp %s" % (r, yv[r, :6].tolist(), exp[r, :6].tolist()))
    return ok


def main():
    filt = sys.argv[1] if len(sys.argv) > 1 else ""
    _lib.call("gcc_check_device")
    results = []

    def want(n):
        return filt in n

    if want("decode"):
        results.append(("decode", decode_case()))
    convs = [
        # name, N, H, W, Cin, Cout, k, stride, pad, transposed, outpad
        ("gemm1x1_c64_r64", 2, 16, 16, 64, 64, 1, 1, 0, 0, 0),
        ("gemm1x1_c128_r128", 2, 16, 16, 128, 128, 1, 1, 0, 0, 0),
        ("gemm1x1_c256_r256", 2, 16, 16, 256, 256, 1, 1, 0, 0, 0),
        ("gemm1x1_c40_r24", 2, 16, 16, 40, 24, 1, 1, 0, 0, 0),
        ("conv3x3_s1", 2, 16, 16, 64, 64, 3, 1, 1, 0, 0),
        ("conv4x4_s1_31", 2, 32, 32, 64, 128, 4, 1, 1, 0, 0),
        ("conv4x4_s1_out1", 2, 31, 31, 128, 1, 4, 1, 1, 0, 0),
        ("conv4x4_s2", 2, 32, 32, 64, 128, 4, 2, 1, 0, 0),
        ("conv4x4_s2_cin6", 2, 64, 64, 6, 128, 4, 2, 1, 0, 0),
        ("conv4x4_s2_small", 4, 2, 2, 64, 64, 4, 2, 1, 0, 0),
        ("conv3x3_s2", 2, 32, 32, 64, 128, 3, 2, 1, 0, 0),
        ("conv7x7_s1", 1, 38, 38, 8, 64, 7, 1, 0, 0, 0),
        ("conv_odd_c37_r65", 2, 16, 16, 37, 65, 4, 2, 1, 0, 0),
        ("convT4x4_s2", 2, 16, 16, 64, 128, 4, 2, 1, 1, 0),
        ("convT4x4_s2_1x1", 4, 1, 1, 64, 64, 4, 2, 1, 1, 0),
        ("convT3x3_s2_op1", 2, 16, 16, 64, 32, 3, 2, 1, 1, 1),
        ("convT4x4_s1", 2, 31, 31, 64, 64, 4, 1, 1, 1, 0),
        ("conv4x4_s1_big", 2, 32, 32, 512, 1024, 4, 1, 1, 0, 0),
    ]
    for c in convs:
        if want(c[0]):
            name, N, H, W, Cin, Cout, k, s, p, tr, op = c
            try:
                results.append((name, conv_case(name, N, H, W, Cin, Cout, k, s, p, tr, op)))
            except Exception as e:  # noqa
                print(name, "EXC", e, flush=True)
                results.append((name, False))
    if want("act"):
        results.append(("act_lrelu", conv_case("act_lrelu", 2, 16, 16, 64, 64, 3, 1, 1, act=1)))
        results.append(("act_tanh", conv_case("act_tanh", 2, 16, 16, 64, 64, 3, 1, 1, act=2)))
    wg = [
        ("wgrad1x1_c64_r64", 2, 16, 16, 64, 64, 1, 1, 0, 0),
        ("wgrad1x1_c128_r128", 2, 16, 16, 128, 128, 1, 1, 0, 0),
        ("wgrad1x1_c256_r256", 2, 16, 16, 256, 256, 1, 1, 0, 0),
        ("wgrad3x3_s1", 2, 16, 16, 64, 64, 3, 1, 1, 0),
        ("wgrad4x4_s2", 2, 32, 32, 64, 128, 4, 2, 1, 0),
        ("wgrad4x4_s1_31", 2, 32, 32, 64, 128, 4, 1, 1, 0),
        ("wgrad_odd_c37_r65", 2, 16, 16, 37, 65, 4, 2, 1, 0),
        ("wgrad_small_2x2", 8, 4, 4, 64, 64, 4, 2, 1, 0),
        ("wgrad4x4_s1_out1", 2, 31, 31, 128, 1, 4, 1, 1, 0),
        ("gram_c128", 3, 16, 16, 128, 128, 1, 1, 0, 1),
        ("gram_c40_31", 2, 31, 31, 40, 40, 1, 1, 0, 1),
        ("wgrad_big", 4, 32, 32, 256, 256, 4, 1, 1, 0),
    ]
    for c in wg:
        if want(c[0]):
            name, N, H, W, Cin, Cout, k, s, p, bt = c
            try:
                results.append((name, wgrad_case(name, N, H, W, Cin, Cout, k, s, p, bt)))
            except Exception as e:  # noqa
                print(name, "EXC", e, flush=True)
                results.append((name, False))
    nfail = sum(1 for _, ok in results if not ok)
    print("SUMMARY: %d/%d ok" % (len(results) - nfail, len(results)))
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
