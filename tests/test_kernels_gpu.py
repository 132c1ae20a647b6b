"""Per-kernel parity on the B200: every C-ABI kernel against a plain torch fp32 computation of the same
op on the same bf16-rounded inputs.  Metric: relative L2 per tensor (|a - b|_2 / |b|_2); tolerance 6e-3 for bf16
outputs (rounding alone is 2^-9 / sqrt(3) = 1.1e-3), 2e-3 for fp32 outputs.  The torch reference runs in true fp32
(TF32 off for cuDNN and matmul)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

BF16_TOL = 6e-3


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gcc_b200 import _lib, arena, nets, ops
    _lib.call("gcc_check_device")

    class NS:
        pass

    ns = NS()
    ns.lib, ns.arena, ns.nets, ns.ops = _lib, arena, nets, ops
    return ns


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def nhwc(x, G):
    return G.ops.to_nhwc(x)


def nchw(x, c, G):
    return G.ops.to_nchw(x, c)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def bf(x):
    return x.to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------ conv
CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, transposed, outpad
    (2, 16, 16, 64, 64, 1, 1, 0, 0, 0),
    (2, 16, 16, 40, 24, 1, 1, 0, 0, 0),
    (2, 32, 32, 64, 128, 4, 1, 1, 0, 0),
    (2, 31, 31, 128, 1, 4, 1, 1, 0, 0),
    (2, 64, 64, 6, 128, 4, 2, 1, 0, 0),
    (4, 2, 2, 64, 64, 4, 2, 1, 0, 0),
    (2, 32, 32, 64, 128, 3, 2, 1, 0, 0),
    (1, 38, 38, 3, 64, 7, 1, 0, 0, 0),
    (2, 16, 16, 37, 65, 4, 2, 1, 0, 0),
    (2, 16, 16, 64, 128, 4, 2, 1, 1, 0),
    (4, 1, 1, 64, 64, 4, 2, 1, 1, 0),
    (2, 16, 16, 64, 32, 3, 2, 1, 1, 1),
    (2, 8, 8, 77, 45, 4, 2, 1, 1, 0),
    (1, 32, 32, 512, 1024, 4, 1, 1, 0, 0),
    (2, 32, 32, 3, 64, 4, 2, 1, 0, 0),    # col path (im2col): first conv of G
    (3, 16, 16, 6, 128, 4, 2, 1, 0, 0),   # col path: first conv of D
    (2, 16, 16, 128, 3, 4, 2, 1, 1, 0),   # col path (col2im): last ConvTranspose of the U-Net
    (2, 8, 8, 45, 3, 4, 2, 1, 1, 0),
    (3, 31, 31, 256, 1, 4, 1, 1, 0, 0),   # head path: 1x1 GEMM + fold
    (2, 15, 17, 72, 3, 4, 1, 1, 0, 0),
    (2, 30, 26, 3, 32, 7, 1, 0, 0, 0),    # stem path (row-window GEMM): k7 on a pre-padded image
    (2, 20, 24, 3, 64, 9, 1, 4, 0, 0),    # stem path: k9 p4 (SRResNet), two 8-tap blocks per kernel row
    (2, 16, 16, 3, 64, 3, 1, 1, 0, 0),    # stem path: k3 p1 (VGG19 / SRGAN discriminator first conv)
    (3, 22, 22, 24, 3, 7, 1, 0, 0, 0),    # fold path: k7 head, 3 output channels
    (2, 16, 20, 64, 3, 9, 1, 4, 0, 0),    # fold path: k9 p4 head (SRResNet)
    (2, 12, 12, 40, 6, 3, 1, 1, 0, 0),    # fold path with an 8-channel group
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_layer_fwd_bwd(G, case):
    """ConvFn (tcgen05 fprop + dgrad + wgrad + bias grad) vs torch autograd."""
    N, H, W, Cin, Cout, k, s, p, tr, op = case
    A = G.arena.ParamArena("cuda")
    layer = G.nets.ConvLayer(A, "c", "convT" if tr else "conv", Cin, Cout, k, s, p, outpad=op, bias=True)
    A.finalize()
    layer.bind()
    wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
    w0 = rnd(wshape, 1, 0.1)
    b0 = rnd((Cout,), 2)
    with torch.no_grad():
        layer.weight.copy_(w0)
        layer.bias.copy_(b0)
    A.mark_dirty()
    x = rnd((N, Cin, H, W), 3)
    xh = nhwc(x, G).requires_grad_(True)
    y = layer(xh)
    OH, OW = y.shape[1], y.shape[2]
    gy = rnd((N, Cout, OH, OW), 4)
    A.zero_grad()
    y.backward(nhwc(gy, G))
    # torch reference on bf16-rounded operands
    xr = bf(x).requires_grad_(True)
    wr = bf(w0).requires_grad_(True)
    br = b0.clone().requires_grad_(True)
    if tr:
        yr = F.conv_transpose2d(xr, wr, br, stride=s, padding=p, output_padding=op)
    else:
        yr = F.conv2d(xr, wr, br, stride=s, padding=p)
    yr.backward(bf(gy))
    assert rel_err(nchw(y.detach(), Cout, G), yr.detach()) < BF16_TOL
    assert rel_err(nchw(xh.grad, Cin, G), xr.grad) < BF16_TOL
    assert rel_err(layer.weight.grad, wr.grad) < 2e-3
    assert rel_err(layer.bias.grad, br.grad) < 2e-3
    if Cout % 8:
        assert bool((y.detach()[..., Cout:] == 0).all())


# 512 tiles on the 148-SM persistent grid (3.46 waves): the last 68 tiles run as 2 x 68 half-K work items + the finalize
# kernel (conv_gemm.cu, "tail-wave split").  fprop k4 s1; the data gradient of a k4 s1 conv; a stride-2 transposed conv
# (four sub-pixel classes).
TAIL_CASES = [
    (32, 32, 32, 128, 512, 4, 1, 1, 0),
    (32, 31, 31, 512, 128, 4, 1, 1, 0),
    (32, 16, 16, 512, 512, 4, 2, 1, 1),
]


@pytest.mark.parametrize("case", TAIL_CASES)
def test_conv_tail_wave_split(G, case):
    """The tail-wave split against torch, against the un-split kernel (debug bit 11 switches it off) and with the fused
    BatchNorm statistics."""
    N, H, W, Cin, Cout, k, s, p, tr = case
    A = G.arena.ParamArena("cuda")
    layer = G.nets.ConvLayer(A, "c", "convT" if tr else "conv", Cin, Cout, k, s, p, bias=True)
    A.finalize()
    layer.bind()
    wshape = (Cin, Cout, k, k) if tr else (Cout, Cin, k, k)
    w0, b0 = rnd(wshape, 1, 0.05), rnd((Cout,), 2)
    with torch.no_grad():
        layer.weight.copy_(w0)
        layer.bias.copy_(b0)
    A.mark_dirty()
    A.ensure_packed()
    x = rnd((N, Cin, H, W), 3)
    res = {}
    for flags in (0, 2048):
        G.lib.lib().gcc_debug_set_flags(flags)
        G.lib.lib().gcc_debug_set_tail_min_kb(32)   # (by default only tiles of >= 128 k-blocks are split)
        try:
            l0 = G.lib.lib().gcc_launch_count()
            xh = nhwc(x, G).requires_grad_(True)
            y, sums = layer.with_stats(xh)
            OH, OW = y.shape[1], y.shape[2]
            gy = rnd((N, Cout, OH, OW), 4)
            A.zero_grad()
            y.backward(nhwc(gy, G))
            torch.cuda.synchronize()
            res[flags] = (y.detach().clone(), xh.grad.clone(), sums.clone(), G.lib.lib().gcc_launch_count() - l0)
        finally:
            G.lib.lib().gcc_debug_set_flags(0)
            G.lib.lib().gcc_debug_set_tail_min_kb(0)
    (y1, dx1, st1, n1), (y0, dx0, st0, n0) = res[0], res[2048]
    assert n1 == n0 + 1, "exactly one of the convs of this case takes the tail-wave split (one extra finalize launch)"
    xr = bf(x).requires_grad_(True)
    wr, br = bf(w0), b0
    yr = F.conv_transpose2d(xr, wr, br, stride=s, padding=p) if tr else F.conv2d(xr, wr, br, stride=s, padding=p)
    yr.backward(bf(gy))
    for y_, dx_, st_ in ((y1, dx1, st1), (y0, dx0, st0)):
        assert rel_err(nchw(y_, Cout, G), yr.detach()) < BF16_TOL
        assert rel_err(nchw(dx_, Cin, G), xr.grad) < BF16_TOL
        yf = nchw(y_, Cout, G).double()
        cp = st_.numel() // 2
        assert rel_err(st_[:Cout], yf.sum((0, 2, 3))) < 1e-4
        assert rel_err(st_[cp:cp + Cout], (yf * yf).sum((0, 2, 3))) < 1e-4
    # split and un-split results differ only by the fp32 summation order of the two K halves
    assert rel_err(y1.float(), y0.float()) < 1e-3 and rel_err(dx1.float(), dx0.float()) < 1e-3


# ------------------------------------------------------------------------------------------ norm
@pytest.mark.parametrize("mode,gated,act,dual,C", [("bn", False, 1, False, 64), ("bn", True, 1, False, 40),
                                                   ("bn", False, 1, True, 24), ("in", False, 2, False, 32),
                                                   ("id", True, 1, False, 16), ("id", False, 1, True, 8),
                                                   ("bn", False, 2, False, 130)])
def test_norm_block_fwd_bwd(G, mode, gated, act, dual, C):
    N, H, W = 3, 9, 7
    A, GA = G.arena.ParamArena("cuda"), G.arena.ParamArena("cuda")
    layer = G.nets.NormLayer(A, "n", C, mode, "cuda", GA if gated else None, "g" if gated else None, 0.5)
    A.finalize()
    GA.finalize()
    layer.bind()
    gam = rnd((C,), 1, 0.3) + 1.0
    bet = rnd((C,), 2, 0.5)
    alpha = torch.tensor(([0.7, 0.5, 0.2, 1.0] * C)[:C], device="cuda")
    with torch.no_grad():
        if mode == "bn":
            layer.gamma.copy_(gam)
            layer.beta.copy_(bet)
        if gated:
            layer.alpha.copy_(alpha)
    x = rnd((N, C, H, W), 3, 2.0) + 0.5
    xh = nhwc(x, G).requires_grad_(True)
    out = layer(xh, act, 2 if dual else None)
    y, y2 = (out if dual else (out, None))
    gy, gy2 = rnd((N, C, H, W), 4), rnd((N, C, H, W), 5)
    A.zero_grad()
    GA.zero_grad()
    loss_parts = [(y, gy)] + ([(y2, gy2)] if dual else [])
    torch.autograd.backward([t for t, _ in loss_parts], [nhwc(g, G) for _, g in loss_parts])
    # reference
    xr = bf(x).requires_grad_(True)
    gr = gam.clone().requires_grad_(True)
    btr = bet.clone().requires_grad_(True)
    ar = alpha.clone().requires_grad_(True)
    if mode == "bn":
        z = F.batch_norm(xr, None, None, gr, btr, True, 0.1, 1e-5)
    elif mode == "in":
        z = F.instance_norm(xr, eps=1e-5)
    else:
        z = xr
    mask = 1.0
    if gated:
        hard = (torch.sign(ar.detach() - 0.5) + 1) / 2
        mask = (hard + ar - ar.detach())[None, :, None, None]
    actf = (lambda t: F.leaky_relu(t, 0.2)) if act == 1 else F.relu
    if mode == "id" and gated:
        yr = actf(z) * mask
    else:
        yr = actf(z * mask)
    outs, grads = [yr], [bf(gy)]
    if dual:
        outs.append(F.relu(z * mask))
        grads.append(bf(gy2))
    torch.autograd.backward(outs, grads)
    assert rel_err(nchw(y.detach(), C, G), yr.detach()) < BF16_TOL
    if dual:
        assert rel_err(nchw(y2.detach(), C, G), outs[1].detach()) < BF16_TOL
    assert rel_err(nchw(xh.grad, C, G), xr.grad) < 2 * BF16_TOL
    if mode == "bn":
        assert rel_err(layer.gamma.grad, gr.grad) < 5e-3
        assert rel_err(layer.beta.grad, btr.grad) < 5e-3
        m = bf(x).mean(dim=(0, 2, 3))
        v = bf(x).var(dim=(0, 2, 3), unbiased=True)
        assert torch.allclose(layer.running_mean, 0.1 * m, rtol=1e-3, atol=1e-4)
        assert torch.allclose(layer.running_var, 0.9 + 0.1 * v, rtol=1e-3, atol=1e-4)
    if gated:
        assert rel_err(layer.alpha.grad, ar.grad) < 5e-3
    if C % 8:
        assert bool((y.detach()[..., C:] == 0).all())


def test_bn_eval_mode(G):
    C = 24
    A = G.arena.ParamArena("cuda")
    layer = G.nets.NormLayer(A, "n", C, "bn", "cuda")
    A.finalize()
    layer.bind()
    with torch.no_grad():
        layer.gamma.copy_(rnd((C,), 1, 0.3) + 1)
        layer.beta.copy_(rnd((C,), 2))
        layer.running_mean.copy_(rnd((C,), 3))
        layer.running_var.copy_(rnd((C,), 4).abs() + 0.5)
    layer.training = False
    x = rnd((2, C, 5, 6), 5)
    with torch.no_grad():
        y = layer(nhwc(x, G), 2)
    ref = F.relu(F.batch_norm(bf(x), layer.running_mean, layer.running_var, layer.gamma, layer.beta, False, 0.1, 1e-5))
    assert rel_err(nchw(y, C, G), ref) < BF16_TOL


# ----------------------------------------------------------------------------------- elementwise
def test_layout_roundtrip_and_cat(G):
    x = rnd((2, 5, 6, 7), 1)
    h = nhwc(x, G)
    assert h.shape == (2, 6, 7, 8) and bool((h[..., 5:] == 0).all())
    assert torch.equal(nchw(h, 5, G), bf(x))
    a, b = rnd((2, 3, 6, 7), 2), rnd((2, 11, 6, 7), 3)
    ah, bh = nhwc(a, G).requires_grad_(True), nhwc(b, G).requires_grad_(True)
    c = G.ops.CatFn.apply(ah, bh, 3, 11)
    assert torch.equal(nchw(c.detach(), 14, G), bf(torch.cat([a, b], 1)))
    assert bool((c.detach()[..., 14:] == 0).all())
    g = rnd((2, 14, 6, 7), 4)
    c.backward(nhwc(g, G))
    assert torch.equal(nchw(ah.grad, 3, G), bf(g[:, :3]))
    assert torch.equal(nchw(bh.grad, 11, G), bf(g[:, 3:]))
    # aligned fast path
    a, b = rnd((2, 16, 6, 7), 5), rnd((2, 8, 6, 7), 6)
    ah, bh = nhwc(a, G).requires_grad_(True), nhwc(b, G).requires_grad_(True)
    c = G.ops.CatFn.apply(ah, bh, 16, 8)
    assert torch.equal(nchw(c.detach(), 24, G), bf(torch.cat([a, b], 1)))
    g = rnd((2, 24, 6, 7), 7)
    c.backward(nhwc(g, G))
    assert torch.equal(nchw(ah.grad.contiguous(), 16, G), bf(g[:, :16]))
    assert torch.equal(nchw(bh.grad.contiguous(), 8, G), bf(g[:, 16:]))


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_activation(G, mode):
    x = rnd((2, 16, 5, 5), 1)
    xh = nhwc(x, G).requires_grad_(True)
    y = G.ops.ActFn.apply(xh, mode, 0.2)
    g = rnd((2, 16, 5, 5), 2)
    y.backward(nhwc(g, G))
    xr = bf(x).requires_grad_(True)
    yr = [None, F.leaky_relu(xr, 0.2), F.relu(xr), torch.tanh(xr)][mode]
    yr.backward(bf(g))
    assert rel_err(nchw(y.detach(), 16, G), yr.detach()) < BF16_TOL
    assert rel_err(nchw(xh.grad, 16, G), xr.grad) < 2 * BF16_TOL


def test_dropout_statistics_and_replay(G):
    x = torch.ones(4, 32, 32, 64, dtype=torch.bfloat16, device="cuda").requires_grad_(True)
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    y = G.ops.DropoutFn.apply(x, 0.5, seed, 3)
    keep = (y.detach() != 0).float().mean().item()
    assert abs(keep - 0.5) < 0.01
    assert set(y.detach().unique().tolist()) == {0.0, 2.0}
    y.backward(torch.ones_like(y))
    assert torch.equal(x.grad, y.detach())  # same mask, same scale
    y2 = G.ops.DropoutFn.apply(x, 0.5, seed + 1, 3)
    assert not torch.equal(y2.detach(), y.detach())


def test_reflect_pad_and_dwconv(G):
    C = 24
    x = rnd((2, C, 9, 11), 1)
    xh = nhwc(x, G).requires_grad_(True)
    y = G.ops.ReflectPadFn.apply(xh, 3)
    g = rnd((2, C, 15, 17), 2)
    y.backward(nhwc(g, G))
    xr = bf(x).requires_grad_(True)
    yr = F.pad(xr, (3, 3, 3, 3), mode="reflect")
    yr.backward(bf(g))
    assert torch.equal(nchw(y.detach(), C, G), yr.detach())
    assert rel_err(nchw(xh.grad, C, G), xr.grad) < BF16_TOL
    # depthwise 3x3 with fused reflection pad
    A = G.arena.ParamArena("cuda")
    dw = G.nets.DwConvLayer(A, "dw", C)
    A.finalize()
    dw.bind()
    w0, b0 = rnd((C, 1, 3, 3), 3, 0.3), rnd((C,), 4)
    with torch.no_grad():
        dw.weight.copy_(w0)
        dw.bias.copy_(b0)
    xh = nhwc(x, G).requires_grad_(True)
    y = dw(xh)
    g = rnd((2, C, 9, 11), 5)
    A.zero_grad()
    y.backward(nhwc(g, G))
    xr = bf(x).requires_grad_(True)
    wr, br = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    yr = F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), wr, br, groups=C)
    yr.backward(bf(g))
    assert rel_err(nchw(y.detach(), C, G), yr.detach()) < BF16_TOL
    assert rel_err(nchw(xh.grad, C, G), xr.grad) < 2 * BF16_TOL
    assert rel_err(dw.weight.grad, wr.grad) < 5e-3
    assert rel_err(dw.bias.grad, br.grad) < 5e-3


# ----------------------------------------------------------------------------------------- losses
@pytest.mark.parametrize("mode", ["hinge", "lsgan", "vanilla", "wgangp"])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_gan_loss(G, mode, kind):
    if kind == 2 and mode != "hinge":
        pytest.skip("generator form only differs for hinge")
    pred = rnd((3, 1, 30, 30), 7, 1.5)
    ph = nhwc(pred, G).requires_grad_(True)
    loss = G.ops.GanLossFn.apply(ph, 1, G.ops.GAN_MODES[mode], kind)
    (loss * 3.0).backward()
    pr = bf(pred).requires_grad_(True)
    if mode == "hinge":
        z = torch.zeros_like(pr)
        lr = [-torch.min(pr - 1, z).mean(), -torch.min(-pr - 1, z).mean(), -pr.mean()][kind]
    elif mode == "lsgan":
        lr = ((pr - (0.0 if kind == 1 else 1.0)) ** 2).mean()
    elif mode == "vanilla":
        lr = F.binary_cross_entropy_with_logits(pr, torch.full_like(pr, 0.0 if kind == 1 else 1.0))
    else:
        lr = pr.mean() if kind == 1 else -pr.mean()
    (lr * 3.0).backward()
    assert abs(loss.item() - lr.item()) < 1e-4 * max(1.0, abs(lr.item()))
    assert rel_err(nchw(ph.grad, 1, G), pr.grad) < BF16_TOL
    assert bool((ph.grad[..., 1:] == 0).all())


def test_l1_and_rmse(G):
    for C in (3, 16):
        a, b = rnd((2, C, 8, 9), 1), rnd((2, C, 8, 9), 2)
        for mode in (0, 1):
            ah = nhwc(a, G).requires_grad_(True)
            loss = G.ops.DiffLossFn.apply(ah, nhwc(b, G), C, mode)
            (loss * 2.0).backward()
            ar = bf(a).requires_grad_(True)
            lr = (ar - bf(b)).abs().mean() if mode == 0 else torch.sqrt(((ar - bf(b)) ** 2).mean())
            (lr * 2.0).backward()
            assert abs(loss.item() - lr.item()) < 1e-4 * abs(lr.item())
            assert rel_err(nchw(ah.grad, C, G), ar.grad) < BF16_TOL


@pytest.mark.parametrize("C,H,W", [(16, 8, 8), (40, 31, 31), (128, 16, 16)])
def test_gram_rmse(G, C, H, W):
    f, t = rnd((3, C, H, W), 1), rnd((3, C, H, W), 2)

    def gram(x):
        b, c, h, w = x.shape
        v = x.reshape(b, c, h * w)
        return torch.bmm(v, v.transpose(1, 2)) / (c * h * w)

    fh = nhwc(f, G).requires_grad_(True)
    gt = G.ops.gram_matrix(nhwc(t, G), C)
    assert rel_err(gt, gram(bf(t))) < 1e-4
    loss = G.ops.GramRmseFn.apply(fh, gt, C)
    (loss * 1e4).backward()
    fr = bf(f).requires_grad_(True)
    lr = torch.sqrt(((gram(fr) - gram(bf(t))) ** 2).mean())
    (lr * 1e4).backward()
    assert abs(loss.item() - lr.item()) < 1e-3 * abs(lr.item())
    assert rel_err(nchw(fh.grad, C, G), fr.grad) < 2 * BF16_TOL


# ------------------------------------------------------------------------------------------ adam
def test_adam_arena_matches_torch(G):
    A = G.arena.ParamArena("cuda", lr=2e-4, betas=(0.5, 0.999))
    A.add("w", (8, 5, 3, 3), "conv")
    A.add("b", (13,), "vec")
    A.finalize()
    w0, b0 = rnd((8, 5, 3, 3), 1), rnd((13,), 2)
    with torch.no_grad():
        A.params["w"].copy_(w0)
        A.params["b"].copy_(b0)
    wr, br = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    ref = torch.optim.Adam([wr, br], lr=2e-4, betas=(0.5, 0.999))
    for i in range(4):
        gw, gb = rnd((8, 5, 3, 3), 10 + i, 10.0 ** (i - 2)), rnd((13,), 20 + i, 10.0 ** (i - 2))
        A.zero_grad()
        A.params["w"].grad.copy_(gw)
        A.params["b"].grad.copy_(gb)
        A.step()
        wr.grad, br.grad = gw.clone(), gb.clone()
        ref.step()
    assert torch.allclose(A.params["w"], wr, rtol=1e-5, atol=1e-7)
    assert torch.allclose(A.params["b"], br, rtol=1e-5, atol=1e-7)
    A.ensure_packed()
    pk = A.packs["w"]
    assert torch.equal(pk.direct[..., :5].float(), bf(A.params["w"].detach().permute(0, 2, 3, 1).reshape(8, 9, 5)))
    assert torch.equal(pk.transposed[..., :8].float(), bf(A.params["w"].detach().permute(1, 2, 3, 0).reshape(5, 9, 8)))


# ------------------------------------------------------------------------------------ slab kernels
@pytest.mark.parametrize("N,H,W,C", [(2, 16, 16, 24), (3, 9, 7, 13), (2, 64, 64, 40), (1, 2, 3, 8)])
def test_dw_in_slab_matches_torch(G, N, H, W, C):
    """InstanceNorm(dw3x3(ReflectionPad(x)) + b) in one launch and its one-launch backward (csrc/slab.cu)."""
    A = G.arena.ParamArena("cuda")
    dw = G.nets.DwConvLayer(A, "dw", C)
    A.finalize()
    dw.bind()
    w0, b0 = rnd((C, 1, 3, 3), 1, 0.5), rnd((C,), 2)
    with torch.no_grad():
        dw.weight.copy_(w0)
        dw.bias.copy_(b0)
    x = rnd((N, C, H, W), 3)
    xh = nhwc(x, G).requires_grad_(True)
    z = G.ops.DwInSlabFn.apply(xh, dw.weight, dw.bias, dw)
    gz = rnd((N, C, H, W), 4)
    A.zero_grad()
    z.backward(nhwc(gz, G))
    xr, wr, br = bf(x).requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    yr = F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), wr, br, groups=C)
    zr = F.instance_norm(yr, eps=1e-5)
    zr.backward(bf(gz))
    assert rel_err(nchw(z.detach(), C, G), zr.detach()) < BF16_TOL
    assert rel_err(nchw(xh.grad, C, G), xr.grad) < 1.2e-2     # (dy1 is rounded to bf16 before the data-gradient taps)
    assert rel_err(dw.weight.grad, wr.grad) < 1e-2
    if C % 8:
        assert bool((z.detach()[..., C:] == 0).all())


@pytest.mark.parametrize("N,H,W,C,act,res", [(2, 16, 16, 24, 2, False), (3, 9, 7, 13, 0, True), (2, 64, 64, 40, 1, False),
                                             (2, 31, 31, 64, 1, False)])
def test_in_act_slab_matches_torch(G, N, H, W, C, act, res):
    y = rnd((N, C, H, W), 5, 2.0) + 0.3
    yh = nhwc(y, G).requires_grad_(True)
    r = rnd((N, C, H, W), 6)
    rh = nhwc(r, G).requires_grad_(True) if res else None
    z = G.ops.InActSlabFn.apply(yh, rh, C, act, 0.2)
    gz = rnd((N, C, H, W), 7)
    z.backward(nhwc(gz, G))
    yr = bf(y).requires_grad_(True)
    rr = bf(r).requires_grad_(True)
    t = F.instance_norm(yr, eps=1e-5)
    t = F.leaky_relu(t, 0.2) if act == 1 else (F.relu(t) if act == 2 else t)
    zr = t + rr if res else t
    zr.backward(bf(gz))
    assert rel_err(nchw(z.detach(), C, G), zr.detach()) < BF16_TOL
    assert rel_err(nchw(yh.grad, C, G), yr.grad) < 1.2e-2
    if res:
        assert rel_err(nchw(rh.grad, C, G), rr.grad) < BF16_TOL
