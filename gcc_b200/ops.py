"""autograd.Function wrappers over the C-ABI kernels.  Activations are NHWC bf16 tensors
[N, H, W, Cp] (Cp = channels padded to a multiple of 8, pad channels are exactly zero).

Parameter gradients are not returned to autograd: the kernels accumulate them straight into the
owning ``ParamArena``'s flat gradient buffer (zeroed by ``arena.zero_grad()``), honouring the
parameter's ``requires_grad`` flag exactly like the reference's set_requires_grad toggles
(models/Pix2Pix.py:574-590).
"""
import torch

from . import _lib
from .arena import rp8

call = _lib.call
BN_EPS = 1e-5
BN_MOM = 0.1

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3

# Global-batch parity under data parallel, off by default: when set (models built with --sync_bn
# inside an initialised process group), BatchNorm statistics (forward: sum x, sum x^2; backward: sum dg, sum dg*xhat)
# and every loss mean are all-reduced BEFORE any non-linearity that follows (sqrt of the MSE / Gram MSE, |.| of the
# arch-step loss differences), so that N ranks x batch b compute the step of one device at batch N*b (running
# statistics, the teacher EMA scalar and the reported losses then stay replicated).
GLOBAL_BATCH_SYNC = False


def _world():
    return torch.distributed.get_world_size() if GLOBAL_BATCH_SYNC else 1


def _allreduce_sum(t):
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)


def _global_mean(t):
    """Per-rank mean (device scalar, in place) -> mean over the global batch.  The backward kernels keep the per-rank
    scaling: every rank holds the gradient of the GLOBAL loss times the world size, the convention under which
    averaging the parameter gradients over ranks (the data-parallel all-reduce) yields the global-batch gradient."""
    if GLOBAL_BATCH_SYNC:
        _allreduce_sum(t)
        t.mul_(1.0 / _world())
    return t



def _st():
    return torch.cuda.current_stream().cuda_stream


def _check(x):
    if not x.is_cuda:
        raise _lib.GccB200Error("gcc_b200 ops need CUDA tensors: there is no CPU fallback")
    return x


def _window(t):
    """(pointer, channel stride) of an NHWC gradient that may be a channel-window view."""
    n, h, w, c = t.shape
    if t.stride(3) == 1:
        cs = t.stride(2)
        if t.stride(1) == w * cs and t.stride(0) == h * w * cs and cs % 8 == 0 and t.data_ptr() % 16 == 0:
            return t, cs
    t = t.contiguous()
    return t, t.shape[3]


_TAIL_WS = {}
_TAIL_WS_ELEMS = 148 * 128 * 256


def _tail_ws(device):
    """Per-device fp32 scratch of the conv kernel's tail-wave split (include/gcc_b200.h, gcc_conv_gemm_bf16 (b)): 19 MB,
    allocated once (before any graph capture: the eager warm-up iterations touch it), contents undefined between calls.
    Every convolution of the step runs on ONE stream, so consecutive launches may share it."""
    ws = _TAIL_WS.get(device)
    if ws is None:
        ws = _TAIL_WS[device] = torch.empty(_TAIL_WS_ELEMS, dtype=torch.float32, device=device)
    return ws


def _splitk_ws(n, oh, ow, rows, device, stats=False):
    """fp32 scratch of a conv launch, (ptr, elems, keep-alive): layers with very few output pixels (U-Net inner levels)
    get a split-K buffer of their own (not with fused statistics), the large ones the shared tail-wave scratch."""
    pix = n * oh * ow
    if pix > 16384:
        ws = _tail_ws(device)
        return ws.data_ptr(), ws.numel(), ws
    if stats:
        return None, 0, None
    elems = pix * rp8(rows)
    ws = torch.empty(elems, dtype=torch.float32, device=device)
    return ws.data_ptr(), elems, ws


class _ZeroPool:
    """Zero-initialised fp32 scratch handed out by bump allocation and re-zeroed with ONE fill per training phase
    (``reset()`` at the start of optimize_parameters / optimizer_netD_arch) instead of one tiny fill kernel per
    BatchNorm statistics buffer (~90 per iteration).  Only for scratch that dies with the phase's autograd graph."""

    def __init__(self):
        self.bufs = {}

    def take(self, n, device):
        ent = self.bufs.get(device)
        if ent is None:
            ent = self.bufs[device] = [torch.zeros(1 << 22, dtype=torch.float32, device=device), 0]
        buf, off = ent
        n8 = (n + 7) // 8 * 8
        if off + n8 > buf.numel():
            return torch.zeros(n, dtype=torch.float32, device=device)
        ent[1] = off + n8
        return buf[off:off + n]

    def reset(self):
        for ent in self.bufs.values():
            if ent[1]:
                ent[0][:ent[1]].zero_()
                ent[1] = 0


zero_pool = _ZeroPool()


def conv_out_hw(h, w, k, stride, pad, transposed, outpad=0):
    if not transposed:
        return (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    return (h - 1) * stride - 2 * pad + k + outpad, (w - 1) * stride - 2 * pad + k + outpad


class ConvFn(torch.autograd.Function):
    """nn.Conv2d / nn.ConvTranspose2d (+ bias, + fused LeakyReLU/Tanh epilogue) on tcgen05."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, act, slope, stats=None):
        _check(x)
        x = x.contiguous()
        layer.arena.ensure_packed()
        n, h, w, cx = x.shape
        tr = 1 if layer.kind == "convT" else 0
        oh, ow = conv_out_hw(h, w, layer.k, layer.stride, layer.pad, tr, layer.outpad)
        cop = rp8(layer.cout)
        y = torch.empty(n, oh, ow, cop, dtype=torch.bfloat16, device=x.device)
        pk = layer.packs
        wp = pk.direct if not tr else pk.transposed  # [cout][T][cin_p]
        epi = {ACT_NONE: 0, ACT_LRELU: 1, ACT_TANH: 2}[act]
        wsp, wse, _keep = _splitk_ws(n, oh, ow, layer.cout, x.device, stats is not None)
        call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cx, wp.data_ptr(), layer.cout, layer.k * layer.k,
             wp.shape[2], None if bias is None else bias.data_ptr(), y.data_ptr(), oh, ow, cop, 0, tr, layer.k,
             layer.k, layer.stride, layer.pad, epi, slope, 0, wsp, wse,
             None if stats is None else stats.data_ptr(), cop, _st())
        ctx.layer, ctx.act, ctx.slope = layer, act, slope
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        x, y = ctx.saved_tensors
        dy = dy.contiguous()
        st = _st()
        if ctx.act != ACT_NONE:
            dpre = torch.empty_like(dy)
            call("gcc_act_bwd_bf16", y.data_ptr(), dy.data_ptr(), dpre.data_ptr(), dy.numel(),
                 1 if ctx.act == ACT_LRELU else 3, ctx.slope, st)
        else:
            dpre = dy
        n, h, w, cx = x.shape
        _, oh, ow, cop = dpre.shape
        tr = layer.kind == "convT"
        T = layer.k * layer.k
        dx = None
        if ctx.needs_input_grad[0]:
            layer.arena.ensure_packed()
            pk = layer.packs
            wp = pk.transposed if not tr else pk.direct  # [cin][T][cout_p]
            dx = torch.empty(n, h, w, rp8(layer.cin), dtype=torch.bfloat16, device=x.device)
            wsp, wse, _keep = _splitk_ws(n, h, w, layer.cin, x.device)
            call("gcc_conv_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, wp.data_ptr(), layer.cin, T, wp.shape[2], None,
                 dx.data_ptr(), h, w, dx.shape[3], 0, 0 if tr else 1, layer.k, layer.k, layer.stride, layer.pad, 0,
                 0.0, 0, wsp, wse, None, 0, st)
            if dx.shape[3] != cx:
                raise _lib.GccB200Error("conv input channel padding mismatch")
        if ctx.needs_input_grad[1]:
            gw = layer.arena.flat_grad[layer.wname]
            if not tr:  # dW[co][tap][ci] = sum dy[.., co] * x[gather, ci]
                call("gcc_wgrad_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, x.data_ptr(), h, w, cx, gw.data_ptr(),
                     layer.cout, layer.cin, layer.k, layer.k, layer.stride, layer.pad, 0, 1, 1.0, st)
            else:  # dW[ci][tap][co] = sum x[.., ci] * dy[gather, co]
                call("gcc_wgrad_gemm_bf16", x.data_ptr(), n, h, w, cx, dpre.data_ptr(), oh, ow, cop, gw.data_ptr(),
                     layer.cin, layer.cout, layer.k, layer.k, layer.stride, layer.pad, 0, 1, 1.0, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = layer.arena.flat_grad[layer.bname]
            call("gcc_bias_grad_bf16", dpre.data_ptr(), n * oh * ow, cop, 0, layer.cout, gb.data_ptr(), 1, st)
        return dx, None, None, None, None, None, None


def _tap_major_weight(layer, pack, c, key):
    """Rows of ``pack`` ([c][16 taps][K] bf16) re-ordered tap-major with the channels padded to a group of 4
    or 8 (zero rows): the GEMM then writes columns col2im can read with one 8/16-byte load per tap.  Built by one
    kernel when the arena's packs were refreshed (cached per layer on the pack version)."""
    cg = 4 if c <= 4 else 8

    def build(old):
        buf = old if old is not None else torch.empty(16, cg, pack.shape[2], dtype=torch.bfloat16, device=pack.device)
        call("gcc_fold_weight_pack_bf16", pack.data_ptr(), buf.data_ptr(), 0, c, 16, cg, 0, pack.shape[2], 0, 16 * cg, _st())
        return buf
    return _derived(layer, key, build), cg, (2 if cg == 4 else 0)


class ColConvFn(torch.autograd.Function):
    """k4 s2 p1 layers whose image side has <= 8 channels.
    Conv2d (<= 8 input channels: first layer of the U-Net / PatchGANs): forward and weight gradient run in the GEMM
    kernels' image mode (the 16 taps x 8 channels are gathered by TMA into K = 128, no column buffer); the data
    gradient is one 1x1 GEMM producing tap-major columns + col2im.
    ConvTranspose2d (<= 8 output channels: last layer of the U-Net): forward = 1x1 GEMM + col2im; backward = im2col
    of the output gradient + 1x1 GEMMs (see elementwise.cu, "im2col / col2im")."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, act, slope):
        _check(x)
        x = x.contiguous()
        layer.arena.ensure_packed()
        st = _st()
        n, h, w, cx = x.shape
        pk = layer.packs
        bp = None if bias is None else bias.data_ptr()
        if layer.kind == "conv":
            oh, ow = h // 2, w // 2
            cop = rp8(layer.cout)
            # the GEMM kernel gathers the 16 taps x 8 channels itself (image mode: K = 128 straight from the image)
            y = torch.empty(n, oh, ow, cop, dtype=torch.bfloat16, device=x.device)
            epi = {ACT_NONE: 0, ACT_LRELU: 1, ACT_TANH: 2}[act]
            call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, 8, pk.direct.data_ptr(), layer.cout, 16, 8, bp,
                 y.data_ptr(), oh, ow, cop, 0, 0, 4, 4, 2, 1, epi, slope, 0, None, 0, None, 0, st)
            saved = x
        else:
            oh, ow = 2 * h, 2 * w
            wp, cg, order = _tap_major_weight(layer, pk.transposed, layer.cout, "tapmajor_fwd")  # [16*cg][cin_p]
            ycol = torch.empty(n, h, w, 16 * cg, dtype=torch.bfloat16, device=x.device)
            call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cx, wp.data_ptr(), 16 * cg, 1, wp.shape[2], None,
                 ycol.data_ptr(), h, w, 16 * cg, 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
            y = torch.empty(n, oh, ow, 8, dtype=torch.bfloat16, device=x.device)
            if act not in (ACT_NONE, ACT_TANH):
                raise _lib.GccB200Error("col-path ConvTranspose supports none/tanh epilogues")
            call("gcc_col2im_k4s2_c8", ycol.data_ptr(), 16 * cg, order, layer.cout, bp, 2 if act == ACT_TANH else 0,
                 y.data_ptr(), n, oh, ow, st)
            saved = x
        ctx.layer, ctx.act, ctx.slope = layer, act, slope
        ctx.has_bias = bias is not None
        ctx.xshape = (n, h, w, cx)
        ctx.save_for_backward(saved, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        saved, y = ctx.saved_tensors
        dy = dy.contiguous()
        st = _st()
        dev = dy.device
        if ctx.act != ACT_NONE:
            dpre = torch.empty_like(dy)
            call("gcc_act_bwd_bf16", y.data_ptr(), dy.data_ptr(), dpre.data_ptr(), dy.numel(),
                 1 if ctx.act == ACT_LRELU else 3, ctx.slope, st)
        else:
            dpre = dy
        n, h, w, cx = ctx.xshape
        pk = layer.packs
        dx = None
        if layer.kind == "conv":
            _, oh, ow, cop = dpre.shape
            if ctx.needs_input_grad[0]:
                layer.arena.ensure_packed()
                wp, cg, order = _tap_major_weight(layer, pk.transposed, layer.cin, "tapmajor_dgrad")  # [16*cg][cout_p]
                dcol = torch.empty(n, oh, ow, 16 * cg, dtype=torch.bfloat16, device=dev)
                call("gcc_conv_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, wp.data_ptr(), 16 * cg, 1,
                     wp.shape[2], None, dcol.data_ptr(), oh, ow, 16 * cg, 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0,
                     st)
                dx = torch.empty(n, h, w, 8, dtype=torch.bfloat16, device=dev)
                call("gcc_col2im_k4s2_c8", dcol.data_ptr(), 16 * cg, order, layer.cin, None, 0, dx.data_ptr(), n, h, w,
                     st)
            if ctx.needs_input_grad[1]:
                tmp = torch.empty(layer.cout, 128, dtype=torch.float32, device=dev)
                call("gcc_wgrad_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, saved.data_ptr(), h, w, 8,
                     tmp.data_ptr(), layer.cout, 8, 4, 4, 2, 1, 0, 0, 1.0, st)
                call("gcc_unpad_wgrad_c8", tmp.data_ptr(), layer.arena.flat_grad[layer.wname].data_ptr(), layer.cout,
                     layer.cin, st)
            npix_out = n * oh * ow
            bias_cp = cop
        else:
            x = saved
            dcol = torch.empty(n, h, w, 128, dtype=torch.bfloat16, device=dev)
            call("gcc_im2col_k4s2_c8", dpre.data_ptr(), dcol.data_ptr(), n, 2 * h, 2 * w, st)
            if ctx.needs_input_grad[0]:
                layer.arena.ensure_packed()
                wp = pk.direct  # [cin][16][8] viewed as [cin][1][128]
                dx = torch.empty(n, h, w, rp8(layer.cin), dtype=torch.bfloat16, device=dev)
                call("gcc_conv_gemm_bf16", dcol.data_ptr(), n, h, w, 128, wp.data_ptr(), layer.cin, 1, 128, None,
                     dx.data_ptr(), h, w, dx.shape[3], 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
            if ctx.needs_input_grad[1]:
                tmp = torch.empty(layer.cin, 128, dtype=torch.float32, device=dev)
                call("gcc_wgrad_gemm_bf16", x.data_ptr(), n, h, w, cx, dcol.data_ptr(), h, w, 128, tmp.data_ptr(),
                     layer.cin, 128, 1, 1, 1, 0, 0, 0, 1.0, st)
                call("gcc_unpad_wgrad_c8", tmp.data_ptr(), layer.arena.flat_grad[layer.wname].data_ptr(), layer.cin,
                     layer.cout, st)
            npix_out = n * 4 * h * w
            bias_cp = 8
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = layer.arena.flat_grad[layer.bname]
            call("gcc_bias_grad_bf16", dpre.data_ptr(), npix_out, bias_cp, 0, layer.cout, gb.data_ptr(), 1, st)
        return dx, None, None, None, None, None


class HeadConvFn(torch.autograd.Function):
    """k4 s1 p1 Conv2d with <= 8 output channels (the PatchGAN logits head): one 1x1 GEMM over the input
    (read once) producing the 16 per-tap partial dot products, then a fold (see elementwise.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, act, slope):
        _check(x)
        if act != ACT_NONE:
            raise _lib.GccB200Error("head conv has no fused activation")
        x = x.contiguous()
        layer.arena.ensure_packed()
        st = _st()
        n, h, w, cx = x.shape
        pk = layer.packs
        rows = layer.cout * 16
        ycol = torch.empty(n, h, w, rp8(rows), dtype=torch.bfloat16, device=x.device)
        wp = pk.direct  # [cout][16][cin_p] viewed as [cout*16][1][cin_p]
        call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cx, wp.data_ptr(), rows, 1, wp.shape[2], None, ycol.data_ptr(),
             h, w, ycol.shape[3], 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
        y = torch.empty(n, h - 1, w - 1, 8, dtype=torch.bfloat16, device=x.device)
        call("gcc_fold_k4s1_c8", ycol.data_ptr(), ycol.shape[3], layer.cout, None if bias is None else bias.data_ptr(),
             y.data_ptr(), n, h, w, st)
        ctx.layer = layer
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        st = _st()
        n, h, w, cx = x.shape
        dev = dy.device
        dcol = torch.empty(n, h, w, 128, dtype=torch.bfloat16, device=dev)
        call("gcc_unfold_k4s1_c8", dy.data_ptr(), dcol.data_ptr(), n, h, w, st)
        dx = None
        if ctx.needs_input_grad[0]:
            layer.arena.ensure_packed()
            wp = layer.packs.transposed  # [cin][16][8] viewed as [cin][1][128]
            dx = torch.empty(n, h, w, rp8(layer.cin), dtype=torch.bfloat16, device=dev)
            call("gcc_conv_gemm_bf16", dcol.data_ptr(), n, h, w, 128, wp.data_ptr(), layer.cin, 1, 128, None,
                 dx.data_ptr(), h, w, dx.shape[3], 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
        if ctx.needs_input_grad[1]:
            tmp = torch.empty(128, layer.cin, dtype=torch.float32, device=dev)
            call("gcc_wgrad_gemm_bf16", dcol.data_ptr(), n, h, w, 128, x.data_ptr(), h, w, cx, tmp.data_ptr(), 128,
                 layer.cin, 1, 1, 1, 0, 0, 0, 1.0, st)
            call("gcc_unpad_wgrad_rows", tmp.data_ptr(), layer.arena.flat_grad[layer.wname].data_ptr(), layer.cout,
                 layer.cin, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = layer.arena.flat_grad[layer.bname]
            call("gcc_bias_grad_bf16", dy.data_ptr(), n * (h - 1) * (w - 1), 8, 0, layer.cout, gb.data_ptr(), 1, st)
        return dx, None, None, None, None, None


def alloc_slack(shape, device, zero=False):
    """bf16 tensor of `shape` whose storage continues for >= 128 bytes behind the last element (the row-window stem's
    tensor map reads up to 7 pixels past a row end, against zero weights)."""
    n = 1
    for d in shape:
        n *= int(d)
    flat = (torch.zeros if zero else torch.empty)(n + 64, dtype=torch.bfloat16, device=device)
    if not zero:
        flat[n:].zero_()
    return flat[:n].view(*shape)


def _has_slack(t):
    return t.is_contiguous() and t.untyped_storage().nbytes() - (t.storage_offset() + t.numel()) * 2 >= 128


def _derived(layer, key, build):
    """Per-layer operand derived from the arena's bf16 packs, rebuilt when the packs were refreshed."""
    layer.arena.ensure_packed()
    cache = layer.__dict__.setdefault("_derived_ops", {})
    ent = cache.get(key)
    if ent is None or ent[0] != layer.arena.version:
        buf = build(None if ent is None else ent[1])
        cache[key] = (layer.arena.version, buf)
        return buf
    return ent[1]


class StemConvFn(torch.autograd.Function):
    """Stride-1 Conv2d whose INPUT has <= 8 channels and whose kernel is large (k7 stem of the MobileResNet generator,
    k9 stem of the SRResNet, VGG's first k3 conv): row-window implicit GEMM (gcc_conv_rowwin_bf16) instead of one
    7/8-empty k-block per tap; the data gradient (needed when the input is itself generated, e.g. CycleGAN's
    G_B(G_A(x))) is a 1x1 GEMM + fold."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, act, slope):
        _check(x)
        st = _st()
        n, h, w, cx = x.shape
        k, p = layer.k, layer.pad
        if cx != 8:
            raise _lib.GccB200Error("stem conv expects an 8-channel (padded) image")
        if p > 0:
            xp = alloc_slack((n, h + 2 * p, w + 2 * p, 8), x.device)
            call("gcc_zero_pad_bf16", x.contiguous().data_ptr(), xp.data_ptr(), n, h, w, 8, p, 0, 0, st)
        else:
            xp = x if _has_slack(x) else alloc_slack(x.shape, x.device).copy_(x)
        hp, wp = h + 2 * p, w + 2 * p
        oh, ow = hp - k + 1, wp - k + 1
        kb = (k + 7) // 8
        pk = layer.packs

        def build(old):
            buf = old if old is not None else torch.empty(layer.cout, k * kb, 64, dtype=torch.bfloat16, device=x.device)
            call("gcc_rowwin_weight_pack_bf16", pk.direct.data_ptr(), buf.data_ptr(), layer.cout, k, k, st)
            return buf
        wrow = _derived(layer, "rowwin", build)
        cop = rp8(layer.cout)
        y = torch.empty(n, oh, ow, cop, dtype=torch.bfloat16, device=x.device)
        epi = {ACT_NONE: 0, ACT_LRELU: 1, ACT_TANH: 2}[act]
        call("gcc_conv_rowwin_bf16", xp.data_ptr(), n, hp, wp, wrow.data_ptr(), layer.cout, k, k,
             None if bias is None else bias.data_ptr(), y.data_ptr(), oh, ow, cop, epi, slope, None, 0, st)
        ctx.layer, ctx.act, ctx.slope, ctx.has_bias = layer, act, slope, bias is not None
        ctx.geom = (n, h, w, hp, wp, oh, ow, cop)
        ctx.save_for_backward(xp, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        xp, y = ctx.saved_tensors
        n, h, w, hp, wp, oh, ow, cop = ctx.geom
        k, p = layer.k, layer.pad
        kb = (k + 7) // 8
        st = _st()
        dev = dy.device
        dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dpre = torch.empty_like(dy)
            call("gcc_act_bwd_bf16", y.data_ptr(), dy.data_ptr(), dpre.data_ptr(), dy.numel(),
                 1 if ctx.act == ACT_LRELU else 3, ctx.slope, st)
        else:
            dpre = dy
        dx = None
        if ctx.needs_input_grad[0]:
            cg = 4 if layer.cin <= 4 else 8
            T = k * k
            ccol = rp8(T * cg)
            pk = layer.packs

            def build(old):
                buf = old if old is not None else torch.empty(ccol, pk.d0p, dtype=torch.bfloat16, device=dev)
                call("gcc_fold_weight_pack_bf16", pk.transposed.data_ptr(), buf.data_ptr(), 0, layer.cin, T, cg, layer.cout,
                     pk.d0p, 0, ccol, st)
                return buf
            w2 = _derived(layer, "stem_dgrad", build)           # [(tap, c_in)][cout_p]
            ycol = torch.empty(n, oh, ow, ccol, dtype=torch.bfloat16, device=dev)
            call("gcc_conv_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, w2.data_ptr(), ccol, 1, w2.shape[1], None,
                 ycol.data_ptr(), oh, ow, ccol, 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
            dxp = torch.empty(n, hp, wp, 8, dtype=torch.bfloat16, device=dev)
            call("gcc_fold_taps_bf16", ycol.data_ptr(), ccol, cg, k, k, layer.cin, None, 0, dxp.data_ptr(), n, oh, ow, hp, wp,
                 -1, 0, st)
            if p > 0:
                dx = torch.empty(n, h, w, 8, dtype=torch.bfloat16, device=dev)
                call("gcc_zero_pad_bf16", dxp.data_ptr(), dx.data_ptr(), n, h, w, 8, p, 0, 1, st)
            else:
                dx = dxp
        if ctx.needs_input_grad[1]:
            tmp = torch.empty(layer.cout, k * kb, 64, dtype=torch.float32, device=dev)
            call("gcc_wgrad_rowwin_bf16", dpre.data_ptr(), n, oh, ow, cop, xp.data_ptr(), hp, wp, tmp.data_ptr(), layer.cout,
                 k, k, st)
            call("gcc_rowwin_wgrad_unpack_f32", tmp.data_ptr(), layer.arena.flat_grad[layer.wname].data_ptr(), layer.cout, k,
                 k, layer.cin, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            call("gcc_bias_grad_bf16", dpre.data_ptr(), n * oh * ow, cop, 0, layer.cout,
                 layer.arena.flat_grad[layer.bname].data_ptr(), 1, st)
        return dx, None, None, None, None, None


class FoldConvFn(torch.autograd.Function):
    """Stride-1 Conv2d with <= 8 OUTPUT channels and a large kernel (k7 head of the MobileResNet generator, k9 head of
    the SRResNet): ONE 1x1 GEMM over the input computes every tap's partial dot product, gcc_fold_taps_bf16 sums the
    shifted partials (+ bias, + tanh); backward = unfold of dy + two 1x1 GEMMs."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer, act, slope):
        _check(x)
        if act not in (ACT_NONE, ACT_TANH):
            raise _lib.GccB200Error("fold-path conv supports none / tanh epilogues")
        x = x.contiguous()
        st = _st()
        n, h, w, cx = x.shape
        k, p = layer.k, layer.pad
        T = k * k
        cg = 4 if layer.cout <= 4 else 8
        ccol = rp8(T * cg)
        oh, ow = h + 2 * p - k + 1, w + 2 * p - k + 1
        pk = layer.packs

        def build(old):
            buf = old if old is not None else torch.empty(ccol, pk.d1p, dtype=torch.bfloat16, device=x.device)
            call("gcc_fold_weight_pack_bf16", pk.direct.data_ptr(), buf.data_ptr(), 0, layer.cout, T, cg, layer.cin, pk.d1p,
                 0, ccol, st)
            return buf
        w2 = _derived(layer, "fold_fwd", build)                  # [(tap, c_out)][cin_p]
        ycol = torch.empty(n, h, w, ccol, dtype=torch.bfloat16, device=x.device)
        call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cx, w2.data_ptr(), ccol, 1, w2.shape[1], None, ycol.data_ptr(), h, w,
             ccol, 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
        y = torch.empty(n, oh, ow, 8, dtype=torch.bfloat16, device=x.device)
        call("gcc_fold_taps_bf16", ycol.data_ptr(), ccol, cg, k, k, layer.cout, None if bias is None else bias.data_ptr(),
             2 if act == ACT_TANH else 0, y.data_ptr(), n, h, w, oh, ow, 1, -p, st)
        ctx.layer, ctx.act, ctx.has_bias = layer, act, bias is not None
        ctx.geom = (n, h, w, cx, oh, ow, cg, ccol)
        ctx.save_for_backward(x, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        x, y = ctx.saved_tensors
        n, h, w, cx, oh, ow, cg, ccol = ctx.geom
        k, p = layer.k, layer.pad
        T = k * k
        st = _st()
        dev = dy.device
        dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dpre = torch.empty_like(dy)
            call("gcc_act_bwd_bf16", y.data_ptr(), dy.data_ptr(), dpre.data_ptr(), dy.numel(), 3, 0.0, st)
        else:
            dpre = dy
        dcol = torch.empty(n, h, w, ccol, dtype=torch.bfloat16, device=dev)
        call("gcc_unfold_taps_bf16", dpre.data_ptr(), dcol.data_ptr(), ccol, cg, k, k, n, h, w, oh, ow, -p, st)
        pk = layer.packs
        dx = None
        if ctx.needs_input_grad[0]:
            def build(old):
                buf = old if old is not None else torch.empty(layer.cin, ccol, dtype=torch.bfloat16, device=dev)
                call("gcc_fold_weight_pack_bf16", pk.transposed.data_ptr(), buf.data_ptr(), 1, layer.cout, T, cg, layer.cin, 0,
                     pk.d0p, ccol, st)
                return buf
            w3 = _derived(layer, "fold_dgrad", build)            # [cin][(tap, c_out)]
            dx = torch.empty(n, h, w, rp8(layer.cin), dtype=torch.bfloat16, device=dev)
            call("gcc_conv_gemm_bf16", dcol.data_ptr(), n, h, w, ccol, w3.data_ptr(), layer.cin, 1, ccol, None, dx.data_ptr(),
                 h, w, dx.shape[3], 0, 0, 1, 1, 1, 0, 0, 0.0, 0, None, 0, None, 0, st)
        if ctx.needs_input_grad[1]:
            tmp = torch.empty(ccol, layer.cin, dtype=torch.float32, device=dev)
            call("gcc_wgrad_gemm_bf16", dcol.data_ptr(), n, h, w, ccol, x.data_ptr(), h, w, cx, tmp.data_ptr(), ccol, layer.cin,
                 1, 1, 1, 0, 0, 0, 1.0, st)
            call("gcc_fold_wgrad_unpack_f32", tmp.data_ptr(), layer.arena.flat_grad[layer.wname].data_ptr(), layer.cout, T, cg,
                 layer.cin, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            call("gcc_bias_grad_bf16", dpre.data_ptr(), n * oh * ow, 8, 0, layer.cout,
                 layer.arena.flat_grad[layer.bname].data_ptr(), 1, st)
        return dx, None, None, None, None, None


class NormActFn(torch.autograd.Function):
    """[BatchNorm | InstanceNorm | identity] -> [channel gate] -> activation, with an optional second
    activation output (the U-Net's relu'd skip copy).

    ``y_into`` / ``y2_into`` = (buffer [N,H,W,Ct], channel offset): the output is written straight into that channel
    window of a wider NHWC buffer and returned as a view of it -- the U-Net writes both halves of every skip
    concatenation (models/Pix2Pix.py:77) this way, so ``torch.cat`` costs no copy (see ``CatViewFn``)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, alpha, layer, act, act2, sums_in=None, y_into=None, y2_into=None):
        _check(x)
        x = x.contiguous()
        n, h, w, cp = x.shape
        st = _st()
        mode = layer.mode  # 'bn', 'in', 'id'
        per_sample = 1 if mode == "in" else 0
        if y_into is not None and act2 is not None:
            raise _lib.GccB200Error("a windowed primary output excludes a second output")
        y = torch.empty_like(x) if y_into is None else None
        # kernel slots: (y pointer, second pointer + pitch + offset + activation)
        if y_into is not None:
            buf, coff = y_into
            k_y, k_y2, k_cp2, k_off2, k_act2 = None, buf.data_ptr(), buf.shape[3], coff, act
            yret = buf[..., coff:coff + cp]
            y2ret = None
        elif act2 is not None:
            if y2_into is not None:
                buf, coff = y2_into
                k_y2, k_cp2, k_off2 = buf.data_ptr(), buf.shape[3], coff
                y2ret = buf[..., coff:coff + cp]
            else:
                y2ret = torch.empty_like(x)
                k_y2, k_cp2, k_off2 = y2ret.data_ptr(), cp, 0
            k_y, k_act2, yret = y.data_ptr(), act2, y
        else:
            k_y, k_y2, k_cp2, k_off2, k_act2, yret, y2ret = y.data_ptr(), None, cp, 0, 0, y, None
        gp = None if gamma is None else gamma.data_ptr()
        bp = None if beta is None else beta.data_ptr()
        ap = None if alpha is None else alpha.data_ptr()
        sums = None
        if mode == "bn" and not layer.training:
            call("gcc_norm_apply_eval_bf16", x.data_ptr(), k_y, n, h * w, cp, layer.c,
                 layer.running_mean.data_ptr(), layer.running_var.data_ptr(), gp, bp, ap, layer.thr, BN_EPS, act,
                 layer.slope, k_y2, k_cp2, k_off2, k_act2, st)
            ctx.eval_bn = True
            ctx.stat_count = 0
        else:
            ctx.eval_bn = False
            if mode != "id":
                if sums_in is not None and not per_sample:
                    sums = sums_in  # accumulated by the producing conv's epilogue
                else:
                    sums = zero_pool.take((n if per_sample else 1) * 2 * cp, x.device)
                    call("gcc_norm_stats_acc_bf16", x.data_ptr(), n, h * w, cp, per_sample, sums.data_ptr(), st)
                if layer.stats_hook is not None:
                    layer.stats_hook(sums)
                if GLOBAL_BATCH_SYNC and mode == "bn":
                    sums = sums.clone()      # the buffer belongs to the shared zero pool
                    _allreduce_sum(sums)
            stat_count = n * h * w * _world() if mode == "bn" else 0
            rm = rv = None
            if mode == "bn" and layer.running_mean is not None:
                rm, rv = layer.running_mean.data_ptr(), layer.running_var.data_ptr()
                layer.num_batches += 1
            call("gcc_norm_apply_bf16", x.data_ptr(), k_y, n, h * w, cp, layer.c, per_sample,
                 None if sums is None else sums.data_ptr(), gp, bp, ap, layer.thr, BN_EPS, rm, rv, BN_MOM, act,
                 layer.slope, 1 if (getattr(layer, "gate_after", mode == "id") and mode == "id" and alpha is not None) else 0,
                 k_y2, k_cp2, k_off2, k_act2, stat_count, st)
            ctx.stat_count = stat_count
        ctx.layer, ctx.act, ctx.act2 = layer, act, act2
        ctx.save_for_backward(x, sums, gamma, beta, alpha)
        ctx.set_materialize_grads(False)
        if y2ret is None:
            return yret
        return yret, y2ret

    @staticmethod
    def backward(ctx, dy, dy2=None):
        layer = ctx.layer
        if ctx.eval_bn:
            raise _lib.GccB200Error("backward through eval-mode batch norm is not implemented")
        x, sums, gamma, beta, alpha = ctx.saved_tensors
        n, h, w, cp = x.shape
        st = _st()
        per_sample = 1 if layer.mode == "in" else 0
        p1 = p2 = None
        c1 = c2 = cp
        if dy is not None:
            dy, c1 = _window(dy)
            p1 = dy.data_ptr()
        if dy2 is not None:
            dy2, c2 = _window(dy2)
            p2 = dy2.data_ptr()
        arena_of = lambda p: p._gcc_arena.flat_grad[p._gcc_name].data_ptr()
        dgamma = arena_of(gamma) if (gamma is not None and ctx.needs_input_grad[1]) else None
        dbeta = arena_of(beta) if (beta is not None and ctx.needs_input_grad[2]) else None
        dalpha = arena_of(alpha) if (alpha is not None and ctx.needs_input_grad[3]) else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        red = zero_pool.take((n if per_sample else 1) * 2 * cp, x.device)   # pre-zeroed: phase | 4 below
        gate_after = 1 if (getattr(layer, "gate_after", layer.mode == "id") and layer.mode == "id" and alpha is not None) else 0

        def bwd(phase, red_param):
            call("gcc_norm_bwd_bf16", x.data_ptr(), n, h * w, cp, layer.c, per_sample,
                 None if sums is None else sums.data_ptr(), None if gamma is None else gamma.data_ptr(),
                 None if beta is None else beta.data_ptr(), None if alpha is None else alpha.data_ptr(), layer.thr, BN_EPS,
                 ctx.act, layer.slope, gate_after, p1, c1, 0, p2, c2, 0, ctx.act2 or 0, red.data_ptr(),
                 None if dx is None else dx.data_ptr(), dgamma, dbeta, dalpha, ctx.stat_count, phase | 4,
                 None if red_param is None else red_param.data_ptr(), st)

        if GLOBAL_BATCH_SYNC and layer.mode == "bn":
            # sync batch norm: dx needs the GLOBAL sums, the parameter gradients this rank's share of them
            bwd(1, None)
            red_local = red.clone()
            _allreduce_sum(red)
            bwd(2, red_local)
        else:
            bwd(0, None)
        return dx, None, None, None, None, None, None, None, None, None


class CatViewFn(torch.autograd.Function):
    """torch.cat([a, b], channel) where a and b already ARE the two channel windows of ``buf`` (written there by their
    producers, ``NormActFn(..., y_into / y2_into)``): no copy in forward, channel-window views of the gradient in
    backward."""

    @staticmethod
    def forward(ctx, a, b, buf):
        ca, cb = a.shape[3], b.shape[3]
        ok = (a.data_ptr() == buf.data_ptr() and b.data_ptr() == buf.data_ptr() + 2 * ca and ca + cb == buf.shape[3]
              and a.stride() == b.stride() == buf.stride())
        if not ok:
            raise _lib.GccB200Error("CatViewFn: the inputs are not the two channel windows of the buffer")
        ctx.ca = ca
        return buf.view(buf.shape)

    @staticmethod
    def backward(ctx, dy):
        ca = ctx.ca
        return dy[..., :ca], dy[..., ca:], None


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode, slope):
        x = _check(x).contiguous()
        y = torch.empty_like(x)
        call("gcc_act_fwd_bf16", x.data_ptr(), y.data_ptr(), x.numel(), mode, slope, _st())
        ctx.mode, ctx.slope = mode, slope
        ctx.save_for_backward(y if mode == ACT_TANH else x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (ref,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        call("gcc_act_bwd_bf16", ref.data_ptr(), dy.data_ptr(), dx.data_ptr(), dy.numel(), ctx.mode, ctx.slope, _st())
        return dx, None, None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p): counter-based mask keyed by (*seed, salt); backward replays the same mask."""

    @staticmethod
    def forward(ctx, x, p, seed, salt):
        x = _check(x).contiguous()
        y = torch.empty_like(x)
        call("gcc_dropout_bf16", x.data_ptr(), y.data_ptr(), x.numel(), p, seed.data_ptr(), salt, _st())
        ctx.p, ctx.salt = p, salt
        ctx.save_for_backward(seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        (seed,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        call("gcc_dropout_bf16", dy.data_ptr(), dx.data_ptr(), dy.numel(), ctx.p, seed.data_ptr(), ctx.salt, _st())
        return dx, None, None, None


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _check(a).contiguous(), b.contiguous()
        y = torch.empty_like(a)
        call("gcc_add_bf16", a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _st())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class CatFn(torch.autograd.Function):
    """torch.cat([a, b], channel) on the logical channels (ca, cb)."""

    @staticmethod
    def forward(ctx, a, b, ca, cb):
        a, b = _check(a).contiguous(), b.contiguous()
        n, h, w, _ = a.shape
        ct = rp8(ca + cb)
        st = _st()
        ctx.dims = (ca, cb, a.shape[3], b.shape[3])
        if ct == 8 and a.shape[3] == 8 and b.shape[3] == 8:      # both images and the result are one vector per pixel
            y = torch.empty(n, h, w, 8, dtype=torch.bfloat16, device=a.device)
            call("gcc_cat_small_bf16", a.data_ptr(), b.data_ptr(), y.data_ptr(), ca, cb, n * h * w, st)
            return y
        alloc = torch.empty if ct == ca + cb else torch.zeros
        y = alloc(n, h, w, ct, dtype=torch.bfloat16, device=a.device)
        call("gcc_copy_channels_bf16", a.data_ptr(), a.shape[3], 0, y.data_ptr(), ct, 0, ca, n * h * w, 0, st)
        call("gcc_copy_channels_bf16", b.data_ptr(), b.shape[3], 0, y.data_ptr(), ct, ca, cb, n * h * w, 0, st)
        return y

    @staticmethod
    def backward(ctx, dy):
        ca, cb, cap, cbp = ctx.dims
        dy = dy.contiguous()
        n, h, w, ct = dy.shape
        st = _st()
        da = db = None
        if ct == 8 and cap == 8 and cbp == 8:
            if ctx.needs_input_grad[0]:
                da = torch.empty(n, h, w, 8, dtype=torch.bfloat16, device=dy.device)
            if ctx.needs_input_grad[1]:
                db = torch.empty(n, h, w, 8, dtype=torch.bfloat16, device=dy.device)
            if da is not None or db is not None:
                call("gcc_split_small_bf16", dy.data_ptr(), None if da is None else da.data_ptr(),
                     None if db is None else db.data_ptr(), ca, cb, n * h * w, st)
            return da, db, None, None
        if ctx.needs_input_grad[0]:
            if ca == cap and ca % 8 == 0:
                da = dy[..., :ca]
            else:
                da = torch.zeros(n, h, w, cap, dtype=torch.bfloat16, device=dy.device)
                call("gcc_copy_channels_bf16", dy.data_ptr(), ct, 0, da.data_ptr(), cap, 0, ca, n * h * w, 0, st)
        if ctx.needs_input_grad[1]:
            if cb == cbp and ca % 8 == 0 and cb % 8 == 0:
                db = dy[..., ca:ca + cb]
            else:
                db = torch.zeros(n, h, w, cbp, dtype=torch.bfloat16, device=dy.device)
                call("gcc_copy_channels_bf16", dy.data_ptr(), ct, ca, db.data_ptr(), cbp, 0, cb, n * h * w, 0, st)
        return da, db, None, None


class ReflectPadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad):
        x = _check(x).contiguous()
        n, h, w, cp = x.shape
        y = alloc_slack((n, h + 2 * pad, w + 2 * pad, cp), x.device)   # (slack: the row-window stem may read it)
        call("gcc_reflect_pad_bf16", x.data_ptr(), y.data_ptr(), n, h, w, cp, pad, 0, _st())
        ctx.pad, ctx.shape = pad, (n, h, w, cp)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, h, w, cp = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty(n, h, w, cp, dtype=torch.bfloat16, device=dy.device)
        call("gcc_reflect_pad_bf16", dy.data_ptr(), dx.data_ptr(), n, h, w, cp, ctx.pad, 1, _st())
        return dx, None


class DwConvFn(torch.autograd.Function):
    """Depthwise 3x3 conv with fused ReflectionPad2d(1) (SeparableConv2d's first conv)."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer):
        x = _check(x).contiguous()
        n, h, w, cp = x.shape
        y = torch.empty_like(x)
        call("gcc_dw3x3_fwd_bf16", x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
             y.data_ptr(), n, h, w, cp, layer.c, _st())
        ctx.layer = layer
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, h, w, cp = x.shape
        dx = dxp = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)      # (the mirrored border taps are folded into the kernel: no padded scratch)
        gw = layer.arena.flat_grad[layer.wname].data_ptr() if ctx.needs_input_grad[1] else None
        gb = layer.arena.flat_grad[layer.bname].data_ptr() if (layer.bname and ctx.needs_input_grad[2]) else None
        call("gcc_dw3x3_bwd_bf16", x.data_ptr(), dy.data_ptr(), weight.data_ptr(),
             None if dxp is None else dxp.data_ptr(), None if dx is None else dx.data_ptr(), gw, gb, n, h, w, cp,
             layer.c, 1, _st())
        return dx, None, None, None


def slab_ok(x):
    """The slab kernels (csrc/slab.cu) cover NHWC tensors whose H * W pixels of 8 channels fit one CTA (<= 4096)."""
    return x.is_cuda and x.dim() == 4 and 4 <= x.shape[1] * x.shape[2] <= 4096 and x.shape[1] >= 2 and x.shape[2] >= 2


class DwInSlabFn(torch.autograd.Function):
    """InstanceNorm2d(depthwise3x3(ReflectionPad2d(1)(x)) + b) in ONE launch (SeparableConv2d's conv.0 + conv.1,
    models/Pix2Pix.py:137-141); backward recomputes the depthwise output from x (saved: x and the 2 statistics per
    (sample, channel)) and produces dx, dw, db in one launch."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer):
        x = _check(x).contiguous()
        n, h, w, cp = x.shape
        z = torch.empty_like(x)
        stats = torch.empty(n, cp, 2, dtype=torch.float32, device=x.device)
        call("gcc_dw_in_slab_fwd_bf16", x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
             z.data_ptr(), stats.data_ptr(), n, h, w, cp, layer.c, BN_EPS, _st())
        ctx.layer = layer
        ctx.save_for_backward(x, weight, bias, stats)
        return z

    @staticmethod
    def backward(ctx, dz):
        layer = ctx.layer
        x, weight, bias, stats = ctx.saved_tensors
        dz = dz.contiguous()
        n, h, w, cp = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = layer.arena.flat_grad[layer.wname].data_ptr() if ctx.needs_input_grad[1] else None
        gb = layer.arena.flat_grad[layer.bname].data_ptr() if (layer.bname and ctx.needs_input_grad[2]) else None
        call("gcc_dw_in_slab_bwd_bf16", x.data_ptr(), dz.data_ptr(), weight.data_ptr(),
             None if bias is None else bias.data_ptr(), stats.data_ptr(), None if dx is None else dx.data_ptr(), gw, gb,
             n, h, w, cp, layer.c, _st())
        return dx, None, None, None


class InActSlabFn(torch.autograd.Function):
    """act(InstanceNorm2d(y)) (+ residual) in ONE launch (the norm after the pointwise conv, the ReLU, and the block's
    skip addition, models/Pix2Pix.py:166-197; also the InstanceNorm + LeakyReLU of CycleGAN's plain discriminator)."""

    @staticmethod
    def forward(ctx, y, res, c, act, slope):
        y = _check(y).contiguous()
        n, h, w, cp = y.shape
        z = torch.empty_like(y)
        stats = torch.empty(n, cp, 2, dtype=torch.float32, device=y.device)
        call("gcc_in_act_slab_fwd_bf16", y.data_ptr(), None if res is None else res.contiguous().data_ptr(), z.data_ptr(),
             stats.data_ptr(), n, h * w, cp, c, BN_EPS, act, slope, _st())
        ctx.args = (c, act, slope, res is not None)
        ctx.save_for_backward(y, stats)
        return z

    @staticmethod
    def backward(ctx, dz):
        y, stats = ctx.saved_tensors
        c, act, slope, has_res = ctx.args
        dz = dz.contiguous()
        n, h, w, cp = y.shape
        dy = None
        if ctx.needs_input_grad[0]:
            dy = torch.empty_like(y)
            call("gcc_in_act_slab_bwd_bf16", y.data_ptr(), dz.data_ptr(), stats.data_ptr(), dy.data_ptr(), n, h * w, cp, c,
                 act, slope, _st())
        return dy, (dz if (has_res and ctx.needs_input_grad[1]) else None), None, None, None


# ------------------------------------------------------------------------------------- losses
GAN_MODES = {"hinge": 0, "lsgan": 1, "vanilla": 2, "wgangp": 3}


class GanLossFn(torch.autograd.Function):
    """GANLoss.__call__ (models/GANLoss.py:38-59) on an NHWC bf16 prediction with c logical channels."""

    @staticmethod
    def forward(ctx, pred, c, mode, kind):
        pred = _check(pred).contiguous()
        out = torch.zeros((), dtype=torch.float32, device=pred.device)
        npix = pred.numel() // pred.shape[-1]
        call("gcc_gan_loss_fwd_bf16", pred.data_ptr(), npix, pred.shape[-1], c, mode, kind, out.data_ptr(), _st())
        _global_mean(out)
        ctx.args = (npix, c, mode, kind)
        ctx.save_for_backward(pred)
        return out

    @staticmethod
    def backward(ctx, gout):
        (pred,) = ctx.saved_tensors
        npix, c, mode, kind = ctx.args
        gout = gout.contiguous().float()
        dp = torch.empty_like(pred)
        call("gcc_gan_loss_bwd_bf16", pred.data_ptr(), npix, pred.shape[-1], c, mode, kind, gout.data_ptr(),
             dp.data_ptr(), _st())
        return dp, None, None, None


class DiffLossFn(torch.autograd.Function):
    """mode 0: mean |a - b| (nn.L1Loss);  mode 1: sqrt(mean (a - b)^2) (sqrt(nn.MSELoss));  mode 2: mean (a - b)^2
    (nn.MSELoss).  b is constant."""

    @staticmethod
    def forward(ctx, a, b, c, mode):
        a, b = _check(a).contiguous(), b.contiguous()
        npix = a.numel() // a.shape[-1]
        acc = torch.zeros((), dtype=torch.float32, device=a.device)
        st = _st()
        call("gcc_diff_reduce_bf16", a.data_ptr(), b.data_ptr(), npix, a.shape[-1], c, 1 if mode else 0, acc.data_ptr(), st)
        _global_mean(acc)               # mean over the global batch BEFORE the square root
        if mode == 1:
            out = torch.empty_like(acc)
            call("gcc_scalar_sqrt", acc.data_ptr(), out.data_ptr(), st)
        else:
            out = acc
        ctx.args = (npix, c, mode)
        ctx.save_for_backward(a, b, acc)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b, acc = ctx.saved_tensors
        npix, c, mode = ctx.args
        gout = gout.contiguous().float()
        da = torch.empty_like(a)
        call("gcc_diff_bwd_bf16", a.data_ptr(), b.data_ptr(), npix, a.shape[-1], c, mode, gout.data_ptr(),
             acc.data_ptr(), da.data_ptr(), _st())
        return da, None, None, None


def gram_matrix(f, c):
    """Per-sample f f^T / (c h w) as fp32 [N, c, c] (models/Pix2Pix.py:733-740) on the tensor cores."""
    f = _check(f).contiguous()
    n, h, w, cp = f.shape
    g = torch.empty(n, c, 1, c, dtype=torch.float32, device=f.device)
    call("gcc_wgrad_gemm_bf16", f.data_ptr(), n, h, w, cp, f.data_ptr(), h, w, cp, g.data_ptr(), c, c, 1, 1, 1, 0, 1, 0,
         1.0 / (c * h * w), _st())
    return g.view(n, c, c)


class GramRmseFn(torch.autograd.Function):
    """sqrt(MSE(gram(f), gram_target)) (models/Pix2Pix.py:542) or, with mse=True, MSE(gram(f), gram_target)
    (models/CycleGAN.py:513); gram_target is a constant fp32 [N,c,c]."""

    @staticmethod
    def forward(ctx, f, gt, c, mse=False):
        f = _check(f).contiguous()
        gs = gram_matrix(f, c)
        acc = torch.zeros((), dtype=torch.float32, device=f.device)
        st = _st()
        call("gcc_sqdiff_reduce_f32", gs.data_ptr(), gt.data_ptr(), gs.numel(), acc.data_ptr(), st)
        _global_mean(acc)
        if mse:
            out = acc
        else:
            out = torch.empty_like(acc)
            call("gcc_scalar_sqrt", acc.data_ptr(), out.data_ptr(), st)
        ctx.c, ctx.mse = c, mse
        ctx.save_for_backward(f, gs, gt, acc)
        return out

    @staticmethod
    def backward(ctx, gout):
        f, gs, gt, acc = ctx.saved_tensors
        c = ctx.c
        n, h, w, cp = f.shape
        st = _st()
        gout = gout.contiguous().float()
        m = torch.empty(n, c, cp, dtype=torch.bfloat16, device=f.device)
        call("gcc_gram_bwd_matrix", gs.data_ptr(), gt.data_ptr(), n, c, cp, 1.0 / (c * h * w), gout.data_ptr(),
             acc.data_ptr(), 1 if ctx.mse else 0, m.data_ptr(), st)
        df = torch.empty_like(f)
        call("gcc_conv_gemm_bf16", f.data_ptr(), n, h, w, cp, m.data_ptr(), c, 1, cp, None, df.data_ptr(), h, w, cp, 0,
             0, 1, 1, 1, 0, 0, 0.0, 1, None, 0, None, 0, st)
        return df, None, None, None


# ------------------------------------------------------------------------------- boundary helpers
def to_nhwc(x_nchw, out=None, c_off=0, cp=None):
    """NCHW fp32 (device) -> NHWC bf16 channel window; returns the NHWC tensor."""
    x_nchw = _check(x_nchw).contiguous().float()
    n, c, h, w = x_nchw.shape
    if out is None:
        cp = cp or rp8(c)
        out = torch.empty(n, h, w, cp, dtype=torch.bfloat16, device=x_nchw.device)
        zero_to = cp
    else:
        cp = out.shape[3]
        zero_to = c_off + c
    call("gcc_nchw_f32_to_nhwc_bf16", x_nchw.data_ptr(), out.data_ptr(), n, c, h * w, cp, c_off, zero_to, _st())
    return out


def to_nchw(x_nhwc, c, c_off=0):
    """NHWC bf16 -> NCHW fp32 (the reference's tensor convention at the model boundary)."""
    x_nhwc = _check(x_nhwc).contiguous()
    n, h, w, cp = x_nhwc.shape
    out = torch.empty(n, c, h, w, dtype=torch.float32, device=x_nhwc.device)
    call("gcc_nhwc_bf16_to_nchw_f32", x_nhwc.data_ptr(), out.data_ptr(), n, c, h * w, cp, c_off, 0, _st())
    return out
