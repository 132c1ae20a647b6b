"""SAGAN on the B200 kernels: spectral-normed generator / (gated) discriminator with self-attention and the
``SAGANModel`` wrapper (reference: /root/reference/models/SAGAN.py:14-764).  Class / method / attribute / state-dict
names follow the reference so that ``train.py`` drives this class unchanged.

Reference behaviours reproduced on purpose (see oracle/sagan_oracle.py, pinned against the reference):
  * one power iteration per forward updates weight_u / weight_v in place; a backward of an EARLIER forward reads the
    vectors' current values (the reference's ``.data =`` assignment has the same effect on autograd's saved leaves);
  * after ``set_requires_grad(netD, True)`` the discriminator's u / v carry a gradient and optimizer_D steps them;
  * the student's optimizers list the spectral-normed convs' and the attention convs' parameters twice: those
    arenas are stepped twice per ``step()``;
  * the student's reported ``G_GAN`` is the whole generator loss (in-place accumulation into an alias).
"""
import copy
import math
import os
from collections import OrderedDict

import torch

from . import _lib, ops
from .arena import ParamArena, rp8
from .nets import ConvLayer, NormLayer, _Tree
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, GAN_MODES, _check, _st, call, conv_out_hw
from .base import GccModelMixin, capturing
from .pix2pix import _allreduce_grads, get_scheduler
from .srgan import _merge, _SRNet

LRELU = 0.1


# ------------------------------------------------------------------------------------------- autograd functions
class SNConvFn(torch.autograd.Function):
    """SpectralNorm(nn.Conv2d | nn.ConvTranspose2d) (SAGAN.py:14-71): power iteration, w = w_bar / sigma packed to
    bf16 for THIS forward, tcgen05 conv; backward: data gradient with the saved packs, weight gradient of w mapped
    back to w_bar (and to u, v when they require grad)."""

    @staticmethod
    def forward(ctx, x, w_bar, bias, u, v, layer, act, slope):
        _check(x)
        x = x.contiguous()
        st = _st()
        dev = x.device
        n, h, w, cx = x.shape
        tr = 1 if layer.kind == "convT" else 0
        T = layer.k * layer.k
        d0, d1 = (layer.cout, layer.cin) if not tr else (layer.cin, layer.cout)
        height, width = d0, T * d1
        t_saved = torch.empty(height, dtype=torch.float32, device=dev)
        sigma = torch.empty(1, dtype=torch.float32, device=dev)
        scratch = torch.empty(2, dtype=torch.float32, device=dev)
        call("gcc_spectral_norm_fwd", w_bar.data_ptr(), u.data_ptr(), v.data_ptr(), height, width, t_saved.data_ptr(),
             sigma.data_ptr(), scratch.data_ptr(), st)
        direct = torch.zeros(d0, T, rp8(d1), dtype=torch.bfloat16, device=dev)
        transposed = torch.zeros(d1, T, rp8(d0), dtype=torch.bfloat16, device=dev)
        call("gcc_pack_weight_scaled_bf16", w_bar.data_ptr(), sigma.data_ptr(), direct.data_ptr(), transposed.data_ptr(),
             d0, T, d1, direct.shape[2], transposed.shape[2], st)
        oh, ow = conv_out_hw(h, w, layer.k, layer.stride, layer.pad, tr, 0)
        cop = rp8(layer.cout)
        y = torch.empty(n, oh, ow, cop, dtype=torch.bfloat16, device=dev)
        wp = direct if not tr else transposed          # [cout][T][cin_p]
        epi = {ACT_NONE: 0, ACT_LRELU: 1, ACT_TANH: 2}[act]
        wsp, wse, _keep = ops._splitk_ws(n, oh, ow, layer.cout, dev)
        call("gcc_conv_gemm_bf16", x.data_ptr(), n, h, w, cx, wp.data_ptr(), layer.cout, T, wp.shape[2],
             None if bias is None else bias.data_ptr(), y.data_ptr(), oh, ow, cop, 0, tr, layer.k, layer.k, layer.stride,
             layer.pad, epi, slope, 0, wsp, wse, None, 0, st)
        ctx.layer, ctx.act, ctx.slope = layer, act, slope
        ctx.save_for_backward(x, y if act != ACT_NONE else None, direct, transposed, sigma, t_saved, w_bar, bias, u, v)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        x, y, direct, transposed, sigma, t_saved, w_bar, bias, u, v = ctx.saved_tensors
        dy = dy.contiguous()
        st = _st()
        dev = dy.device
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
        d0, d1 = (layer.cout, layer.cin) if not tr else (layer.cin, layer.cout)
        dx = None
        if ctx.needs_input_grad[0]:
            wp = transposed if not tr else direct        # [cin][T][cout_p]
            dx = torch.empty(n, h, w, rp8(layer.cin), dtype=torch.bfloat16, device=dev)
            wsp, wse, _keep = ops._splitk_ws(n, h, w, layer.cin, dev)
            call("gcc_conv_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, wp.data_ptr(), layer.cin, T, wp.shape[2], None,
                 dx.data_ptr(), h, w, dx.shape[3], 0, 0 if tr else 1, layer.k, layer.k, layer.stride, layer.pad, 0, 0.0, 0,
                 wsp, wse, None, 0, st)
        if ctx.needs_input_grad[1]:
            dweff = torch.empty(d0 * T * d1, dtype=torch.float32, device=dev)
            if not tr:
                call("gcc_wgrad_gemm_bf16", dpre.data_ptr(), n, oh, ow, cop, x.data_ptr(), h, w, cx, dweff.data_ptr(),
                     layer.cout, layer.cin, layer.k, layer.k, layer.stride, layer.pad, 0, 0, 1.0, st)
            else:
                call("gcc_wgrad_gemm_bf16", x.data_ptr(), n, h, w, cx, dpre.data_ptr(), oh, ow, cop, dweff.data_ptr(),
                     layer.cin, layer.cout, layer.k, layer.k, layer.stride, layer.pad, 0, 0, 1.0, st)
            g = lambda p: p._gcc_arena.flat_grad[p._gcc_name].data_ptr()
            uv = ctx.needs_input_grad[3] and ctx.needs_input_grad[4]
            scratch = torch.empty(1, dtype=torch.float32, device=dev)
            call("gcc_spectral_norm_bwd", dweff.data_ptr(), w_bar.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(),
                 t_saved.data_ptr(), d0, T * d1, g(w_bar), g(u) if uv else None, g(v) if uv else None, scratch.data_ptr(), st)
        if bias is not None and ctx.needs_input_grad[2]:
            gb = bias._gcc_arena.flat_grad[bias._gcc_name]
            call("gcc_bias_grad_bf16", dpre.data_ptr(), n * oh * ow, cop, 0, layer.cout, gb.data_ptr(), 1, st)
        return dx, None, None, None, None, None, None, None


class AttnFn(torch.autograd.Function):
    """softmax(q^T k) applied to v over the L = H * W positions of each sample (SAGAN.py:96-104)."""

    @staticmethod
    def forward(ctx, q, k, v, d, c):
        q, k, v = _check(q).contiguous(), k.contiguous(), v.contiguous()
        n, h, w, cp = v.shape
        L = h * w
        probs = torch.empty(n, L, L, dtype=torch.bfloat16, device=v.device)
        out = torch.empty_like(v)
        nb = _lib.lib().gcc_attn_workspace_bytes(n, L, q.shape[3], cp, 0)
        ws = torch.empty(nb, dtype=torch.uint8, device=v.device)
        call("gcc_attn_fwd_bf16", q.data_ptr(), k.data_ptr(), v.data_ptr(), n, L, d, q.shape[3], c, cp, probs.data_ptr(),
             out.data_ptr(), ws.data_ptr(), nb, _st())
        ctx.args = (n, L, d, c)
        ctx.save_for_backward(q, k, v, probs)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, probs = ctx.saved_tensors
        n, L, d, c = ctx.args
        dout = dout.contiguous()
        de = torch.empty_like(probs)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        nb = _lib.lib().gcc_attn_workspace_bytes(n, L, q.shape[3], v.shape[3], 1)
        ws = torch.empty(nb, dtype=torch.uint8, device=v.device)
        call("gcc_attn_bwd_bf16", q.data_ptr(), k.data_ptr(), v.data_ptr(), probs.data_ptr(), dout.data_ptr(), n, L, d,
             q.shape[3], c, v.shape[3], de.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), nb, _st())
        return dq, dk, dv, None, None


class ScaleAddFn(torch.autograd.Function):
    """gamma * a + x with a learnable scalar gamma (SAGAN.py:106)."""

    @staticmethod
    def forward(ctx, a, x, gamma):
        a, x = _check(a).contiguous(), x.contiguous()
        y = torch.empty_like(x)
        call("gcc_scale_add_bf16", a.data_ptr(), x.data_ptr(), gamma.data_ptr(), y.data_ptr(), x.numel(), _st())
        ctx.save_for_backward(a, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        a, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        da = torch.empty_like(dy)
        dg = gamma._gcc_arena.flat_grad[gamma._gcc_name].data_ptr() if ctx.needs_input_grad[2] else None
        call("gcc_scale_add_bwd_bf16", dy.data_ptr(), a.data_ptr(), gamma.data_ptr(), da.data_ptr(), dg, dy.numel(), _st())
        return da, dy, None


# -------------------------------------------------------------------------------------------------------- layers
class SNConvLayer:
    """SpectralNorm(conv): parameters ``<name>.module.{bias, weight_u, weight_v, weight_bar}`` (SAGAN.py:53-66)."""

    def __init__(self, arena, name, kind, cin, cout, k, stride, pad):
        self.arena, self.tname, self.kind = arena, name, kind
        self.cin, self.cout, self.k, self.stride, self.pad = cin, cout, k, stride, pad
        shape = (cout, cin, k, k) if kind == "conv" else (cin, cout, k, k)
        m = name + ".module"
        arena.add(m + ".bias", (cout,), "vec")
        arena.add(m + ".weight_u", (shape[0],), "vec")
        arena.add(m + ".weight_v", (shape[1] * k * k,), "vec")
        arena.add(m + ".weight_bar", shape, kind + "_nopack")

    def bind(self):
        P = self.arena.params
        m = self.tname + ".module"
        self.bias, self.u, self.v, self.weight_bar = P[m + ".bias"], P[m + ".weight_u"], P[m + ".weight_v"], P[m + ".weight_bar"]
        self.u.requires_grad = False       # Parameter(..., requires_grad=False) until set_requires_grad touches them
        self.v.requires_grad = False

    def named(self):
        m = self.tname + ".module"
        return [(m + ".bias", self.bias), (m + ".weight_u", self.u), (m + ".weight_v", self.v), (m + ".weight_bar", self.weight_bar)]

    def __call__(self, x, act=ACT_NONE, slope=0.2):
        return SNConvFn.apply(x, self.weight_bar, self.bias, self.u, self.v, self, act, slope)


class SelfAttn:
    """Self_Attn (SAGAN.py:73-107): three 1x1 convs, softmax attention over positions, ``gamma * out + x``."""

    def __init__(self, conv_arena, gamma_arena, name, c):
        self.tname, self.c, self.d = name, c, c // 8
        self.gamma_arena = gamma_arena
        gamma_arena.add(name + ".gamma", (1,), "vec")
        self.q = ConvLayer(conv_arena, name + ".query_conv", "conv", c, c // 8, 1, 1, 0, bias=True)
        self.k = ConvLayer(conv_arena, name + ".key_conv", "conv", c, c // 8, 1, 1, 0, bias=True)
        self.v = ConvLayer(conv_arena, name + ".value_conv", "conv", c, c, 1, 1, 0, bias=True)

    def bind(self):
        self.gamma = self.gamma_arena.params[self.tname + ".gamma"]

    def __call__(self, x):
        out = AttnFn.apply(self.q(x), self.k(x), self.v(x), self.d, self.c)
        return ScaleAddFn.apply(out, x, self.gamma)


class _SANet(_SRNet):
    """Registration of spectral-norm / attention parameters under the reference's names; weight_v is stored in the
    arena's channels-last column order (kh, kw, c) and converted to the reference's (c, kh, kw) at the state-dict
    boundary."""

    def _register_sa(self, plain_layers, sn_layers, attns):
        self._register(plain_layers)
        extra = _Tree()
        self._sn = sn_layers
        for l in sn_layers:
            l.bind()
            for name, p in l.named():
                extra.put(name, p)
        for a in attns:
            a.bind()
            extra.put(a.tname + ".gamma", a.gamma)
        for k, m in extra._modules.items():
            if k in self._modules:
                _merge(self._modules[k], m)
            else:
                self.add_module(k, m)

    def _v_perm(self, l, v, to_ref):
        T = l.k * l.k
        d1 = v.numel() // T
        return v.reshape(T, d1).t().reshape(-1) if to_ref else v.reshape(d1, T).t().reshape(-1)

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        for l in self._sn:
            key = l.tname + ".module.weight_v"
            sd[key] = self._v_perm(l, sd[key], True).contiguous()
        return sd

    def load_state_dict(self, sd, strict=True):
        sd = dict(sd)
        for l in self._sn:
            key = l.tname + ".module.weight_v"
            if key in sd:
                sd[key] = self._v_perm(l, sd[key], False).contiguous()
        return super().load_state_dict(sd, strict)


class Generator(_SANet):
    """SAGAN.py:109-171 (image size 64): 4 x [SpectralNorm(ConvT k4) + BN + ReLU] with attention after levels 3 and 4,
    ConvT k4 s2 + Tanh.  Taps: output of l2 and of attn2."""

    def __init__(self, ngf=64, image_size=64, z_dim=128, filter_cfgs=None, arena=None, sn_arena=None, device="cuda"):
        super().__init__()
        assert image_size == 64, "the reference's forward only supports 64 x 64 (SAGAN.py:140-148,167)"
        self.arena = arena if arena is not None else ParamArena(device, betas=(0.0, 0.9))
        self.sn_arena = sn_arena if sn_arena is not None else self.arena
        self._arenas = [self.arena] if self.sn_arena is self.arena else [self.arena, self.sn_arena]
        c = [ngf * 8, ngf * 4, ngf * 2, ngf] if filter_cfgs is None else list(filter_cfgs)
        self.c = c
        self.sn, self.bn = [], []
        cin = z_dim
        for i, co in enumerate(c):
            self.sn.append(SNConvLayer(self.sn_arena, "l%d.0" % (i + 1), "convT", cin, co, 4, 1 if i == 0 else 2, 0 if i == 0 else 1))
            self.bn.append(NormLayer(self.arena, "l%d.1" % (i + 1), co, "bn", device))
            cin = co
        self.last_conv = ConvLayer(self.arena, "last.0", "convT", c[3], 3, 4, 2, 1, bias=True)
        self.at1 = SelfAttn(self.sn_arena, self.arena, "attn1", c[2])
        self.at2 = SelfAttn(self.sn_arena, self.arena, "attn2", c[3])
        self.taps = []
        if arena is None:
            self.finalize()

    def finalize(self):
        for a in self._arenas:
            if not a.finalized:
                a.finalize()
        plain = list(self.bn) + [self.last_conv]
        for at in (self.at1, self.at2):
            plain += [at.q, at.k, at.v]
        self._register_sa(plain, self.sn, [self.at1, self.at2])

    def forward(self, z):
        out = z
        self.taps = []
        for i in range(4):
            out = self.bn[i](self.sn[i](out), ACT_RELU)
            if i == 1:
                self.taps.append((out, self.c[1]))
            if i == 2:
                out = self.at1(out)
        out = self.at2(out)
        self.taps.append((out, self.c[3]))
        return self.last_conv(out, ACT_TANH)


class Discriminator(_SANet):
    """SAGAN.py:173-222 (``gated=False``) / MaskDiscriminator :224-277: 4 x [SpectralNorm(Conv k4 s2) -> [gate] ->
    LeakyReLU(0.1)], attention after levels 3 and 4, Conv k4 -> one logit per image.  Taps: l2 output, attn2 output."""

    def __init__(self, ndf=64, image_size=64, threshold=0.5, gated=False, arena=None, sn_arena=None, gate_arena=None,
                 device="cuda"):
        super().__init__()
        assert image_size == 64
        self.arena = arena if arena is not None else ParamArena(device, betas=(0.0, 0.9))
        self.sn_arena = sn_arena if sn_arena is not None else self.arena
        if gated and gate_arena is None:
            gate_arena = ParamArena(device)
        self.gate_arena = gate_arena
        self._arenas = [a for a in (self.arena, self.sn_arena if self.sn_arena is not self.arena else None, gate_arena)
                        if a is not None]
        self.gated = gated
        self.sn, self.gates = [], []
        cin = 3
        for i in range(4):
            co = ndf * (2 ** i)
            self.sn.append(SNConvLayer(self.sn_arena, "l%d.0" % (i + 1), "conv", cin, co, 4, 2, 1))
            if gated:
                g = NormLayer(self.arena, "l%d.gate" % (i + 1), co, "id", device, gate_arena=gate_arena,
                              gate_name="l%d.1" % (i + 1), thr=threshold, slope=LRELU)
                g.gate_after = False
                self.gates.append(g)
            cin = co
        self.c = [ndf, ndf * 2, ndf * 4, ndf * 8]
        self.last_conv = ConvLayer(self.arena, "last.0", "conv", cin, 1, 4, 1, 0, bias=True)
        self.at1 = SelfAttn(self.sn_arena, self.arena, "attn1", ndf * 4)
        self.at2 = SelfAttn(self.sn_arena, self.arena, "attn2", ndf * 8)
        self.taps = []
        if arena is None:
            self.finalize()

    def finalize(self):
        for a in self._arenas:
            if not a.finalized:
                a.finalize()
        plain = list(self.gates) + [self.last_conv]
        for at in (self.at1, self.at2):
            plain += [at.q, at.k, at.v]
        self._register_sa(plain, self.sn, [self.at1, self.at2])

    def forward(self, x):
        out = x
        self.taps = []
        for i in range(4):
            if self.gated:
                out = self.gates[i](self.sn[i](out), ACT_LRELU)
            else:
                out = self.sn[i](out, ACT_LRELU, LRELU)
            if i == 1:
                self.taps.append((out, self.c[1]))
            if i == 2:
                out = self.at1(out)
        out = self.at2(out)
        self.taps.append((out, self.c[3]))
        return self.last_conv(out)                      # [N, 1, 1, 8]: one logit per image (``.squeeze()`` upstream)

    def get_current_masks(self):
        return [((torch.sign(g.alpha.detach() - g.thr) + 1) / 2) for g in self.gates]


class MaskDiscriminator(Discriminator):
    def __init__(self, ndf=64, image_size=64, threshold=0.5, arena=None, sn_arena=None, gate_arena=None, device="cuda"):
        super().__init__(ndf, image_size, threshold, True, arena, sn_arena, gate_arena, device)


class _MultiArenaOptimizer:
    """torch.optim-like facade over several arenas of one optimizer; ``times`` = how often the reference's parameter
    list holds the arena's tensors (each occurrence is a full Adam update with the same gradient)."""

    def __init__(self, arenas, lr, betas):
        self.arenas = arenas
        for a, _ in arenas:
            a.lr, a.betas = lr, betas
            if a.finalized:
                a._write_hyper()
        self.param_groups = [{"lr": lr, "initial_lr": lr, "betas": betas}]

    def zero_grad(self):
        for a, _ in self.arenas:
            a.zero_grad()

    def arenas_of(self):
        return [a for a, _ in self.arenas]

    def step(self):
        for a, times in self.arenas:
            if not capturing():
                a.set_lr(self.param_groups[0]["lr"])
            _allreduce_grads(a)
            for _ in range(times):
                a.step()


# ------------------------------------------------------------------------------------------------------- model
class SAGANModel(GccModelMixin, torch.nn.Module):
    """models/SAGAN.py:279-764 on the B200 kernels."""

    def __init__(self, opt, filter_cfgs=None, channel_cfgs=None):
        super().__init__()
        self.opt = opt
        if len(opt.gpu_ids) == 0:
            raise RuntimeError("gcc_b200.SAGANModel needs a CUDA device (gpu_ids): there is no CPU path")
        self.device = dev = torch.device("cuda:%d" % opt.gpu_ids[0])
        call("gcc_check_device")
        self.filter_cfgs, self.channel_cfgs = filter_cfgs, channel_cfgs
        self.loss_names = ["G_GAN", "D_real", "D_fake"]
        self.visual_names = ["fake_img", "real_img"]
        self.current_D_arch_diff_loss = 0.0
        self.teacher_model = None
        self._base_init()
        self.generator_extract_layers = ["l2", "attn2"]
        self.discriminator_extract_layers = ["l2", "attn2"]
        self.distill = bool(opt.online_distillation or getattr(opt, "normal_distillation", False))
        betas = (0.0, 0.9)
        dup_g = 2 if self.distill else 1                  # SAGAN.py:318-326
        dup_d = 2 if opt.darts_discriminator else 1       # SAGAN.py:339-344

        self.arena_G, self.arena_Gs = ParamArena(dev, betas=betas), ParamArena(dev, betas=betas)
        self.transform_convs = []
        if self.distill:
            t_ch = [opt.teacher_ngf * 4, opt.teacher_ngf]
            s_ch = [opt.ngf * 4, opt.ngf] if filter_cfgs is None else [filter_cfgs[1], filter_cfgs[3]]
            for i in range(2):
                self.transform_convs.append(ConvLayer(self.arena_G, "transform.%d" % i, "conv", s_ch[i], t_ch[i], 1, 1, 0))
        self.netG = Generator(ngf=opt.ngf, image_size=opt.crop_size, z_dim=opt.z_dim, filter_cfgs=filter_cfgs,
                              arena=self.arena_G, sn_arena=self.arena_Gs, device=dev)
        self.netG.finalize()
        for t in self.transform_convs:
            t.bind()
        self.optimizer_G = _MultiArenaOptimizer([(self.arena_G, 1), (self.arena_Gs, dup_g)], opt.lr, betas)

        self.arena_D, self.arena_Ds = ParamArena(dev, betas=betas), ParamArena(dev, betas=betas)
        if opt.darts_discriminator:
            self.loss_names += ["D_arch_diff", "D_arch", "teacher_D_arch_diff"]
            self.arena_A = ParamArena(dev)
            self.netD = MaskDiscriminator(ndf=opt.ndf, arena=self.arena_D, sn_arena=self.arena_Ds, gate_arena=self.arena_A,
                                          device=dev)
            self.netD.finalize()
            self.optimizer_arch = _MultiArenaOptimizer([(self.arena_A, 1)], opt.arch_lr, (0.9, 0.999))
        else:
            self.arena_A = None
            self.netD = Discriminator(ndf=opt.ndf, arena=self.arena_D, sn_arena=self.arena_Ds, device=dev)
            self.netD.finalize()
        self.optimizer_D = _MultiArenaOptimizer([(self.arena_D, 1), (self.arena_Ds, dup_d)], opt.lr * 4, betas)
        self.optimizers = []
        self.schedulers = []
        if opt.darts_discriminator and getattr(opt, "arch_lr_step", False):
            # StepLR(step_size=40, gamma=0.1) over optimizer_arch only (SAGAN.py:348-353, stepped at :548-549)
            arch_opt = copy.deepcopy(opt)
            arch_opt.lr_policy, arch_opt.lr_decay_iters = "step", 40
            self.arch_scheduler = get_scheduler(self.optimizer_arch, arch_opt)
            self.schedulers.append(self.arch_scheduler)
        self.init_net()
        self.gan_mode = GAN_MODES.get(opt.gan_mode)
        if self.gan_mode is None:
            raise NotImplementedError("gan mode %s not implemented" % opt.gan_mode)
        self.total_generator_features, self.total_discriminator_features = {}, {}
        self.broadcast_parameters()

    def _gcc_arenas(self):
        return {k: a for k, a in (("G", self.arena_G), ("Gs", self.arena_Gs), ("D", self.arena_D), ("Ds", self.arena_Ds),
                                  ("A", self.arena_A)) if a is not None}

    def _gcc_optimizers(self):
        d = {"G": self.optimizer_G, "D": self.optimizer_D}
        if self.arena_A is not None:
            d["arch"] = self.optimizer_arch
        return d

    def _gcc_nets(self):
        return {"netG": self.netG, "netD": self.netD}

    def _all_arenas(self):
        return [a for a in (self.arena_G, self.arena_Gs, self.arena_D, self.arena_Ds, self.arena_A) if a is not None]

    def init_net(self):
        """util.init_weights (utils/util.py:261-286) skips the spectral-normed convs (no ``.weight`` attribute yet):
        they keep nn.Conv2d's default kaiming-uniform; u, v ~ normalised N(0, 1); attention gamma = 0; alpha = 1."""
        with torch.no_grad():
            for arena in (self.arena_G, self.arena_Gs, self.arena_D, self.arena_Ds):
                for (name, shape, kind) in arena.specs:
                    p = arena.params[name]
                    if name.startswith("transform.") or name.endswith("weight_bar"):
                        fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
                        bound = 1.0 / math.sqrt(fan_in)
                        p.copy_(torch.empty(shape, device=self.device).uniform_(-bound, bound))
                    elif name.endswith("weight_u") or name.endswith("weight_v"):
                        t = torch.empty(shape, device=self.device).normal_(0.0, 1.0)
                        p.copy_(t / (t.norm() + 1e-12))
                    elif name.endswith("gamma"):
                        p.zero_()
                    elif len(shape) == 4:
                        p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 0.02))
                    elif name.endswith(".weight"):
                        p.copy_(torch.empty(shape, device=self.device).normal_(1.0, 0.02))
                    elif name.endswith(".bias"):
                        wname = name[:-4] + "weight"
                        is_bn = wname in arena.params and len(arena.params[wname].shape) == 1
                        p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 1.0)) if is_bn else p.zero_()
                arena.mark_dirty()

    def sync_weights(self):
        for a in self._all_arenas():
            a.mark_dirty()

    # ------------------------------------------------------------------ inputs / forward
    def set_input(self, input):
        self.input = input
        self.z = input["z"].to(self.device, non_blocking=True)
        self._real_nchw = input["real_img"].to(self.device, non_blocking=True)
        self.image_paths = [input.get("img_path"), input.get("img_path")]
        self.z_nhwc = ops.to_nhwc(self.z.reshape(self.z.shape[0], self.z.shape[1], 1, 1))
        self.real_img_nhwc = ops.to_nhwc(self._real_nchw)

    def _adopt_input(self, other):
        """Teacher side of ``T.set_input(self.input)`` (SAGAN.py:500,519): share the converted batch."""
        if other.device != self.device:
            return self.set_input(other.input)
        self.input, self.image_paths = other.input, other.image_paths
        self.z, self._real_nchw = other.z, other._real_nchw
        self.z_nhwc, self.real_img_nhwc = other.z_nhwc, other.real_img_nhwc

    @property
    def real_img(self):
        return self._real_nchw

    @property
    def fake_img(self):
        return ops.to_nchw(self.fake_img_nhwc.detach(), 3)

    @property
    def Tfake_img(self):
        return ops.to_nchw(self.Tfake_img_nhwc, 3)

    def forward(self):
        self.fake_img_nhwc = self.netG(self.z_nhwc)
        self.g_taps = list(self.netG.taps)

    def _D(self, net, x):
        pred = net(x)
        return pred, list(net.taps)

    def _gan(self, pred, kind):
        return ops.GanLossFn.apply(pred, 1, self.gan_mode, kind)

    # ------------------------------------------------------------------ losses / steps
    def backward_D(self):
        pred_real, _ = self._D(self.netD, self.real_img_nhwc)
        self.loss_D_real = self._gan(pred_real, 0)
        pred_fake, self.d_taps = self._D(self.netD, self.fake_img_nhwc.detach())
        self.loss_D_fake = self._gan(pred_fake, 1)
        self.loss_D = self.loss_D_fake + self.loss_D_real
        self.loss_D.backward()

    def get_D_arch_diff(self, isTeacher=False):
        if isTeacher:
            self.set_requires_grad(self.netD, False)
        pred_fake, _ = self._D(self.netD, self.fake_img_nhwc.detach())
        self.loss_D_arch_fake = self._gan(pred_fake, 1)
        self.loss_D_arch_fake_real = self._gan(pred_fake, 2)
        pred_real, _ = self._D(self.netD, self.real_img_nhwc)
        self.loss_D_arch_real = self._gan(pred_real, 0)
        diff = (self.loss_D_arch_fake_real - self.loss_D_arch_fake).abs()
        if isTeacher:
            self.current_D_arch_diff_loss = self._ema_update("D", diff)
        else:
            self.current_D_arch_diff_loss = diff
        return self.current_D_arch_diff_loss, torch.sign(self.loss_D_arch_fake_real - self.loss_D_arch_fake)

    def backward_D_arch(self):
        self.loss_teacher_D_arch_diff, _ = self.teacher_model.get_D_arch_diff(isTeacher=True)
        self.loss_D_arch_diff, _ = self.get_D_arch_diff(isTeacher=False)
        self.loss_D_arch = (self.loss_D_arch_diff - self.loss_teacher_D_arch_diff.detach()).abs()
        self.loss_D_arch = self.loss_D_arch + self.loss_D_arch_real + self.loss_D_arch_fake
        self.loss_D_arch.backward()

    def backward_G(self):
        o = self.opt
        pred_fake, self.d_taps = self._D(self.netD, self.fake_img_nhwc)
        self.loss_G_GAN = self._gan(pred_fake, 2)
        self.loss_G = self.loss_G_GAN
        if self.distill:
            T = self.teacher_model
            self.Tfake_img_nhwc = T.fake_img_nhwc.detach()
            feats = list(self.g_taps)
            _, t_d_taps = self._D(T.netD, self.fake_img_nhwc)        # teacher D (frozen here) on the student fake
            feats += t_d_taps
            self.loss_content = 0.0
            self.loss_gram = 0.0
            for i, (f, c) in enumerate(feats):
                if i < 2:
                    f = self.transform_convs[i](f)
                    c = self.transform_convs[i].cout
                tgt, gram_t = self.target_distillation_features[i], self.target_grams[i]
                self.loss_gram = self.loss_gram + ops.GramRmseFn.apply(f, gram_t, c)
                self.loss_content = self.loss_content + ops.DiffLossFn.apply(f, tgt, c, 1)
            self.loss_gram = o.lambda_gram * self.loss_gram
            self.loss_content = o.lambda_content * self.loss_content
            self.loss_L1 = o.lambda_L1 * ops.DiffLossFn.apply(self.fake_img_nhwc, self.Tfake_img_nhwc, 3, 0)
            self.loss_G = self.loss_G + self.loss_gram + self.loss_content + self.loss_L1
            # the reference accumulates in place into an alias of loss_G_GAN (SAGAN.py:463,483-485)
            self.loss_G_GAN = self.loss_G
        self.loss_G.backward()
        self.L1_sparsity()

    def L1_sparsity(self):
        o, st = self.opt, _st()
        for arena in (self.arena_G, self.arena_Gs):
            names = [s[0] for s in arena.specs]
            for (name, shape, kind) in arena.specs:
                if name.startswith("transform."):
                    continue
                off, n = arena.offsets[names.index(name)]
                if o.lambda_weight > 0.0 and len(shape) == 4 and not name.endswith("weight_bar"):
                    # (spectral-normed convs have no ``.weight`` parameter: the reference's loop reaches the attention
                    # convs and ``last`` only)
                    call("gcc_l1_sparsity_f32", arena.P[off:].data_ptr(), arena.G[off:].data_ptr(), n, o.lambda_weight, st)
                elif o.lambda_weight <= 0.0 and o.lambda_scale > 0.0 and len(shape) == 1 and name.endswith(".1.weight"):
                    call("gcc_l1_sparsity_f32", arena.P[off:].data_ptr(), arena.G[off:].data_ptr(), n, o.lambda_scale, st)

    def _release_graphs(self):
        for k, v in list(vars(self).items()):
            if isinstance(v, torch.Tensor) and v.grad_fn is not None:
                object.__setattr__(self, k, v.detach())
        for name in ("g_taps", "d_taps"):
            taps = getattr(self, name, None)
            if taps:
                setattr(self, name, [(f.detach(), c) for f, c in taps])
        self.netG.taps = []
        self.netD.taps = []

    def optimize_parameters(self):
        ops.zero_pool.reset()
        if self.opt.online_distillation:
            T = self.teacher_model
            T._adopt_input(self)
            T.optimize_parameters()
            feats = [f.detach() for f, _ in (T.g_taps + T.d_taps)]
            chans = [c for _, c in (T.g_taps + T.d_taps)]
            self.target_distillation_features = feats
            self.target_grams = [ops.gram_matrix(f, c) for f, c in zip(feats, chans)]
        self.forward()
        self.set_requires_grad(self.netD, True)
        self.set_netD_arch_grad(False)
        self.optimizer_D.zero_grad()
        self.backward_D()
        self.optimizer_D.step()
        self.set_requires_grad(self.netD, False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        self.optimizer_G.step()
        self._release_graphs()

    def optimizer_netD_arch(self):
        ops.zero_pool.reset()
        self.forward()
        self.teacher_model._adopt_input(self)
        self.teacher_model.forward()
        self.set_requires_grad(self.netD, True)
        self.set_netD_weight_grad(False)
        self.optimizer_arch.zero_grad()
        self.backward_D_arch()
        self.optimizer_arch.step()
        self._release_graphs()
        self.teacher_model._release_graphs()

    # ------------------------------------------------------------------ bookkeeping (reference surface)
    def print_sparse_info(self, logger):
        for i, mask in enumerate(self.netD.get_current_masks() if self.opt.darts_discriminator else []):
            logger.info("netD gate %d sparsity ratio: %.2f" % (i, float((mask == 0.0).sum()) / mask.numel()))

    def adaptive_ema_beta(self, epoch):
        self.opt.ema_beta = 1.0 - epoch / (self.opt.n_epochs + self.opt.n_epochs_decay)

    def update_learning_rate(self, epoch):
        for sch in self.schedulers:
            sch.step()
        self.adaptive_ema_beta(epoch)
        print("learning rate = %.7f" % self.optimizer_G.param_groups[0]["lr"])

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad

    def set_netD_weight_grad(self, requires_grad=False):
        """Conv2d / BatchNorm2d modules only (SAGAN.py:564-568): everything but the attention gammas and the gates."""
        for n, p in self.netD.named_parameters():
            if not (n.endswith("gamma") or n.endswith("alpha")):
                p.requires_grad = requires_grad

    def set_netD_arch_grad(self, requires_grad=False):
        if self.arena_A is not None:
            for p in self.arena_A.params.values():
                p.requires_grad = requires_grad

    def clipping_mask_alpha(self):
        if self.arena_A is not None:
            call("gcc_clamp_f32", self.arena_A.P.data_ptr(), self.arena_A.numel, 0.0, 1.0, _st())

    def model_train(self):
        self.netG.train()
        self.netD.train()

    def model_eval(self):
        self.netG.eval()
        self.netD.eval()

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names if isinstance(n, str) and hasattr(self, n))

    def get_current_losses(self):
        return OrderedDict((n, float(getattr(self, "loss_" + n))) for n in self.loss_names if isinstance(n, str))

    def init_distillation(self):
        self.total_generator_features, self.total_discriminator_features = {}, {}
        if self.distill:
            o = self.opt
            if o.lambda_content > 0.0:
                self.loss_names.append("content")
            if o.lambda_gram > 0.0:
                self.loss_names.append("gram")
            if o.lambda_L1 > 0.0:
                self.loss_names.append("L1")
            self.visual_names.append("Tfake_img")

    def get_distillation_features(self):
        return [ops.to_nchw(f.detach(), c) for f, c in (list(getattr(self, "g_taps", [])) + list(getattr(self, "d_taps", [])))]

    def gram(self, x):
        b, c, h, w = x.size()
        return ops.gram_matrix(ops.to_nhwc(x), c)

    def get_cfg(self):
        return self.filter_cfgs, self.channel_cfgs

    @staticmethod
    def _pop_ops(sd):
        for k in list(sd.keys()):
            if k.endswith("total_ops") or k.endswith("total_params"):
                sd.pop(k)
        return sd

    def save_models(self, epoch, save_dir, fid=None, isbest=False, direction="AtoB"):
        os.makedirs(save_dir, exist_ok=True)
        ckpt = {"G": self._pop_ops(self.netG.state_dict()), "D": self._pop_ops(self.netD.state_dict()), "epoch": epoch,
                "cfg": (self.filter_cfgs, self.channel_cfgs), "fid": fid}
        path = os.path.join(save_dir, "model_best_%s.pth" % direction if isbest else "model_%d.pth" % epoch)
        torch.save(self._ckpt_add_resume(ckpt), path)

    def load_models(self, load_path, load_discriminator=True, resume=None):
        """`resume`: None = continue training exactly (optimizer moments, counters, EMA, teacher) when the file carries
        the `gcc_b200` entry and the discriminator is loaded too; False = weights only, as the reference."""
        ckpt = torch.load(load_path, map_location=self.device, weights_only=False)
        self.netG.load_state_dict(self._pop_ops(ckpt["G"]))
        if load_discriminator:
            self.netD.load_state_dict(self._pop_ops(ckpt["D"]))
        self._ckpt_load_resume(ckpt, load_discriminator, resume)
        print("loading the model from %s" % load_path)
        return ckpt["fid"], float("inf")

    # ------------------------------------------------------------------ pruning front-end (SAGAN.py:694-764)
    def prune(self, threshold, lottery_path=None):
        if self.opt.scale_prune:
            return self.scale_prune(threshold)
        raise NotImplementedError("only scale and norm pruning are supported!!!")

    def max_min_bn_scale(self):
        mx, mn = float("inf"), float("inf")
        for bn in self.netG.bn:
            mx = min(torch.max(bn.gamma.data), mx)
            mn = min(torch.min(bn.gamma.data), mn)
        return mx, mn

    def max_min_conv_norm(self):
        pass

    def scale_prune(self, threshold):
        cfg = [int((bn.gamma.data > threshold).sum()) for bn in self.netG.bn]
        print(cfg)
        return SAGANModel(self.opt, filter_cfgs=cfg)


def build_sa_teacher(model, opt):
    """train.py:92-105."""
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    topt.generator_only = False
    teacher = SAGANModel(topt)
    teacher.model_train()
    model.teacher_model = teacher
    model.init_distillation()
    teacher.init_distillation()
    return teacher
