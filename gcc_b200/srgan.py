"""SRGAN on the B200 kernels: SRResNet generator, (gated) discriminator, truncated VGG19 and the ``SRGAN`` model
wrapper (reference: /root/reference/models/SRGAN.py:15-842, models/GANLoss.py:95-144,
data/sr_dataset.py:15-64).  Class / method / attribute / state-dict names follow the reference so that ``train.py``
drives this class unchanged:

    model.set_input({'lr','hr',...}); model.optimize_parameters()
    model.set_input(val);             model.clipping_mask_alpha(); model.optimizer_netD_arch()

Activations are NHWC bf16 on the device; the convolutions, norms, gates and losses are the kernels of the pix2pix
path, plus the SRGAN-only kernels of csrc/srgan.cu (PReLU, PixelShuffle, MaxPool, pooled linear head, channel affine).
"""
import copy
import math
import os
from collections import OrderedDict

import torch

from . import ops
from .arena import ParamArena, rp8
from .nets import ConvLayer, NormLayer, _Net, _Tree
from .ops import ACT_LRELU, ACT_NONE, ACT_TANH, GAN_MODES, _check, _st, call
from .base import GccModelMixin
from .pix2pix import _ArenaOptimizer, get_scheduler

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512]


# ------------------------------------------------------------------------------------------- autograd functions
class PReLUFn(torch.autograd.Function):
    """nn.PReLU() with one learnable slope (SRGAN.py:49-50,89); the slope gradient accumulates into its arena."""

    @staticmethod
    def forward(ctx, x, slope):
        x = _check(x).contiguous()
        y = torch.empty_like(x)
        call("gcc_prelu_fwd_bf16", x.data_ptr(), y.data_ptr(), x.numel(), slope.data_ptr(), _st())
        ctx.save_for_backward(x, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, slope = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy) if ctx.needs_input_grad[0] else None
        ds = slope._gcc_arena.flat_grad[slope._gcc_name].data_ptr() if ctx.needs_input_grad[1] else None
        call("gcc_prelu_bwd_bf16", x.data_ptr(), dy.data_ptr(), None if dx is None else dx.data_ptr(), dy.numel(),
             slope.data_ptr(), ds, _st())
        return dx, None


class PixelShuffleFn(torch.autograd.Function):
    """nn.PixelShuffle(2) (SRGAN.py:88): [N,H,W,4C] -> [N,2H,2W,C]."""

    @staticmethod
    def forward(ctx, x, c):
        x = _check(x).contiguous()
        n, h, w, cin_p = x.shape
        y = torch.empty(n, 2 * h, 2 * w, rp8(c), dtype=torch.bfloat16, device=x.device)
        call("gcc_pixel_shuffle2_bf16", x.data_ptr(), y.data_ptr(), n, h, w, c, cin_p, y.shape[3], 0, _st())
        ctx.args = (n, h, w, c, cin_p)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, h, w, c, cin_p = ctx.args
        dy = dy.contiguous()
        dx = torch.empty(n, h, w, cin_p, dtype=torch.bfloat16, device=dy.device)
        call("gcc_pixel_shuffle2_bf16", dy.data_ptr(), dx.data_ptr(), n, h, w, c, cin_p, dy.shape[3], 1, _st())
        return dx, None


class MaxPoolFn(torch.autograd.Function):
    """nn.MaxPool2d(2, 2) of torchvision's VGG19."""

    @staticmethod
    def forward(ctx, x):
        x = _check(x).contiguous()
        n, h, w, cp = x.shape
        y = torch.empty(n, h // 2, w // 2, cp, dtype=torch.bfloat16, device=x.device)
        call("gcc_maxpool2_fwd_bf16", x.data_ptr(), y.data_ptr(), n, h, w, cp, _st())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, h, w, cp = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        call("gcc_maxpool2_bwd_bf16", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), n, h, w, cp, _st())
        return dx


class ChannelAffineFn(torch.autograd.Function):
    """y = x * scale[c] + shift[c] on an 8-channel image (convert_image '[-1, 1]' -> 'imagenet-norm')."""

    @staticmethod
    def forward(ctx, x, scale, shift):
        x = _check(x).contiguous()
        y = torch.empty_like(x)
        call("gcc_channel_affine8_bf16", x.data_ptr(), y.data_ptr(), x.numel() // 8, 3, scale.data_ptr(), shift.data_ptr(),
             _st())
        ctx.save_for_backward(scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        call("gcc_channel_affine8_bf16", dy.data_ptr(), dx.data_ptr(), dy.numel() // 8, 3, scale.data_ptr(), None, _st())
        return dx, None, None


class PoolLinearFn(torch.autograd.Function):
    """AdaptiveAvgPool2d((1,1)) + Linear(C, 1) (SRGAN.py:231-245): NHWC bf16 -> logits bf16 [N, 1, 1, 8]."""

    @staticmethod
    def forward(ctx, x, weight, bias, c):
        x = _check(x).contiguous()
        n, h, w, cp = x.shape
        st = _st()
        sums = torch.empty(n * 2 * cp, dtype=torch.float32, device=x.device)
        call("gcc_norm_stats_bf16", x.data_ptr(), n, h * w, cp, 1, sums.data_ptr(), st)
        out = torch.empty(n, 1, 1, 8, dtype=torch.bfloat16, device=x.device)
        call("gcc_pool_linear_fwd", sums.data_ptr(), n, h * w, cp, c, weight.data_ptr(), bias.data_ptr(), out.data_ptr(), st)
        ctx.args = (n, h, w, cp, c)
        ctx.save_for_backward(sums, weight, bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        sums, weight, bias = ctx.saved_tensors
        n, h, w, cp, c = ctx.args
        dout = dout.contiguous()
        dx = torch.empty(n, h, w, cp, dtype=torch.bfloat16, device=dout.device) if ctx.needs_input_grad[0] else None
        dw = weight._gcc_arena.flat_grad[weight._gcc_name].data_ptr() if ctx.needs_input_grad[1] else None
        db = bias._gcc_arena.flat_grad[bias._gcc_name].data_ptr() if ctx.needs_input_grad[2] else None
        call("gcc_pool_linear_bwd", dout.data_ptr(), sums.data_ptr(), weight.data_ptr(), n, h * w, cp, c,
             None if dx is None else dx.data_ptr(), dw, db, _st())
        return dx, None, None, None


def convert_to_imagenet(x_nhwc, consts):
    """convert_image(x, '[-1, 1]', 'imagenet-norm') on an NHWC bf16 image: ((x + 1) / 2 - mean) / std."""
    return ChannelAffineFn.apply(x_nhwc, consts[0], consts[1])


def _imagenet_consts(device):
    scale = torch.tensor([0.5 / s for s in IMAGENET_STD] + [0.0] * 5, dtype=torch.float32, device=device)
    shift = torch.tensor([(0.5 - m) / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)] + [0.0] * 5, dtype=torch.float32,
                         device=device)
    return scale, shift


# -------------------------------------------------------------------------------------------------------- nets
class _Scalar:
    """A 1-element learnable parameter (nn.PReLU weight) living in an arena."""

    def __init__(self, arena, name):
        self.arena, self.tname, self.name = arena, name, name
        arena.add(name + ".weight", (1,), "vec")

    def bind(self):
        self.weight = self.arena.params[self.name + ".weight"]


class _SRNet(_Net):
    def _register_sr(self, layers):
        """_Net._register plus PReLU scalars and the linear head."""
        tree_layers = [l for l in layers if not isinstance(l, (_Scalar, _Linear))]
        self._register(tree_layers)
        extra = _Tree()
        for l in layers:
            if isinstance(l, _Scalar):
                l.bind()
                extra.put(l.tname + ".weight", l.weight)
            elif isinstance(l, _Linear):
                l.bind()
                extra.put(l.tname + ".weight", l.weight)
                extra.put(l.tname + ".bias", l.bias)
        # merge: same top-level names may exist in both trees
        for k, m in extra._modules.items():
            if k in self._modules:
                _merge(self._modules[k], m)
            else:
                self.add_module(k, m)


def _merge(dst, src):
    for k, p in src._parameters.items():
        dst.register_parameter(k, p)
    for k, m in src._modules.items():
        if k in dst._modules:
            _merge(dst._modules[k], m)
        else:
            dst.add_module(k, m)


class _Linear:
    def __init__(self, arena, name, cin):
        self.arena, self.tname, self.cin = arena, name, cin
        arena.add(name + ".weight", (1, cin), "vec")
        arena.add(name + ".bias", (1,), "vec")

    def bind(self):
        self.weight = self.arena.params[self.tname + ".weight"]
        self.bias = self.arena.params[self.tname + ".bias"]


class Generator(_SRNet):
    """SRResNet (SRGAN.py:139-197): Conv9+PReLU, n_blocks x [Conv3+BN+PReLU, Conv3+BN] + skip, Conv3+BN + skip,
    2 x [Conv3 (-> 4C) + PixelShuffle(2) + PReLU], Conv9 + Tanh.  Taps: outputs of residual blocks 3/7/11/15."""

    def __init__(self, large_kernel_size=9, small_kernel_size=3, n_channels=64, n_blocks=16, scaling_factor=4,
                 filter_cfgs=None, arena=None, prelu_arena=None, device="cuda"):
        super().__init__()
        assert int(scaling_factor) == 4 and large_kernel_size == 9 and small_kernel_size == 3
        self.arena = arena if arena is not None else ParamArena(device)
        self.prelu_arena = prelu_arena if prelu_arena is not None else self.arena
        self._arenas = [self.arena] if self.prelu_arena is self.arena else [self.arena, self.prelu_arena]
        A, PA, C = self.arena, self.prelu_arena, n_channels
        self.n_channels, self.n_blocks = C, n_blocks
        L = []
        self.c1 = ConvLayer(A, "conv_block1.conv_block.0", "conv", 3, C, 9, 1, 4, bias=True)
        self.p1 = _Scalar(PA, "conv_block1.conv_block.1")
        L += [self.c1, self.p1]
        self.blocks = []
        for i in range(n_blocks):
            inner = C if filter_cfgs is None else filter_cfgs[i]
            p = "residual_blocks.%d." % i
            b = (ConvLayer(A, p + "conv_block1.conv_block.0", "conv", C, inner, 3, 1, 1, bias=True),
                 NormLayer(A, p + "conv_block1.conv_block.1", inner, "bn", device),
                 _Scalar(PA, p + "conv_block1.conv_block.2"),
                 ConvLayer(A, p + "conv_block2.conv_block.0", "conv", inner, C, 3, 1, 1, bias=True),
                 NormLayer(A, p + "conv_block2.conv_block.1", C, "bn", device))
            self.blocks.append(b)
            L += list(b)
        self.c2 = ConvLayer(A, "conv_block2.conv_block.0", "conv", C, C, 3, 1, 1, bias=True)
        self.n2 = NormLayer(A, "conv_block2.conv_block.1", C, "bn", device)
        L += [self.c2, self.n2]
        self.sub = []
        for i in range(2):
            s = (ConvLayer(A, "subpixel_convolutional_blocks.%d.conv" % i, "conv", C, 4 * C, 3, 1, 1, bias=True),
                 _Scalar(PA, "subpixel_convolutional_blocks.%d.prelu" % i))
            self.sub.append(s)
            L += list(s)
        self.c3 = ConvLayer(A, "conv_block3.conv_block.0", "conv", C, 3, 9, 1, 4, bias=True)
        L.append(self.c3)
        self._layers = L
        self.taps = []
        if arena is None:
            self.finalize()

    def finalize(self):
        for a in self._arenas:
            if not a.finalized:
                a.finalize()
        self._register_sr(self._layers)

    def forward(self, x):
        C = self.n_channels
        out = PReLUFn.apply(self.c1(x), self.p1.weight)
        residual = out
        self.taps = []
        for i, (ca, na, pa, cb, nb) in enumerate(self.blocks):
            r = out
            t, s = ca.with_stats(out)
            t = PReLUFn.apply(na(t, ACT_NONE, None, s), pa.weight)
            t, s = cb.with_stats(t)
            out = ops.AddFn.apply(nb(t, ACT_NONE, None, s), r)
            if i in (3, 7, 11, 15):
                self.taps.append((out, C))
        t, s = self.c2.with_stats(out)
        out = ops.AddFn.apply(self.n2(t, ACT_NONE, None, s), residual)
        for conv, pr in self.sub:
            out = PReLUFn.apply(PixelShuffleFn.apply(conv(out), C), pr.weight)
        return self.c3(out, ACT_TANH)


def sr_disc_channels(ndf, n_blocks=4):
    chans, cin = [], 3
    for i in range(n_blocks):
        cout = (ndf if i == 0 else cin * 2) if i % 2 == 0 else cin
        chans.append((cin, cout, 1 if i % 2 == 0 else 2, i != 0))
        cin = cout
    return chans


class Discriminator(_SRNet):
    """SRGAN.py:199-245 (``gated=False``) / MaskDiscriminator :247-295 (``gated=True``): n_blocks x [Conv3 (stride 1/2)
    -> BN (not on block 0) -> [DifferentiableOP] -> LeakyReLU(0.2)], global average pool, Linear(C, 1).
    Taps: outputs of conv_blocks.1 and conv_blocks.3."""

    def __init__(self, kernel_size=3, n_channels=64, n_blocks=4, threshold=0.5, gated=False, arena=None, gate_arena=None,
                 device="cuda"):
        super().__init__()
        self.arena = arena if arena is not None else ParamArena(device)
        if gated and gate_arena is None:
            gate_arena = ParamArena(device)
        self.gate_arena = gate_arena
        self._arenas = [self.arena] + ([gate_arena] if gate_arena is not None else [])
        self.gated = gated
        L, self.blocks = [], []
        for i, (cin, cout, stride, has_bn) in enumerate(sr_disc_channels(n_channels, n_blocks)):
            p = "conv_blocks.%d.conv_block." % i
            conv = ConvLayer(self.arena, p + "0", "conv", cin, cout, 3, stride, 1, bias=True)
            gname = (p + ("2" if has_bn else "1")) if gated else None
            norm = NormLayer(self.arena, p + "1", cout, "bn" if has_bn else "id", device,
                             gate_arena=gate_arena if gated else None, gate_name=gname, thr=threshold)
            norm.gate_after = False      # conv -> [BN] -> gate -> LeakyReLU: the gate sits before the activation
            self.blocks.append((conv, norm))
            L += [conv, norm]
        self.cout = cout
        self.head = _Linear(self.arena, "fc1", cout)
        L.append(self.head)
        self._layers = L
        self.taps = []
        if arena is None:
            self.finalize()

    def finalize(self):
        for a in self._arenas:
            if not a.finalized:
                a.finalize()
        self._register_sr(self._layers)

    def forward(self, x):
        self.taps = []
        out = x
        for i, (conv, norm) in enumerate(self.blocks):
            if norm.mode == "id" and norm.alpha is None:
                out = conv(out, ACT_LRELU, 0.2)              # LeakyReLU fused into the conv epilogue
            elif norm.mode == "bn":
                t, s = conv.with_stats(out)
                out = norm(t, ACT_LRELU, None, s)
            else:
                out = norm(conv(out), ACT_LRELU)
            if i in (1, 3):
                self.taps.append((out, conv.cout))
        return PoolLinearFn.apply(out, self.head.weight, self.head.bias, self.cout)

    def get_current_masks(self):
        return [((torch.sign(n.alpha.detach() - n.thr) + 1) / 2) for _, n in self.blocks if n.alpha is not None]


class MaskDiscriminator(Discriminator):
    def __init__(self, kernel_size=3, n_channels=64, n_blocks=4, threshold=0.5, arena=None, gate_arena=None,
                 device="cuda"):
        super().__init__(kernel_size, n_channels, n_blocks, threshold, True, arena, gate_arena, device)


class TruncatedVGG19(_SRNet):
    """torchvision vgg19.features[:36] (GANLoss.py:95-144 with i=5, j=4): 16 x [Conv3 + ReLU], 4 x MaxPool2d(2, 2).
    Frozen: only forward and the data gradient are computed."""

    def __init__(self, i=5, j=4, device="cuda"):
        super().__init__()
        assert (i, j) == (5, 4)
        self.arena = ParamArena(device)
        self._arenas = [self.arena]
        self.seq, L, cin, idx = [], [], 3, 0
        for v in VGG_CFG:
            if v == "M":
                self.seq.append(None)
                idx += 1
                continue
            conv = ConvLayer(self.arena, "truncated_vgg19.%d" % idx, "conv", cin, v, 3, 1, 1, bias=True)
            self.seq.append(conv)
            L.append(conv)
            cin = v
            idx += 2
        self._layers = L
        self.arena.finalize()
        self._register_sr(L)
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        for l in self.seq:
            x = MaxPoolFn.apply(x) if l is None else l(x, ACT_LRELU, 0.0)     # LeakyReLU with slope 0 = ReLU
        return x


# ------------------------------------------------------------------------------------------------------- model
class SRGAN(GccModelMixin, torch.nn.Module):
    """models/SRGAN.py:296-842 on the B200 kernels."""

    def __init__(self, opt, filter_cfgs=None, channel_cfgs=None):
        super().__init__()
        self.opt = opt
        if len(opt.gpu_ids) == 0:
            raise RuntimeError("gcc_b200.SRGAN needs a CUDA device (gpu_ids): there is no CPU path")
        self.device = dev = torch.device("cuda:%d" % opt.gpu_ids[0])
        call("gcc_check_device")
        self.filter_cfgs, self.channel_cfgs = filter_cfgs, channel_cfgs
        self.current_epoch = 0
        self.current_D_arch_diff_loss = 0.0
        self.teacher_model = None
        self._base_init()
        self.visual_names = ["real_lr", "fake_hr", "real_hr"]
        generator_only = bool(getattr(opt, "generator_only", False))
        self.loss_names = ["content"] if generator_only else ["G_GAN", "D_real", "D_fake", "content", "perceptual"]
        self.generator_extract_layers = ["residual_blocks.3", "residual_blocks.7", "residual_blocks.11",
                                         "residual_blocks.15"]
        self.discriminator_extract_layers = ["conv_blocks.1", "conv_blocks.3"]
        self.distill = bool(opt.online_distillation or getattr(opt, "normal_distillation", False))

        # optimizer_G: all generator parameters; under distillation the transform convs come first and the PReLU
        # slopes are left out (SRGAN.py:337-353: only Conv / BatchNorm / Linear modules are collected)
        self.arena_G = ParamArena(dev)
        self.arena_P = ParamArena(dev) if self.distill else None
        self.transform_convs = []
        if self.distill:
            for i in range(4):
                self.transform_convs.append(ConvLayer(self.arena_G, "transform.%d" % i, "conv", opt.ngf, opt.teacher_ngf,
                                                      1, 1, 0))
        self.netG = Generator(n_channels=opt.ngf, filter_cfgs=filter_cfgs, arena=self.arena_G, prelu_arena=self.arena_P,
                              device=dev)
        self.netG.finalize()
        for t in self.transform_convs:
            t.bind()
        self.optimizer_G = _ArenaOptimizer(self.arena_G, opt.lr, (0.9, 0.999))

        self.truncated_vgg19 = TruncatedVGG19(5, 4, device=dev)

        self.arena_D = ParamArena(dev)
        if opt.darts_discriminator:
            self.loss_names += ["D_arch_diff", "D_arch", "teacher_D_arch_diff"]
            self.arena_A = ParamArena(dev)
            self.netD = MaskDiscriminator(n_channels=opt.ndf, threshold=opt.threshold, arena=self.arena_D,
                                          gate_arena=self.arena_A, device=dev)
            self.netD.finalize()
            self.optimizer_D = _ArenaOptimizer(self.arena_D, opt.lr, (0.9, 0.999))
            self.optimizer_arch = _ArenaOptimizer(self.arena_A, opt.arch_lr, (0.9, 0.999))
        else:
            self.arena_A = None
            self.netD = Discriminator(n_channels=opt.ndf, arena=self.arena_D, device=dev)
            self.netD.finalize()
            self.optimizer_D = _ArenaOptimizer(self.arena_D, opt.lr, (0.9, 0.999))
        self.optimizers = []
        if opt.darts_discriminator and getattr(opt, "arch_lr_step", False):
            self.optimizers.append(self.optimizer_arch)
        if generator_only:
            self.optimizers.clear()
        self.init_net()
        self.gan_mode = GAN_MODES.get(opt.gan_mode)
        if self.gan_mode is None:
            raise NotImplementedError("gan mode %s not implemented" % opt.gan_mode)
        self.optimizers.append(self.optimizer_G)
        self.optimizers.append(self.optimizer_D)
        self.schedulers = [get_scheduler(o, opt) for o in self.optimizers]
        self._consts = _imagenet_consts(dev)
        self.total_generator_features, self.total_discriminator_features = {}, {}
        self.broadcast_parameters()

    def _gcc_arenas(self):
        return {k: a for k, a in (("G", self.arena_G), ("P", self.arena_P), ("D", self.arena_D), ("A", self.arena_A))
                if a is not None}

    def _gcc_optimizers(self):
        d = {"G": self.optimizer_G, "D": self.optimizer_D}
        if self.arena_A is not None:
            d["arch"] = self.optimizer_arch
        return d

    def _gcc_nets(self):
        return {"netG": self.netG, "netD": self.netD}

    # ------------------------------------------------------------------ init (util.init_weights + nn.PReLU default)
    def _arenas(self):
        return [a for a in (self.arena_G, self.arena_P, self.arena_D, self.arena_A) if a is not None]

    def init_net(self):
        with torch.no_grad():
            for arena in (self.arena_G, self.arena_P, self.arena_D):
                if arena is None:
                    continue
                for (name, shape, kind) in arena.specs:
                    p = arena.params[name]
                    if name.startswith("transform."):
                        bound = 1.0 / math.sqrt(shape[1])
                        p.copy_(torch.empty(shape, device=self.device).uniform_(-bound, bound))
                    elif len(shape) in (2, 4):
                        p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 0.02))
                    elif shape == (1,) and name.endswith(".weight") and "fc1" not in name:
                        p.fill_(0.25)                                            # nn.PReLU()
                    elif name.endswith(".weight"):
                        p.copy_(torch.empty(shape, device=self.device).normal_(1.0, 0.02))
                    elif name.endswith(".bias"):
                        wname = name[:-4] + "weight"
                        is_bn = wname in arena.params and len(arena.params[wname].shape) == 1 and "fc1" not in name
                        if is_bn:
                            p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 1.0))
                        else:
                            p.zero_()
                arena.mark_dirty()
            # the truncated VGG19 has no pretrained file here: He-normal stand-in (load real weights with
            # truncated_vgg19.load_state_dict(torchvision_vgg19.features[:36].state_dict()))
            for (name, shape, kind) in self.truncated_vgg19.arena.specs:
                p = self.truncated_vgg19.arena.params[name]
                if len(shape) == 4:
                    p.copy_(torch.empty(shape, device=self.device).normal_(0.0, math.sqrt(2.0 / (shape[1] * 9))))
                else:
                    p.zero_()
            self.truncated_vgg19.arena.mark_dirty()

    def sync_weights(self):
        for a in self._arenas() + [self.truncated_vgg19.arena]:
            a.mark_dirty()

    # ------------------------------------------------------------------ inputs / forward
    def set_input(self, input):
        self.input = input
        self._lr_nchw = input["lr"].to(self.device, non_blocking=True)
        self._hr_nchw = input["hr"].to(self.device, non_blocking=True)
        self.image_paths = [input.get("lr_names"), input.get("hr_names")]
        self.real_lr_nhwc = ops.to_nhwc(self._lr_nchw)
        self.real_hr_nhwc = ops.to_nhwc(self._hr_nchw)
        self._hr_raw_nhwc = self.real_hr_nhwc

    def _adopt_input(self, other):
        """Teacher side of ``T.set_input(self.input)`` (SRGAN.py:487,508): share the converted batch (the hr image
        BEFORE the student's in-place re-binding to its imagenet-normalised version)."""
        if other.device != self.device:
            return self.set_input(other.input)
        self.input, self.image_paths = other.input, other.image_paths
        self._lr_nchw, self._hr_nchw = other._lr_nchw, other._hr_nchw
        self.real_lr_nhwc = other.real_lr_nhwc
        self.real_hr_nhwc = self._hr_raw_nhwc = other._hr_raw_nhwc

    @property
    def real_lr(self):
        return self._lr_nchw

    @property
    def real_hr(self):
        return ops.to_nchw(self.real_hr_nhwc.detach(), 3)

    @property
    def fake_hr(self):
        return ops.to_nchw(self.fake_hr_nhwc.detach(), 3)

    def forward(self):
        self.fake_hr_nhwc = self.netG(self.real_lr_nhwc)
        self.g_taps = list(self.netG.taps)

    def _D(self, net, x):
        pred = net(x)
        return pred, list(net.taps)

    def _gan(self, pred, kind):
        return ops.GanLossFn.apply(pred, 1, self.gan_mode, kind)

    def _to_imagenet(self):
        """The reference re-binds real_hr / fake_hr to their imagenet-normalised versions (SRGAN.py:419-420,444-445)."""
        self.real_hr_nhwc = convert_to_imagenet(self.real_hr_nhwc, self._consts)
        self.fake_hr_nhwc = convert_to_imagenet(self.fake_hr_nhwc, self._consts)

    # ------------------------------------------------------------------ losses / steps
    def backward_D(self):
        pred_real, _ = self._D(self.netD, self.real_hr_nhwc)
        pred_fake, self.d_taps = self._D(self.netD, self.fake_hr_nhwc.detach())
        self.loss_D_real = self._gan(pred_real, 0)
        self.loss_D_fake = self._gan(pred_fake, 1)
        self.loss_D = self.loss_D_real + self.loss_D_fake
        self.loss_D.backward()

    def get_D_arch_diff(self, isTeacher=False):
        self._to_imagenet()
        if isTeacher:
            self.set_requires_grad(self.netD, False)
        pred_fake, _ = self._D(self.netD, self.fake_hr_nhwc.detach())
        self.loss_D_arch_fake = self._gan(pred_fake, 1)
        self.loss_D_arch_fake_real = self._gan(pred_fake, 2)
        pred_real, _ = self._D(self.netD, self.real_hr_nhwc)
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
        self.loss_content = ops.DiffLossFn.apply(self.fake_hr_nhwc, self.real_hr_nhwc, 3, 2) * o.lambda_SR_content
        self._to_imagenet()
        pred_fake, self.d_taps = self._D(self.netD, self.fake_hr_nhwc)
        self.loss_G_GAN = self._gan(pred_fake, 2) * o.lambda_SR_adversarial
        fake_vgg = self.truncated_vgg19(self.fake_hr_nhwc)
        with torch.no_grad():
            real_vgg = self.truncated_vgg19(self.real_hr_nhwc)
        self.loss_perceptual = ops.DiffLossFn.apply(fake_vgg, real_vgg, 512, 2) * o.lambda_SR_perceptual
        self.loss_G = self.loss_content + self.loss_G_GAN + self.loss_perceptual
        if self.distill:
            T = self.teacher_model
            self.Tfake_hr_nhwc = T.fake_hr_nhwc.detach()
            feats = list(self.g_taps)
            # teacher D on the student fake.  (The reference leaves the teacher D's requires_grad on here; the
            # gradients it collects are zeroed before the teacher's next D step and never used: skip computing them.)
            T.set_requires_grad(T.netD, False)
            _, t_d_taps = self._D(T.netD, self.fake_hr_nhwc)
            feats += t_d_taps
            self.loss_content = 0.0
            self.loss_gram = 0.0
            for i, (f, c) in enumerate(feats):
                if i < 4:
                    f = self.transform_convs[i](f)
                    c = self.transform_convs[i].cout
                tgt, gram_t = self.target_distillation_features[i], self.target_grams[i]
                self.loss_gram = self.loss_gram + ops.GramRmseFn.apply(f, gram_t, c)
                self.loss_content = self.loss_content + ops.DiffLossFn.apply(f, tgt, c, 1)
            self.loss_gram = o.lambda_gram * self.loss_gram
            self.loss_content = o.lambda_content * self.loss_content
            self.loss_L1 = o.lambda_L1 * ops.DiffLossFn.apply(self.fake_hr_nhwc, self.Tfake_hr_nhwc, 3, 0)
            self.loss_G = self.loss_G + self.loss_gram + self.loss_content + self.loss_L1
        self.loss_G.backward()

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
        self.set_requires_grad(self.netD, False)
        self.optimizer_G.zero_grad()
        if self.arena_P is not None:
            self.arena_P.zero_grad()
        self.backward_G()
        self.optimizer_G.step()
        self.set_requires_grad(self.netD, True)
        self.set_netD_arch_grad(False)
        self.optimizer_D.zero_grad()
        self.backward_D()
        self.optimizer_D.step()
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

    def optimize_content_parameters(self):
        self.forward()
        self.optimizer_G.zero_grad()
        self.loss_content = ops.DiffLossFn.apply(self.fake_hr_nhwc, self.real_hr_nhwc, 3, 2)
        self.loss_content.backward()
        self.L1_sparsity()
        self.optimizer_G.step()
        self._release_graphs()

    def L1_sparsity(self):
        o, st = self.opt, _st()
        names = [s[0] for s in self.arena_G.specs]
        for (name, shape, kind) in self.arena_G.specs:
            if name.startswith("transform."):
                continue
            off, n = self.arena_G.offsets[names.index(name)]
            if o.lambda_weight > 0.0 and len(shape) == 4:
                call("gcc_l1_sparsity_f32", self.arena_G.P[off:].data_ptr(), self.arena_G.G[off:].data_ptr(), n,
                     o.lambda_weight, st)
            elif o.lambda_weight <= 0.0 and o.lambda_scale > 0.0 and len(shape) == 1 and name.endswith(".weight") and \
                    (name[:-6] + "bias") in names:          # BatchNorm scales (PReLU slopes have no bias)
                call("gcc_l1_sparsity_f32", self.arena_G.P[off:].data_ptr(), self.arena_G.G[off:].data_ptr(), n,
                     o.lambda_scale, st)

    # ------------------------------------------------------------------ bookkeeping (reference surface)
    def print_sparse_info(self, logger):
        for i, mask in enumerate(self.netD.get_current_masks() if self.opt.darts_discriminator else []):
            logger.info("netD gate %d sparsity ratio: %.2f" % (i, float((mask == 0.0).sum()) / mask.numel()))

    def adaptive_ema_beta(self, epoch):
        self.opt.ema_beta = 1.0 - epoch / (self.opt.n_epochs + self.opt.n_epochs_decay)

    def update_learning_rate(self, epoch):
        for s in self.schedulers:
            s.step()
        self.adaptive_ema_beta(epoch)
        self.current_epoch = epoch
        print("learning rate = %.7f" % self.optimizers[0].param_groups[0]["lr"])

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad

    def set_netD_weight_grad(self, requires_grad=False):
        for p in self.arena_D.params.values():
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
            self.visual_names.append("Tfake_hr")

    @property
    def Tfake_hr(self):
        return ops.to_nchw(self.Tfake_hr_nhwc, 3)

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
                "cfg": (self.filter_cfgs, self.channel_cfgs), "psnr": fid}
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
        return ckpt["psnr"], float("inf")

    # ------------------------------------------------------------------ pruning front-end (SRGAN.py:707-837)
    def prune(self, threshold, lottery_path=None):
        if self.opt.scale_prune:
            return self.scale_prune(threshold, lottery_path)
        if self.opt.norm_prune:
            return self.norm_prune(threshold, lottery_path)
        raise NotImplementedError("only scale and norm pruning are supported!!!")

    def _named_convs_bns(self):
        sd = self.netG.state_dict()
        convs = [(k[:-len(".weight")], v) for k, v in sd.items() if v.dim() == 4]
        bns = [(k[:-len(".weight")], v) for k, v in sd.items()
               if k.endswith(".weight") and v.dim() == 1 and (k[:-6] + "running_mean") in sd]
        return convs, bns

    def max_min_bn_scale(self):
        """SRGAN.py:719-737 (its unprunable list holds 'conv_block2.conv_block.1' and, 16 times,
        'residual_blocks.15.conv_block2.conv_block.1')."""
        unprunable = {"conv_block2.conv_block.1", "residual_blocks.15.conv_block2.conv_block.1"}
        pmax, omin = float("inf"), float("inf")
        for name, w in self._named_convs_bns()[1]:
            if name in unprunable:
                continue
            pmax = min(torch.max(w), pmax)
            omin = min(torch.min(w), omin)
        return pmax, omin

    def max_min_conv_norm(self):
        """SRGAN.py:739-768 (the missing comma of the reference's list fuses two names into one that matches nothing)."""
        unprunable = {"conv_block3.0", "conv_block2.conv_block.0subpixel_convolutional_blocks.0.conv",
                      "subpixel_convolutional_blocks.1.conv", "residual_blocks.15.conv_block2.conv_block.0"}
        pmax, omin = float("inf"), float("inf")
        for name, w in self._named_convs_bns()[0]:
            if name in unprunable:
                continue
            norm = torch.sum(torch.abs(w), (0, 2, 3))
            pmax = min(torch.max(norm), pmax)
            omin = min(torch.min(norm), omin)
        return pmax, omin

    def norm_prune(self, threshold, lottery_path=None):
        unprunable = {"conv_block1.conv_block.0", "conv_block2.conv_block.0", "subpixel_convolutional_blocks.0.conv",
                      "subpixel_convolutional_blocks.1.conv"} | \
            {"residual_blocks.%d.conv_block2.conv_block.0" % i for i in range(16)}
        cfg = [int((torch.sum(torch.abs(w), (1, 2, 3)) > threshold).sum()) for name, w in self._named_convs_bns()[0]
               if name not in unprunable]
        return SRGAN(self.opt, filter_cfgs=cfg)

    def scale_prune(self, threshold, lottery_path=None):
        unprunable = {"conv_block2.conv_block.1"} | {"residual_blocks.%d.conv_block2.conv_block.1" % i for i in range(16)}
        cfg = [int((w > threshold).sum()) for name, w in self._named_convs_bns()[1] if name not in unprunable]
        return SRGAN(self.opt, filter_cfgs=cfg)


def build_sr_teacher(model, opt):
    """train.py:92-105: the teacher is the same class at teacher widths with a plain discriminator."""
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    topt.generator_only = False
    teacher = SRGAN(topt)
    teacher.model_train()
    model.teacher_model = teacher
    model.init_distillation()
    teacher.init_distillation()
    return teacher
