"""MobileCycleGANModel: the reference's CycleGAN GCC wrapper (/root/reference/models/CycleGAN.py:223-897) on
the B200 kernels.  Same constructor (`opt, cfg_AtoB, cfg_BtoA`), method and attribute names as the reference so
that `train.py` / `test.py` drive it unchanged.

Deliberate differences: NHWC bf16 activations, fused arena optimizers, NCCL gradient all-reduce when a process
group exists, and `forward()` runs 6 generator passes instead of 8: the reference recomputes G_A(real_A) and
G_B(real_B) a second time only "to get hook" (CycleGAN.py:381-382); the second pass is numerically the same
function of the same weights, so the first pass's output and taps are reused (the gradients add up identically).
"""
import copy
import os
import random
from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops
from .arena import ParamArena
from .base import GccModelMixin
from .nets import ConvLayer, MaskNLayerDiscriminator, MobileResnetGenerator, NLayerDiscriminator
from .ops import GAN_MODES
from .pix2pix import _ArenaOptimizer, get_scheduler


class ImagePool:
    """History buffer of generated images (utils/image_pool.py:22-54), on NHWC bf16 device tensors."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.images = []

    def query(self, images):
        if self.pool_size == 0:
            return images.detach()
        out = []
        for image in images.detach().unbind(0):
            image = image.unsqueeze(0)
            if len(self.images) < self.pool_size:
                self.images.append(image)
                out.append(image)
            elif random.uniform(0, 1) > 0.5:
                i = random.randint(0, self.pool_size - 1)
                out.append(self.images[i].clone())
                self.images[i] = image
            else:
                out.append(image)
        return torch.cat(out, 0)


class DeviceImagePool:
    """ImagePool with the buffer, its fill count and the random stream on the device (`gcc_image_pool_query_bf16`):
    same policy as utils/image_pool.py:22-54, but capturable in a CUDA graph -- every replay draws fresh decisions.
    The random stream is a device counter (not Python's `random`), so the swap choices differ from a host pool's."""

    def __init__(self, pool_size, seed=0):
        self.pool_size = pool_size
        self.buf = self.state = self.dec = None
        self.seed = seed

    def query(self, images):
        images = images.detach().contiguous()
        if self.pool_size == 0:
            return images
        b = images.shape[0]
        per = images[0].numel()
        if self.buf is None:
            self.buf = torch.zeros((self.pool_size,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
            self.state = torch.tensor([0, self.seed], dtype=torch.int64, device=images.device)
        if self.dec is None or self.dec.numel() < 2 * b:
            self.dec = torch.zeros(2 * b, dtype=torch.int32, device=images.device)
        out = torch.empty_like(images)
        ops.call("gcc_image_pool_query_bf16", images.data_ptr(), self.buf.data_ptr(), self.state.data_ptr(),
                 self.dec.data_ptr(), out.data_ptr(), b, per, self.pool_size, ops._st())
        return out


class MobileCycleGANModel(GccModelMixin, nn.Module):

    def __init__(self, opt, cfg_AtoB=None, cfg_BtoA=None):
        super().__init__()
        self.opt = opt
        if len(opt.gpu_ids) == 0:
            raise RuntimeError("gcc_b200.MobileCycleGANModel needs a CUDA device (gpu_ids): there is no CPU path")
        self.device = dev = torch.device("cuda:%d" % opt.gpu_ids[0])
        ops.call("gcc_check_device")
        self.cfg_AtoB, self.cfg_BtoA = cfg_AtoB, cfg_BtoA
        self.loss_names = ["D_A", "G_A", "cycle_A", "idt_A", "D_B", "G_B", "cycle_B", "idt_B"]
        self.visual_names = ["real_A", "fake_B", "rec_A", "idt_B", "real_B", "fake_A", "rec_B", "idt_A"]
        self.generator_extract_layers = ["model.9", "model.12", "model.15", "model.18"]
        self.discriminator_extract_layers = ["model.4", "model.12"] if opt.darts_discriminator else ["model.3", "model.9"]
        self.distill = bool(opt.online_distillation or getattr(opt, "normal_distillation", False))
        self.teacher_model = None
        self._base_init()

        self.arena_G = ParamArena(dev)
        self.transform_A_convs, self.transform_B_convs = [], []
        if self.distill:
            for key, cfg, lst in (("A", cfg_AtoB, self.transform_A_convs), ("B", cfg_BtoA, self.transform_B_convs)):
                s_ch = opt.ngf * 4 if cfg is None else cfg[2]
                for i in range(4):
                    lst.append(ConvLayer(self.arena_G, "transform_%s.%d" % (key, i), "conv", s_ch, opt.teacher_ngf * 4, 1, 1, 0))
        self.netG_A = MobileResnetGenerator(ngf=opt.ngf, cfg=cfg_AtoB, arena=self.arena_G, device=dev, prefix="A.")
        self.netG_B = MobileResnetGenerator(ngf=opt.ngf, cfg=cfg_BtoA, arena=self.arena_G, device=dev, prefix="B.")
        self.arena_G.finalize()
        self.netG_A.finalize()
        self.netG_B.finalize()
        for t in self.transform_A_convs + self.transform_B_convs:
            t.bind()
        self.optimizer_G = _ArenaOptimizer(self.arena_G, opt.lr, (0.5, 0.999))

        self.arena_D = ParamArena(dev)
        if opt.darts_discriminator:
            self.loss_names += ["D_arch_diff_A", "D_arch_A", "D_arch_diff_B", "D_arch_B", "teacher_netD_A_arch_diff",
                                "teacher_netD_B_arch_diff"]
            self.arena_A = ParamArena(dev)
            mk = lambda p: MaskNLayerDiscriminator(input_nc=3, ndf=opt.ndf, threshold=opt.threshold, arena=self.arena_D,
                                                   gate_arena=self.arena_A, device=dev, prefix=p)
        else:
            self.arena_A = None
            mk = lambda p: NLayerDiscriminator(input_nc=3, ndf=opt.ndf, arena=self.arena_D, device=dev, norm="in", prefix=p)
        self.netD_A, self.netD_B = mk("A."), mk("B.")
        self.arena_D.finalize()
        if self.arena_A is not None:
            self.arena_A.finalize()
        self.netD_A.finalize()
        self.netD_B.finalize()
        self.optimizer_D = _ArenaOptimizer(self.arena_D, opt.lr, (0.5, 0.999))
        if self.arena_A is not None:
            self.optimizer_arch = _ArenaOptimizer(self.arena_A, opt.arch_lr, (0.9, 0.999))
        self.init_net()

        self.fake_A_pool, self.fake_B_pool = ImagePool(50), ImagePool(50)
        self.gan_mode = GAN_MODES.get(opt.gan_mode)
        if self.gan_mode is None:
            raise NotImplementedError("gan mode %s not implemented" % opt.gan_mode)
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.schedulers = [get_scheduler(o, opt) for o in self.optimizers]
        if opt.darts_discriminator and getattr(opt, "arch_lr_step", False):
            arch_opt = copy.deepcopy(opt)
            arch_opt.lr_policy, arch_opt.lr_decay_iters = "step", opt.n_epochs - 1
            self.arch_scheduler = get_scheduler(self.optimizer_arch, arch_opt)
            self.schedulers.append(self.arch_scheduler)
        self.broadcast_parameters()

    def _gcc_arenas(self):
        return {k: a for k, a in (("G", self.arena_G), ("D", self.arena_D), ("A", self.arena_A)) if a is not None}

    def _gcc_optimizers(self):
        d = {"G": self.optimizer_G, "D": self.optimizer_D}
        if self.arena_A is not None:
            d["arch"] = self.optimizer_arch
        return d

    def _gcc_nets(self):
        return {"netG_A": self.netG_A, "netG_B": self.netG_B, "netD_A": self.netD_A, "netD_B": self.netD_B}

    def _gcc_extra_state(self):
        st = {}
        for k, pool in (("A", self.fake_A_pool), ("B", self.fake_B_pool)):
            if isinstance(pool, DeviceImagePool):
                st[k] = ("device", None if pool.buf is None else pool.buf.cpu(), None if pool.state is None else pool.state.cpu())
            else:
                st[k] = ("host", [t.detach().cpu() for t in pool.images], None)
        return st

    def _gcc_load_extra_state(self, st):
        for k, name in (("A", "fake_A_pool"), ("B", "fake_B_pool")):
            kind, data, state = st[k]
            if kind == "device":
                self.use_device_pools()
                pool = getattr(self, name)
                if data is not None:
                    pool.buf, pool.state = data.to(self.device), state.to(self.device)
            else:
                getattr(self, name).images = [t.to(self.device) for t in data]

    def use_device_pools(self):
        """Swap the host-random image pools for device-resident ones (needed before CUDA-graph capture)."""
        if not isinstance(self.fake_A_pool, DeviceImagePool):
            rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
            self.fake_A_pool = DeviceImagePool(self.fake_A_pool.pool_size, seed=1 + 7919 * rank)
            self.fake_B_pool = DeviceImagePool(self.fake_B_pool.pool_size, seed=2 + 7919 * rank)

    def init_net(self):
        """util.init_weights (utils/util.py:261-286); transform convs keep the nn.Conv2d default; alpha = 1."""
        import math
        with torch.no_grad():
            for arena in (self.arena_G, self.arena_D):
                for (name, shape, kind) in arena.specs:
                    p = arena.params[name]
                    if name.startswith("transform_"):
                        b = 1.0 / math.sqrt(shape[1])
                        p.copy_(torch.empty(shape, device=self.device).uniform_(-b, b))
                    elif len(shape) == 4:
                        p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 0.02))
                    elif name.endswith(".weight"):
                        p.copy_(torch.empty(shape, device=self.device).normal_(1.0, 0.02))
                    elif name.endswith(".bias"):
                        wn = name[:-4] + "weight"
                        if wn in arena.params and len(arena.params[wn].shape) == 1:
                            p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 1.0))
                        else:
                            p.zero_()
                arena.mark_dirty()

    def sync_weights(self):
        for a in (self.arena_G, self.arena_D, self.arena_A):
            if a is not None:
                a.mark_dirty()

    # ------------------------------------------------------------------ inputs / forward
    def set_input(self, input):
        self.input = input
        AtoB = self.opt.direction == "AtoB"
        A = input["A" if AtoB else "B"].to(self.device, non_blocking=True)
        B = input["B" if AtoB else "A"].to(self.device, non_blocking=True)
        self.image_paths = [input.get("A_paths" if AtoB else "B_paths"), input.get("B_paths" if AtoB else "A_paths")]
        self.real_A, self.real_B = A, B
        self.real_A_nhwc, self.real_B_nhwc = ops.to_nhwc(A), ops.to_nhwc(B)

    def _adopt_input(self, other):
        """Teacher side of ``T.set_input(self.input)`` (CycleGAN.py:569,592): share the converted batch."""
        if other.opt.direction != self.opt.direction or other.device != self.device:
            return self.set_input(other.input)
        self.input, self.image_paths = other.input, other.image_paths
        self.real_A, self.real_B = other.real_A, other.real_B
        self.real_A_nhwc, self.real_B_nhwc = other.real_A_nhwc, other.real_B_nhwc

    def _nchw(self, name):
        return ops.to_nchw(getattr(self, name + "_nhwc").detach(), 3)

    fake_A = property(lambda self: self._nchw("fake_A"))
    fake_B = property(lambda self: self._nchw("fake_B"))
    rec_A = property(lambda self: self._nchw("rec_A"))
    rec_B = property(lambda self: self._nchw("rec_B"))
    idt_A = property(lambda self: self._nchw("idt_A"))
    idt_B = property(lambda self: self._nchw("idt_B"))
    Tfake_A = property(lambda self: self.teacher_model.fake_A)
    Tfake_B = property(lambda self: self.teacher_model.fake_B)

    def forward(self):
        self.fake_B_nhwc = self.netG_A(self.real_A_nhwc)
        self.g_taps_A = list(self.netG_A.taps)
        self.rec_A_nhwc = self.netG_B(self.fake_B_nhwc)
        self.fake_A_nhwc = self.netG_B(self.real_B_nhwc)
        self.g_taps_B = list(self.netG_B.taps)
        self.rec_B_nhwc = self.netG_A(self.fake_A_nhwc)
        self.idt_A_nhwc = self.netG_A(self.real_B_nhwc)
        self.idt_B_nhwc = self.netG_B(self.real_A_nhwc)

    def visual_forward(self):
        self.fake_B_nhwc = self.netG_A(self.real_A_nhwc)

    def _gan(self, pred, kind):
        return ops.GanLossFn.apply(pred, 1, self.gan_mode, kind)

    def _D(self, net, x):
        pred = net(x)
        return pred, list(net.taps)

    # ------------------------------------------------------------------ losses / steps
    def backward_D_basic(self, netD, real, fake):
        pred_real, _ = self._D(netD, real)
        loss_real = self._gan(pred_real, 0)
        pred_fake, taps = self._D(netD, fake.detach())
        loss_fake = self._gan(pred_fake, 1)
        loss = (loss_real + loss_fake) * 0.5
        loss.backward()
        return loss, taps

    def backward_D_A(self):
        fake_B = self.fake_B_pool.query(self.fake_B_nhwc)
        self.loss_D_A, self.d_taps_A = self.backward_D_basic(self.netD_A, self.real_B_nhwc, fake_B)

    def backward_D_B(self):
        fake_A = self.fake_A_pool.query(self.fake_A_nhwc)
        self.loss_D_B, self.d_taps_B = self.backward_D_basic(self.netD_B, self.real_A_nhwc, fake_A)

    def get_D_arch_diff(self, isTeacher=False):
        if isTeacher:
            self.set_requires_grad([self.netD_A, self.netD_B], False)
        out = {}
        for k, net, fake, real in (("A", self.netD_A, self.fake_B_nhwc, self.real_B_nhwc),
                                   ("B", self.netD_B, self.fake_A_nhwc, self.real_A_nhwc)):
            pred_fake, _ = self._D(net, fake.detach())
            l_fake = self._gan(pred_fake, 1)
            l_fake_real = self._gan(pred_fake, 2)
            pred_real, _ = self._D(net, real)
            l_real = self._gan(pred_real, 0)
            setattr(self, "loss_D_%s_arch_fake" % k, l_fake)
            setattr(self, "loss_D_%s_arch_fake_real" % k, l_fake_real)
            setattr(self, "loss_D_%s_arch_real" % k, l_real)
            diff = (l_fake_real - l_fake).abs()
            if isTeacher:  # EMA state in one persistent device scalar per discriminator (base.py)
                out[k] = self._ema_update(k, diff)
            else:
                out[k] = diff
        self.current_netD_A_arch_diff_loss, self.current_netD_B_arch_diff_loss = out["A"], out["B"]
        return out["A"], out["B"]

    def backward_D_arch(self):
        self.loss_teacher_netD_A_arch_diff, self.loss_teacher_netD_B_arch_diff = \
            self.teacher_model.get_D_arch_diff(isTeacher=True)
        self.loss_D_arch_diff_A, self.loss_D_arch_diff_B = self.get_D_arch_diff(isTeacher=False)
        self.loss_D_arch_A = (self.loss_D_arch_diff_A - self.loss_teacher_netD_A_arch_diff.detach()).abs() + \
            (self.loss_D_A_arch_fake + self.loss_D_A_arch_real) * 0.5
        self.loss_D_arch_B = (self.loss_D_arch_diff_B - self.loss_teacher_netD_B_arch_diff.detach()).abs() + \
            (self.loss_D_B_arch_fake + self.loss_D_B_arch_real) * 0.5
        self.loss_D_arch_A.backward()
        self.loss_D_arch_B.backward()

    def backward_G(self):
        o = self.opt
        l1 = lambda a, b: ops.DiffLossFn.apply(a, b, 3, 0)
        self.loss_idt_A = l1(self.idt_A_nhwc, self.real_B_nhwc) * (o.lambda_B * o.lambda_identity)
        self.loss_idt_B = l1(self.idt_B_nhwc, self.real_A_nhwc) * (o.lambda_A * o.lambda_identity)
        self.loss_G_A = self._gan(self._D(self.netD_A, self.fake_B_nhwc)[0], 0)   # lsgan: target real
        self.loss_G_B = self._gan(self._D(self.netD_B, self.fake_A_nhwc)[0], 0)
        self.loss_cycle_A = l1(self.rec_A_nhwc, self.real_A_nhwc) * o.lambda_A
        self.loss_cycle_B = l1(self.rec_B_nhwc, self.real_B_nhwc) * o.lambda_B
        self.loss_G = self.loss_G_A + self.loss_G_B + self.loss_cycle_A + self.loss_cycle_B + self.loss_idt_A + self.loss_idt_B
        if self.distill:
            T = self.teacher_model
            for k, fake, tfake, taps, tnet, tconvs in (
                    ("A", self.fake_B_nhwc, T.fake_B_nhwc, self.g_taps_A, T.netD_A, self.transform_A_convs),
                    ("B", self.fake_A_nhwc, T.fake_A_nhwc, self.g_taps_B, T.netD_B, self.transform_B_convs)):
                with torch.no_grad():   # teacher D on the DETACHED student fake (CycleGAN.py:492-493): constants
                    _, t_d = self._D(tnet, fake.detach())
                feats = list(taps) + t_d
                gram_l = content_l = l1_l = 0.0
                targets, grams = getattr(self, "target_distillation_%s_features" % k), getattr(self, "target_grams_" + k)
                for i, (f, c) in enumerate(feats):
                    if i < 4:
                        f = tconvs[i](f)
                        c = tconvs[i].cout
                    gram_l = gram_l + ops.GramRmseFn.apply(f, grams[i], c, True)
                    content_l = content_l + ops.DiffLossFn.apply(f, targets[i], c, 2)
                    if o.lambda_L1 > 0.0:
                        l1_l = l1_l + l1(fake, tfake.detach())
                setattr(self, "loss_gram_" + k, o.lambda_gram * gram_l)
                setattr(self, "loss_content_" + k, o.lambda_content * content_l)
                setattr(self, "loss_L1_" + k, o.lambda_L1 * l1_l)
            self.loss_G = self.loss_G + self.loss_gram_A + self.loss_gram_B + self.loss_content_A + self.loss_content_B
            if o.lambda_L1 > 0.0:
                self.loss_G = self.loss_G + self.loss_L1_A + self.loss_L1_B
        self.loss_G.backward()
        self.L1_sparsity()

    def L1_sparsity(self):
        """CycleGAN.py:539-559: lambda (x2 on model.1/.4/.22, x1000 on model.19) * sign(w) on every conv weight."""
        lam = self.opt.lambda_weight
        if lam <= 0.0:
            return
        names = [s[0] for s in self.arena_G.specs]
        st = ops._st()
        for (name, shape, kind) in self.arena_G.specs:
            if len(shape) != 4 or name.startswith("transform_"):
                continue
            short = name[2:-len(".weight")]
            scale = 1000.0 if short == "model.19" else (2.0 if short in ("model.1", "model.4", "model.22") else 1.0)
            off, n = self.arena_G.offsets[names.index(name)]
            ops.call("gcc_l1_sparsity_f32", self.arena_G.P[off:].data_ptr(), self.arena_G.G[off:].data_ptr(), n, lam * scale, st)

    def _release_graphs(self):
        for k, v in list(vars(self).items()):
            if isinstance(v, torch.Tensor) and v.grad_fn is not None:
                object.__setattr__(self, k, v.detach())
        for name in ("g_taps_A", "g_taps_B", "d_taps_A", "d_taps_B"):
            taps = getattr(self, name, None)
            if taps:
                setattr(self, name, [(f.detach(), c) for f, c in taps])
        for net in (self.netG_A, self.netG_B, self.netD_A, self.netD_B):
            net.taps = []

    def optimize_parameters(self):
        ops.zero_pool.reset()
        if self.opt.online_distillation:
            T = self.teacher_model
            T._adopt_input(self)
            T.optimize_parameters()
            for k in "AB":
                feats = [(f.detach(), c) for f, c in (getattr(T, "g_taps_" + k) + getattr(T, "d_taps_" + k))]
                setattr(self, "target_distillation_%s_features" % k, [f for f, _ in feats])
                setattr(self, "target_grams_" + k, [ops.gram_matrix(f, c) for f, c in feats])
        self.forward()
        self.set_requires_grad([self.netD_A, self.netD_B], False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        self.optimizer_G.step()
        self.set_requires_grad([self.netD_A, self.netD_B], True)
        self.set_netD_arch_grad(False)
        self.optimizer_D.zero_grad()
        self.backward_D_A()
        self.backward_D_B()
        self.optimizer_D.step()
        self._release_graphs()

    def optimizer_netD_arch(self):
        ops.zero_pool.reset()
        self.forward()
        self.teacher_model._adopt_input(self)
        self.teacher_model.forward()
        self.set_requires_grad([self.netD_A, self.netD_B], True)
        self.set_netD_weight_grad(False)
        self.optimizer_arch.zero_grad()
        self.backward_D_arch()
        self.optimizer_arch.step()
        self._release_graphs()
        self.teacher_model._release_graphs()

    # ------------------------------------------------------------------ bookkeeping (reference surface)
    def print_sparse_info(self, logger):
        if not self.opt.darts_discriminator:
            return
        for tag, net in (("netD_A", self.netD_A), ("netD_B", self.netD_B)):
            for i, mask in enumerate(net.get_current_masks()):
                logger.info("%s gate %d sparsity ratio: %.2f" % (tag, i, float((mask == 0.0).sum()) / mask.numel()))

    def adaptive_ema_beta(self, epoch):
        self.opt.ema_beta = 1.0 - epoch / (self.opt.n_epochs + self.opt.n_epochs_decay)

    def update_learning_rate(self, epoch):
        for s in self.schedulers:
            s.step()
        self.adaptive_ema_beta(epoch)
        print("learning rate = %.7f" % self.optimizers[0].param_groups[0]["lr"])

    def set_requires_grad(self, nets, requires_grad=False):
        for net in (nets if isinstance(nets, list) else [nets]):
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
            ops.call("gcc_clamp_f32", self.arena_A.P.data_ptr(), self.arena_A.numel, 0.0, 1.0, ops._st())

    def model_train(self):
        for n in (self.netG_A, self.netG_B, self.netD_A, self.netD_B):
            n.train()

    def model_eval(self):
        for n in (self.netG_A, self.netG_B, self.netD_A, self.netD_B):
            n.eval()

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names)

    def get_current_losses(self):
        return OrderedDict((n, float(getattr(self, "loss_" + n).detach())) for n in self.loss_names)

    def init_distillation(self):
        if self.distill:
            if self.opt.lambda_content > 0.0:
                self.loss_names += ["content_A", "content_B"]
            if self.opt.lambda_gram > 0.0:
                self.loss_names += ["gram_A", "gram_B"]
            if self.opt.lambda_L1 > 0.0:
                self.loss_names += ["L1_A", "L1_B"]
            self.visual_names += ["Tfake_A", "Tfake_B"]

    def get_distillation_features(self, AorB="A"):
        taps = getattr(self, "g_taps_" + AorB) + getattr(self, "d_taps_" + AorB, [])
        return [ops.to_nchw(f.detach(), c) for f, c in taps]

    def gram(self, x):
        return ops.gram_matrix(ops.to_nhwc(x), x.shape[1])

    def get_cfg(self):
        return self.cfg_AtoB, self.cfg_BtoA

    def save_models(self, epoch, save_dir, fid=None, isbest=False, direction="AtoB"):
        os.makedirs(save_dir, exist_ok=True)
        ckpt = {"G_A": self.netG_A.state_dict(), "G_B": self.netG_B.state_dict(), "D_A": self.netD_A.state_dict(),
                "D_B": self.netD_B.state_dict(), "epoch": epoch, "cfg": (self.cfg_AtoB, self.cfg_BtoA), "fid": fid}
        torch.save(self._ckpt_add_resume(ckpt),
                   os.path.join(save_dir, "model_best_%s.pth" % direction if isbest else "model_%d.pth" % epoch))

    def load_models(self, load_path, load_discriminator=True, resume=None):
        """`resume`: None = continue training exactly (optimizer moments, counters, EMA, teacher) when the file carries
        the `gcc_b200` entry and the discriminator is loaded too; False = weights only, as the reference."""
        ckpt = torch.load(load_path, map_location=self.device, weights_only=False)
        drop = lambda sd: {k: v for k, v in sd.items() if not (k.endswith("total_ops") or k.endswith("total_params"))}
        self.netG_A.load_state_dict(drop(ckpt["G_A"]))
        self.netG_B.load_state_dict(drop(ckpt["G_B"]))
        if load_discriminator:
            self.netD_A.load_state_dict(drop(ckpt["D_A"]))
            self.netD_B.load_state_dict(drop(ckpt["D_B"]))
        self._ckpt_load_resume(ckpt, load_discriminator, resume)
        print("loading the model from %s" % load_path)

    # ------------------------------------------------------------------ pruning (index selection)
    def max_min_conv_norm(self, netG):
        from . import prune as P
        return P.cyclegan_max_min_conv_norm(netG.state_dict())

    def get_prunenet_cfg(self, netG, threshold):
        from . import prune as P
        return P.cyclegan_prunenet_cfg(netG.state_dict(), threshold)

    def resnet_prune(self, threshold_AtoB, threshold_BtoA):
        return MobileCycleGANModel(self.opt, cfg_AtoB=self.get_prunenet_cfg(self.netG_A, threshold_AtoB),
                                   cfg_BtoA=self.get_prunenet_cfg(self.netG_B, threshold_BtoA))

    def prune(self, threshold, lottery_path=None):
        return self.resnet_prune(threshold, lottery_path)


def build_cycle_teacher(model, opt):
    """train.py:92-105 for --model cyclegan."""
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    topt.generator_only = False
    teacher = MobileCycleGANModel(topt)
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    return teacher
