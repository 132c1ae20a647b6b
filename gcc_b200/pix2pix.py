"""Pix2PixModel: the reference's model wrapper surface (/root/reference/models/Pix2Pix.py:350-952)
driving the B200 kernels.  Method names, argument meaning, attribute names (``netG``, ``netD``,
``teacher_model``, ``fake_B`` ...) and error behaviour follow the reference so that its ``train.py`` /
``test.py`` loops drive this class unchanged:

    model.set_input(data); model.optimize_parameters()
    model.set_input(val);  model.clipping_mask_alpha(); model.optimizer_netD_arch()

Differences that are deliberate: activations live on the device as NHWC bf16, the
three Adam optimizers are fused flat-arena steps, and when ``torch.distributed`` is initialised the
gradient arenas are all-reduced over NCCL before every optimizer step (data parallel).
"""
import copy
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops
from .arena import ParamArena, rp8
from .base import GccModelMixin, capturing, dist_on
from .nets import (ConvLayer, MaskNLayerDiscriminator, MobileResnetGenerator, NLayerDiscriminator, UnetGenertor,
                   unet_level_prefixes)
from .ops import GAN_MODES


class _ArenaOptimizer:
    """torch.optim-like facade over a ParamArena (zero_grad / step / param_groups[0]['lr']).

    ``step(defer=True)`` (data parallel only): the gradient all-reduce is launched asynchronously and the Adam update
    waits until ``finish()`` -- called before the parameters are next read.  The teacher generator's arena (60 % of the
    exchanged bytes) uses it: its new weights are first needed by the gate step's teacher forward, so the exchange
    hides behind the student's discriminator / generator steps."""

    def __init__(self, arena, lr, betas):
        self.arena = arena
        arena.lr, arena.betas = lr, betas
        if arena.finalized:
            arena._write_hyper()
        self.param_groups = [{"lr": lr, "initial_lr": lr, "betas": betas}]
        self._pending = None

    def zero_grad(self):
        self.finish()
        self.arena.zero_grad()

    def step(self, defer=False):
        self.finish()
        if not capturing():
            self.arena.set_lr(self.param_groups[0]["lr"])
        if defer and _dist_on():
            self._pending = _allreduce_grads(self.arena, asynchronous=True)
            return
        _allreduce_grads(self.arena)
        self.arena.step()

    def finish(self):
        """Complete a deferred step: wait for the exchange, run the fused Adam update."""
        if self._pending is not None:
            pending, self._pending = self._pending, None
            if pending == "segment":
                _graph_segmenter(self.arena, "wait")       # capture: cut; run() waits for the exchange here
            else:
                pending.wait()
            self.arena.step()


_dist_on = dist_on


# Set by gcc_b200.graph.GraphedIteration while it captures a data-parallel iteration: the gradient exchange is a
# cut point between two CUDA graphs (the collective itself is launched eagerly between the replays).
# Called as _graph_segmenter(arena, action) with action in {"allreduce", "allreduce_async", "wait"}.
_graph_segmenter = None


def _allreduce_grads(arena, asynchronous=False):
    """Data parallel: average the flat gradient arena over ranks (NCCL over NVLink) before the step.  Asynchronous:
    returns a handle (``.wait()`` before the gradients are consumed); the collective runs on NCCL's own stream and
    overlaps whatever is queued on the compute stream meanwhile."""
    if not _dist_on():
        return None
    if _graph_segmenter is not None:
        _graph_segmenter(arena, "allreduce_async" if asynchronous else "allreduce")
        return "segment" if asynchronous else None
    dist = torch.distributed
    if dist.get_backend() == "nccl":
        work = dist.all_reduce(arena.G, op=dist.ReduceOp.AVG, async_op=asynchronous)
    else:
        work = dist.all_reduce(arena.G, op=dist.ReduceOp.SUM, async_op=False)
        arena.G.mul_(1.0 / dist.get_world_size())
        if asynchronous:
            work = _DoneWork()
    return work if asynchronous else None


class _DoneWork:
    def wait(self):
        return True


class _LambdaLR:
    def __init__(self, optimizer, fn):
        self.optimizer, self.fn, self.epoch = optimizer, fn, 0
        self.base = optimizer.param_groups[0]["lr"]

    def step(self):
        self.epoch += 1
        self.optimizer.param_groups[0]["lr"] = self.base * self.fn(self.epoch)


class _PlateauLR:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', factor=0.2, threshold=0.01 (relative), patience=5),
    the reference's 'plateau' policy (utils/util.py:299-300).  ``step(metric)`` needs the monitored value; the
    reference's update_learning_rate calls ``scheduler.step()`` without one, which raises TypeError upstream too."""

    def __init__(self, optimizer, factor=0.2, threshold=0.01, patience=5, eps=1e-8):
        self.optimizer, self.factor, self.threshold, self.patience, self.eps = optimizer, factor, threshold, patience, eps
        self.best, self.bad, self.epoch = float("inf"), 0, 0

    def step(self, *metrics):
        if not metrics:
            raise TypeError("step() missing 1 required positional argument: 'metrics'")
        cur = float(metrics[0])
        self.epoch += 1
        if cur < self.best * (1.0 - self.threshold):
            self.best, self.bad = cur, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            g = self.optimizer.param_groups[0]
            new = g["lr"] * self.factor
            if g["lr"] - new > self.eps:
                g["lr"] = new
            self.bad = 0


def get_scheduler(optimizer, opt):
    """utils/util.py:288-303 (linear / step / plateau / cosine)."""
    import math
    if opt.lr_policy == "plateau":
        return _PlateauLR(optimizer)
    if opt.lr_policy == "linear":
        return _LambdaLR(optimizer, lambda e: 1.0 - max(0, e + opt.epoch_count - opt.n_epochs) /
                         float(opt.n_epochs_decay + 1))
    if opt.lr_policy == "step":
        return _LambdaLR(optimizer, lambda e: 0.1 ** (e // opt.lr_decay_iters))
    if opt.lr_policy == "cosine":
        return _LambdaLR(optimizer, lambda e: (1 + math.cos(math.pi * e / opt.n_epochs)) / 2)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


class Pix2PixModel(GccModelMixin, nn.Module):

    def __init__(self, opt, filter_cfgs=None, channel_cfgs=None):
        super().__init__()
        self.opt = opt
        if len(opt.gpu_ids) == 0:
            raise RuntimeError("gcc_b200.Pix2PixModel needs a CUDA device (gpu_ids): there is no CPU path")
        self.device = torch.device("cuda:%d" % opt.gpu_ids[0])
        ops.call("gcc_check_device")
        self.filter_cfgs, self.channel_cfgs = filter_cfgs, channel_cfgs
        self.loss_names = ["G_GAN", "G_L1", "D_real", "D_fake"]
        self.visual_names = ["real_A", "fake_B", "real_B"]
        self.current_D_arch_diff_loss = 0.0
        self.teacher_model = None
        dev = self.device
        self._base_init()
        self.distill = bool(opt.online_distillation or getattr(opt, "normal_distillation", False))

        # ---- generator + optimizer_G arena (transform convs first, as in Pix2Pix.py:403-415)
        self.arena_G = ParamArena(dev)
        self.transform_convs = []
        if opt.backbone == "resnet":
            self.generator_extract_layers = ["model.9", "model.12", "model.15", "model.18"]
        else:
            self.generator_extract_layers = ["model.model.1.model.2", "model.model.1.model.3.model.3.model.2",
                                             "model.model.1.model.3.model.3.model.4", "model.model.1.model.4"]
        self.discriminator_extract_layers = ["model.4", "model.12"] if opt.darts_discriminator else \
            ["model.3", "model.9"]
        if self.distill:
            if opt.backbone == "resnet":
                t_ch = [opt.teacher_ngf * 4] * 4
                s_ch = [opt.ngf * 4] * 4 if filter_cfgs is None else [filter_cfgs[2]] * 4
            else:
                t_ch = [opt.teacher_ngf * 2, opt.teacher_ngf * 8, opt.teacher_ngf * 16, opt.teacher_ngf * 4]
                s_ch = [opt.ngf * 2, opt.ngf * 8, opt.ngf * 16, opt.ngf * 4] if channel_cfgs is None else \
                    [channel_cfgs[1], channel_cfgs[3], channel_cfgs[-4], channel_cfgs[-2]]
            for i in range(4):
                self.transform_convs.append(ConvLayer(self.arena_G, "transform.%d" % i, "conv", s_ch[i], t_ch[i], 1, 1, 0))
        if opt.backbone == "resnet":
            self.netG = MobileResnetGenerator(3, 3, ngf=opt.ngf, cfg=filter_cfgs, arena=self.arena_G, device=dev, opt=opt)
        else:
            self.netG = UnetGenertor(3, 3, opt.num_downs, ngf=opt.ngf, use_dropout=not opt.no_dropout,
                                     filter_cfgs=filter_cfgs, channel_cfgs=channel_cfgs, arena=self.arena_G, device=dev)
        self.arena_G.finalize()
        self.netG.finalize()
        for t in self.transform_convs:
            t.bind()
        self.optimizer_G = _ArenaOptimizer(self.arena_G, opt.lr, (0.5, 0.999))

        # ---- discriminator (+ gate arena)
        self.arena_D = ParamArena(dev)
        if opt.darts_discriminator:
            self.loss_names += ["D_arch_diff", "D_arch", "teacher_D_arch_diff"]
            self.arena_A = ParamArena(dev)
            self.netD = MaskNLayerDiscriminator(input_nc=6, ndf=opt.ndf, threshold=opt.threshold, arena=self.arena_D,
                                                gate_arena=self.arena_A, device=dev)
            self.arena_D.finalize()
            self.arena_A.finalize()
            self.netD.finalize()
            self.optimizer_D = _ArenaOptimizer(self.arena_D, opt.lr, (0.5, 0.999))
            self.optimizer_arch = _ArenaOptimizer(self.arena_A, opt.arch_lr, (0.9, 0.999))
        else:
            self.arena_A = None
            self.netD = NLayerDiscriminator(input_nc=6, ndf=opt.ndf, arena=self.arena_D, device=dev)
            self.arena_D.finalize()
            self.netD.finalize()
            self.optimizer_D = _ArenaOptimizer(self.arena_D, opt.lr, (0.5, 0.999))
        self.init_net()

        self.gan_mode = GAN_MODES.get(opt.gan_mode)
        if self.gan_mode is None:
            raise NotImplementedError("gan mode %s not implemented" % opt.gan_mode)
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.schedulers = [get_scheduler(o, opt) for o in self.optimizers]
        if opt.darts_discriminator and getattr(opt, "arch_lr_step", False):
            arch_opt = copy.deepcopy(opt)
            arch_opt.lr_policy = "step"
            arch_opt.lr_decay_iters = opt.n_epochs - 1
            self.arch_scheduler = get_scheduler(self.optimizer_arch, arch_opt)
            self.schedulers.append(self.arch_scheduler)
        self.total_generator_features = {}
        self.total_discriminator_features = {}
        self.broadcast_parameters()

    def _gcc_arenas(self):
        return {k: a for k, a in (("G", self.arena_G), ("D", self.arena_D), ("A", self.arena_A)) if a is not None}

    def _gcc_optimizers(self):
        d = {"G": self.optimizer_G, "D": self.optimizer_D}
        if self.arena_A is not None:
            d["arch"] = self.optimizer_arch
        return d

    def _gcc_nets(self):
        return {"netG": self.netG, "netD": self.netD}

    # ------------------------------------------------------------------ init (util.init_weights)
    def init_net(self):
        """N(0, .02) conv weights, zero conv biases, BN gamma N(1, .02), BN beta N(0, 1)
        (utils/util.py:261-286); transform convs keep nn.Conv2d's default kaiming-uniform; alpha = 1."""
        import math
        with torch.no_grad():
            for arena in (self.arena_G, self.arena_D):
                for (name, shape, kind) in arena.specs:
                    p = arena.params[name]
                    if name.startswith("transform."):
                        bound = 1.0 / math.sqrt(shape[1])
                        p.copy_(torch.empty(shape, device=self.device).uniform_(-bound, bound))
                    elif len(shape) == 4:
                        p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 0.02))
                    elif name.endswith(".weight"):
                        p.copy_(torch.empty(shape, device=self.device).normal_(1.0, 0.02))
                    elif name.endswith(".bias"):
                        is_bn = (name[:-4] + "weight") in arena.params and len(arena.params[name[:-4] + "weight"].shape) == 1
                        if is_bn:
                            p.copy_(torch.empty(shape, device=self.device).normal_(0.0, 1.0))
                        else:
                            p.zero_()
                arena.mark_dirty()

    def sync_weights(self):
        """Call after mutating parameters from outside (tests, checkpoint surgery)."""
        for a in (self.arena_G, self.arena_D, self.arena_A):
            if a is not None:
                a.mark_dirty()

    # ------------------------------------------------------------------ inputs / forward
    def set_input(self, input):
        """Pix2Pix.py:453-458.  The host->device copies and NHWC bf16 conversions of a batch are done once per phase:
        the teacher's ``set_input(self.input)`` inside optimize_parameters / optimizer_netD_arch (Pix2Pix.py:568,587)
        takes the student's converted tensors (``_adopt_input``) instead of converting the same batch again.  Nothing
        is written into the caller's dict and nothing is cached across calls."""
        self.input = input
        AtoB = self.opt.direction == "AtoB"
        ka, kb = ("A", "B") if AtoB else ("B", "A")
        self.image_paths = [input.get(ka + "_paths"), input.get(kb + "_paths")]
        A = input[ka].to(self.device, non_blocking=True)
        B = input[kb].to(self.device, non_blocking=True)
        a_nhwc, b_nhwc = ops.to_nhwc(A), ops.to_nhwc(B)
        real_AB = ops.CatFn.apply(a_nhwc, b_nhwc, 3, 3)
        self._A_nchw, self._B_nchw = A, B
        self.real_A_nhwc, self.real_B_nhwc, self.real_AB = a_nhwc, b_nhwc, real_AB

    def _adopt_input(self, other):
        """Teacher side of ``T.set_input(self.input)``: same batch, same direction -> share the device tensors."""
        if other.opt.direction != self.opt.direction or other.device != self.device:
            return self.set_input(other.input)
        self.input, self.image_paths = other.input, other.image_paths
        self._A_nchw, self._B_nchw = other._A_nchw, other._B_nchw
        self.real_A_nhwc, self.real_B_nhwc, self.real_AB = other.real_A_nhwc, other.real_B_nhwc, other.real_AB

    @property
    def real_A(self):
        return self._A_nchw.float()      # (a prefetched batch arrives as bf16)

    @property
    def real_B(self):
        return self._B_nchw.float()

    @property
    def fake_B(self):
        """NCHW fp32 view of the generator output (the reference's tensor convention)."""
        return ops.to_nchw(self.fake_B_nhwc.detach(), 3)

    def forward(self):
        self.optimizer_G.finish()          # (a deferred data-parallel generator step completes before its weights are read)
        self.fake_B_nhwc = self.netG(self.real_A_nhwc)
        self.g_taps = list(self.netG.taps)

    def _D(self, net, ab):
        pred = net(ab)
        return pred, list(net.taps)

    def _fake_AB(self, fake, detach):
        f = fake.detach() if detach else fake
        return ops.CatFn.apply(self.real_A_nhwc, f, 3, 3)

    def _gan(self, pred, kind):
        return ops.GanLossFn.apply(pred, 1, self.gan_mode, kind)

    # ------------------------------------------------------------------ losses / steps
    def backward_D(self):
        pred_fake, _ = self._D(self.netD, self._fake_AB(self.fake_B_nhwc, True))
        self.loss_D_fake = self._gan(pred_fake, 1)
        pred_real, self.d_taps = self._D(self.netD, self.real_AB)
        self.loss_D_real = self._gan(pred_real, 0)
        self.loss_D = (self.loss_D_fake + self.loss_D_real) * 0.5
        self.loss_D.backward()

    def get_D_arch_diff(self, isTeacher=False):
        if isTeacher:
            self.set_requires_grad(self.netD, False)
        pred_fake, _ = self._D(self.netD, self._fake_AB(self.fake_B_nhwc, True))
        self.loss_D_arch_fake = self._gan(pred_fake, 1)
        self.loss_D_arch_fake_real = self._gan(pred_fake, 2)
        pred_real, _ = self._D(self.netD, self.real_AB)
        self.loss_D_arch_real = self._gan(pred_real, 0)
        diff = (self.loss_D_arch_fake_real - self.loss_D_arch_fake).abs()
        if isTeacher:
            # EMA state (Pix2Pix.py:503-508): persistent device scalar, beta read from device memory (base.py); the
            # teacher D is frozen here, so the state is graph-free as upstream.
            self.current_D_arch_diff_loss = self._ema_update("D", diff)
        else:
            self.current_D_arch_diff_loss = diff
        return self.current_D_arch_diff_loss, torch.sign(self.loss_D_arch_fake_real - self.loss_D_arch_fake)

    def backward_D_arch(self):
        self.loss_teacher_D_arch_diff, _ = self.teacher_model.get_D_arch_diff(isTeacher=True)
        self.loss_D_arch_diff, _ = self.get_D_arch_diff(isTeacher=False)
        self.loss_D_arch = (self.loss_D_arch_diff - self.loss_teacher_D_arch_diff.detach()).abs()
        self.loss_D_arch = self.loss_D_arch + (self.loss_D_arch_real + self.loss_D_arch_fake) * 0.5
        self.loss_D_arch.backward()

    def backward_G(self):
        o = self.opt
        pred_fake, self.d_taps = self._D(self.netD, self._fake_AB(self.fake_B_nhwc, False))
        self.loss_G_GAN = self._gan(pred_fake, 2)
        self.loss_G_L1 = ops.DiffLossFn.apply(self.fake_B_nhwc, self.real_B_nhwc, 3, 0) * o.lambda_L1
        self.loss_G = self.loss_G_GAN + self.loss_G_L1
        if self.distill:
            T = self.teacher_model
            self.Tfake_B_nhwc = T.fake_B_nhwc
            feats = list(self.g_taps)
            _, t_d_taps = self._D(T.netD, self._fake_AB(self.fake_B_nhwc, False))  # teacher D on the student fake
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
            self.loss_G = self.loss_G + self.loss_gram + self.loss_content
        self.loss_G.backward()
        self.L1_sparsity()

    def L1_sparsity(self):
        o = self.opt
        st = ops._st()
        if o.lambda_weight > 0.0:
            for (name, shape, kind) in self.arena_G.specs:
                if len(shape) == 4 and not name.startswith("transform."):
                    off, n = self.arena_G.offsets[[s[0] for s in self.arena_G.specs].index(name)]
                    ops.call("gcc_l1_sparsity_f32", self.arena_G.P[off:].data_ptr(), self.arena_G.G[off:].data_ptr(), n,
                             o.lambda_weight, st)
        elif o.lambda_scale > 0.0:
            names = [s[0] for s in self.arena_G.specs]
            for (name, shape, kind) in self.arena_G.specs:
                if len(shape) == 1 and name.endswith(".weight"):
                    off, n = self.arena_G.offsets[names.index(name)]
                    ops.call("gcc_l1_sparsity_f32", self.arena_G.P[off:].data_ptr(), self.arena_G.G[off:].data_ptr(), n,
                             o.lambda_scale, st)

    def _release_graphs(self):
        """Detach every tensor attribute that still references an autograd graph.  The reference keeps them
        attached but never uses them again; releasing them lets the AccumulateGrad nodes of the parameters expire
        between iterations, which CUDA-graph capture needs (a stale node pins the stream it was created on)."""
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
        self.optimizer_G.step(defer=getattr(self, "_defer_G_step", False))
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
        for s in self.schedulers:
            s.step()
        self.adaptive_ema_beta(epoch)
        lr = self.optimizers[0].param_groups[0]["lr"]
        print("learning rate = %.7f\tema beta = %.7f" % (lr, self.opt.ema_beta))

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
            ops.call("gcc_clamp_f32", self.arena_A.P.data_ptr(), self.arena_A.numel, 0.0, 1.0, ops._st())

    def model_train(self):
        self.netG.train()
        self.netD.train()

    def model_eval(self):
        self.netG.eval()
        self.netD.eval()

    def get_current_visuals(self):
        ret = OrderedDict()
        for name in self.visual_names:
            if name == "Tfake_B":
                ret[name] = ops.to_nchw(self.Tfake_B_nhwc.detach(), 3)
            else:
                ret[name] = getattr(self, name)
        return ret

    def get_current_losses(self):
        ret = OrderedDict()
        for name in self.loss_names:
            ret[name] = float(getattr(self, "loss_" + name).detach())  # AttributeError before the first arch step, as upstream
        return ret

    def init_distillation(self):
        if self.distill:
            if self.opt.lambda_content > 0.0:
                self.loss_names.append("content")
            if self.opt.lambda_gram > 0.0:
                self.loss_names.append("gram")
            self.visual_names.append("Tfake_B")

    def get_distillation_features(self):
        """NCHW fp32 copies of the taps, in the reference's order (generator taps then discriminator taps)."""
        return [ops.to_nchw(f.detach(), c) for f, c in (self.g_taps + getattr(self, "d_taps", []))]

    def gram(self, x):
        b, c, h, w = x.shape
        return ops.gram_matrix(ops.to_nhwc(x), c)

    def get_cfg(self):
        return self.filter_cfgs, self.channel_cfgs

    def save_models(self, epoch, save_dir, fid=None, isbest=False, direction="AtoB"):
        self.finish_pending_steps()
        os.makedirs(save_dir, exist_ok=True)
        ckpt = {"G": self.netG.state_dict(), "D": self.netD.state_dict(), "epoch": epoch,
                "cfg": (self.filter_cfgs, self.channel_cfgs), "fid": fid}
        path = os.path.join(save_dir, "model_best_%s.pth" % direction if isbest else "model_%d.pth" % epoch)
        torch.save(self._ckpt_add_resume(ckpt), path)

    def load_models(self, load_path, load_discriminator=True, resume=None):
        """`resume`: None = continue training exactly (optimizer moments, counters, EMA, teacher) when the file carries
        the `gcc_b200` entry and the discriminator is loaded too; False = weights only, as the reference."""
        ckpt = torch.load(load_path, map_location=self.device, weights_only=False)
        self.netG.load_state_dict(ckpt["G"])
        if load_discriminator:
            self.netD.load_state_dict(ckpt["D"])
        self._ckpt_load_resume(ckpt, load_discriminator, resume)
        print("loading the model from %s" % load_path)
        return ckpt["fid"], float("inf")

    # ------------------------------------------------------------------ pruning (index selection)
    def prune(self, threshold, lottery_path=None):
        from . import prune as P
        if self.opt.backbone == "resnet":
            cfgs = (P.resnet_prune_cfg(self.netG.state_dict(), threshold, P.convt_names(self.netG)), None)
        elif self.opt.scale_prune:
            cfgs = P.unet_scale_prune_cfg(self.netG.state_dict(), self.opt.ngf, threshold)
        elif self.opt.norm_prune:
            cfgs = P.unet_norm_prune_cfg(self.netG.state_dict(), self.opt.ngf, threshold)
        else:
            raise NotImplementedError("only scale and norm pruning are supported!!!")
        return Pix2PixModel(self.opt, filter_cfgs=cfgs[0], channel_cfgs=cfgs[1])

    def max_min_bn_scale(self):
        from . import prune as P
        return P.unet_max_min_bn_scale(self.netG.state_dict())

    def max_min_conv_norm(self):
        from . import prune as P
        return P.max_min_conv_norm(self.netG.state_dict(), self.opt.backbone, P.convt_names(self.netG))


def build_teacher(model, opt):
    """train.py:92-105: teacher = same class, teacher widths, plain D, no distillation."""
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    topt.generator_only = False
    teacher = Pix2PixModel(topt)
    teacher._defer_G_step = True           # data parallel: its generator exchange overlaps the student's steps
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    return teacher
