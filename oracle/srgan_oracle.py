"""CPU oracle for the GCC SRGAN cooperative-compression step.

TEST INFRASTRUCTURE ONLY (same rules as gcc_oracle.py): used by ``tests/`` as the checker, never imported by
``gcc_b200/``.  Functional fp32 torch-CPU restatement of

  SRResNet generator ......... /root/reference/models/SRGAN.py:15-197
  discriminators ............. /root/reference/models/SRGAN.py:199-295
  truncated VGG19 ............ /root/reference/models/GANLoss.py:95-144 (torchvision vgg19.features[:36])
  convert_image .............. /root/reference/data/sr_dataset.py:15-64
  training step .............. /root/reference/models/SRGAN.py:388-528
  distillation taps / gram ... /root/reference/models/SRGAN.py:668-705

Pinned by ``tests/golden/srgan_tiny.pt`` (oracle/make_golden.py runs the unmodified reference; the only shim that
touches values is ``vgg19(weights=None)``: there is no network for the pretrained weights, so the perceptual
branch is checked with name-seeded random VGG weights on both sides).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .gcc_oracle import (Adam, Opt, _make_params, _trainable, batch_norm, det_normal, det_uniform, gan_loss, gate_mask,
                         gram)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512]
VGG_CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28, 30, 32, 34]   # indices in vgg19.features


class SROpt(Opt):
    """options.parse() for --model srgan (options.py:192-205): vanilla GAN loss, lr 1e-4, default Adam betas."""

    def __init__(self, **kw):
        super().__init__(gan_mode="vanilla", lr=1e-4, ngf=24, ndf=32, teacher_ngf=64, teacher_ndf=64,
                         lambda_content=1e-3, lambda_gram=1e-1, lambda_L1=0.0)
        self.lambda_SR_adversarial, self.lambda_SR_content, self.lambda_SR_perceptual = 1e-3, 0.0, 1.0
        self.n_blocks = 16
        for k, v in kw.items():
            setattr(self, k, v)


def convert_to_imagenet(x):
    """convert_image(x, '[-1, 1]', 'imagenet-norm') (data/sr_dataset.py:33-58)."""
    mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=x.dtype).view(1, 3, 1, 1)
    return ((x + 1.0) / 2.0 - mean) / std


# ---------------------------------------------------------------------------------------------- generator
def sr_generator_shapes(ngf, filter_cfgs=None, n_blocks=16):
    s = OrderedDict()

    def conv(name, cin, cout, k):
        s[name + ".weight"] = (cout, cin, k, k)
        s[name + ".bias"] = (cout,)

    def bn(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)
        s[name + ".running_mean"] = (c,)
        s[name + ".running_var"] = (c,)
        s[name + ".num_batches_tracked"] = ()

    conv("conv_block1.conv_block.0", 3, ngf, 9)
    s["conv_block1.conv_block.1.weight"] = (1,)                    # PReLU
    for i in range(n_blocks):
        inner = ngf if filter_cfgs is None else filter_cfgs[i]
        p = "residual_blocks.%d." % i
        conv(p + "conv_block1.conv_block.0", ngf, inner, 3)
        bn(p + "conv_block1.conv_block.1", inner)
        s[p + "conv_block1.conv_block.2.weight"] = (1,)            # PReLU
        conv(p + "conv_block2.conv_block.0", inner, ngf, 3)
        bn(p + "conv_block2.conv_block.1", ngf)
    conv("conv_block2.conv_block.0", ngf, ngf, 3)
    bn("conv_block2.conv_block.1", ngf)
    for i in range(2):
        conv("subpixel_convolutional_blocks.%d.conv" % i, ngf, ngf * 4, 3)
        s["subpixel_convolutional_blocks.%d.prelu.weight" % i] = (1,)
    conv("conv_block3.conv_block.0", ngf, 3, 9)
    return s


def init_sr_params(P, prefix):
    """Name-seeded stand-in for util.init_weights + nn.PReLU's 0.25 (overwrites what _make_params chose for the
    1-element PReLU weights, which init_like_reference would treat as a BatchNorm scale)."""
    with torch.no_grad():
        for k, v in P.items():
            if v.dim() == 1 and v.numel() == 1 and k.endswith(".weight") and (k[:-6] + "running_mean") not in P \
                    and "fc1" not in k:
                v.fill_(0.25)
    return P


def prelu(x, a):
    return torch.where(x > 0, x, a * x)


def sr_generator_forward(P, x, n_blocks=16, train=True):
    def conv(name, t, k):
        return F.conv2d(t, P[name + ".weight"], P[name + ".bias"], padding=k // 2)

    out = prelu(conv("conv_block1.conv_block.0", x, 9), P["conv_block1.conv_block.1.weight"])
    residual = out
    taps = []
    for i in range(n_blocks):
        p = "residual_blocks.%d." % i
        r = out
        t = conv(p + "conv_block1.conv_block.0", out, 3)
        t = prelu(batch_norm(t, P, p + "conv_block1.conv_block.1", train), P[p + "conv_block1.conv_block.2.weight"])
        t = batch_norm(conv(p + "conv_block2.conv_block.0", t, 3), P, p + "conv_block2.conv_block.1", train)
        out = t + r
        if i in (3, 7, 11, 15):
            taps.append(out)
    out = batch_norm(conv("conv_block2.conv_block.0", out, 3), P, "conv_block2.conv_block.1", train) + residual
    for i in range(2):
        out = conv("subpixel_convolutional_blocks.%d.conv" % i, out, 3)
        out = prelu(F.pixel_shuffle(out, 2), P["subpixel_convolutional_blocks.%d.prelu.weight" % i])
    return torch.tanh(conv("conv_block3.conv_block.0", out, 9)), taps


# ------------------------------------------------------------------------------------------ discriminators
def sr_disc_channels(ndf, n_blocks=4):
    chans, cin = [], 3
    for i in range(n_blocks):
        cout = (ndf if i == 0 else cin * 2) if i % 2 == 0 else cin
        chans.append((cin, cout, 1 if i % 2 == 0 else 2, i != 0))
        cin = cout
    return chans


def sr_disc_shapes(ndf, gated=False, n_blocks=4):
    s = OrderedDict()
    for i, (cin, cout, stride, has_bn) in enumerate(sr_disc_channels(ndf, n_blocks)):
        p = "conv_blocks.%d.conv_block." % i
        s[p + "0.weight"] = (cout, cin, 3, 3)
        s[p + "0.bias"] = (cout,)
        j = 1
        if has_bn:
            s[p + "1.weight"] = (cout,)
            s[p + "1.bias"] = (cout,)
            s[p + "1.running_mean"] = (cout,)
            s[p + "1.running_var"] = (cout,)
            s[p + "1.num_batches_tracked"] = ()
            j = 2
        if gated:
            s[p + "%d.alpha" % j] = (cout,)
    s["fc1.weight"] = (1, cout)
    s["fc1.bias"] = (1,)
    return s


def sr_disc_forward(P, x, ndf, gated=False, threshold=0.5, train=True, n_blocks=4):
    """conv -> [BN] -> [gate] -> LeakyReLU(0.2) blocks, global average pool, Linear -> (N, 1)
    (SRGAN.py:199-295; taps = outputs of conv_blocks.1 and conv_blocks.3)."""
    taps = []
    out = x
    for i, (cin, cout, stride, has_bn) in enumerate(sr_disc_channels(ndf, n_blocks)):
        p = "conv_blocks.%d.conv_block." % i
        out = F.conv2d(out, P[p + "0.weight"], P[p + "0.bias"], stride=stride, padding=1)
        j = 1
        if has_bn:
            out = batch_norm(out, P, p + "1", train)
            j = 2
        if gated:
            out = out * gate_mask(P[p + "%d.alpha" % j], threshold)[None, :, None, None]
        out = F.leaky_relu(out, 0.2)
        if i in (1, 3):
            taps.append(out)
    pooled = out.mean(dim=(2, 3))
    return pooled @ P["fc1.weight"].t() + P["fc1.bias"], taps


# ---------------------------------------------------------------------------------------------------- VGG
def vgg_shapes():
    s, cin, idx = OrderedDict(), 3, 0
    for v in VGG_CFG:
        if v == "M":
            idx += 1
            continue
        s["truncated_vgg19.%d.weight" % idx] = (v, cin, 3, 3)
        s["truncated_vgg19.%d.bias" % idx] = (v,)
        cin = v
        idx += 2
    return s


def make_vgg_params():
    """Name-seeded He-normal stand-in for the (unavailable) pretrained weights; shared by teacher and student."""
    P = OrderedDict()
    for k, shape in vgg_shapes().items():
        if k.endswith("weight"):
            P[k] = det_normal("vgg." + k, shape, 0.0, math.sqrt(2.0 / (shape[1] * 9)))
        else:
            P[k] = det_normal("vgg." + k, shape, 0.0, 0.05)
    return P


def vgg_forward(P, x):
    idx = 0
    for v in VGG_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            x = F.relu(F.conv2d(x, P["truncated_vgg19.%d.weight" % idx], P["truncated_vgg19.%d.bias" % idx], padding=1))
            idx += 2
    return x


# -------------------------------------------------------------------------------------------------- model
class SRGANOracle:
    """One SRGAN model (student or teacher): SRGAN.py:296-528."""

    def __init__(self, opt, tag, filter_cfgs=None):
        self.opt, self.tag, self.filter_cfgs = opt, tag, filter_cfgs
        self.G = init_sr_params(_make_params(sr_generator_shapes(opt.ngf, filter_cfgs, opt.n_blocks), tag + ".netG."),
                                tag + ".netG.")
        self.gated = bool(opt.darts_discriminator)
        self.D = _make_params(sr_disc_shapes(opt.ndf, self.gated), tag + ".netD.")
        self.vgg = make_vgg_params()
        self.transform = []
        prelu_names = [k for k in self.G if self.G[k].numel() == 1 and self.G[k].dim() == 1 and k.endswith(".weight")
                       and (k[:-6] + "running_mean") not in self.G]
        self.prelu_names = prelu_names
        if opt.online_distillation:
            for i in range(4):
                w = det_uniform("%s.transform.%d" % (tag, i), (opt.teacher_ngf, opt.ngf, 1, 1), 1.0 / math.sqrt(opt.ngf))
                w.requires_grad_(True)
                self.transform.append(w)
            # optimizer_G = transform convs + Conv / BatchNorm parameters: the PReLU slopes are left out (SRGAN.py:349-353)
            g_params = list(self.transform) + [v for k, v in self.G.items() if v.requires_grad and k not in prelu_names]
        else:
            g_params = _trainable(self.G)
        self.opt_G = Adam(g_params, opt.lr)
        self.d_weights = [v for k, v in self.D.items() if v.requires_grad and not k.endswith("alpha")]
        self.d_alphas = [v for k, v in self.D.items() if k.endswith("alpha")]
        self.opt_D = Adam(self.d_weights, opt.lr)
        self.opt_arch = Adam(self.d_alphas, opt.arch_lr) if self.gated else None
        self.teacher = None
        self.current_D_arch_diff_loss = 0.0
        self.train_mode = True

    def netG(self, x):
        return sr_generator_forward(self.G, x, self.opt.n_blocks, self.train_mode)

    def netD(self, x):
        return sr_disc_forward(self.D, x, self.opt.ndf, self.gated, self.opt.threshold, self.train_mode)

    def _set_grad(self, tensors, flag):
        for t in tensors:
            t.requires_grad_(flag)

    def set_input(self, lr, hr):
        self.real_lr, self.real_hr = lr, hr
        self._raw = (lr, hr)

    def forward(self):
        self.fake_hr, self.g_taps = self.netG(self.real_lr)

    def backward_D(self):
        m = self.opt.gan_mode
        pred_real, _ = self.netD(self.real_hr)
        pred_fake, self.d_taps = self.netD(self.fake_hr.detach())
        self.loss_D_real = gan_loss(m, pred_real, True)
        self.loss_D_fake = gan_loss(m, pred_fake, False)
        self.loss_D = self.loss_D_real + self.loss_D_fake
        self.loss_D.backward()

    def backward_G(self):
        o = self.opt
        self.loss_content = ((self.fake_hr - self.real_hr) ** 2).mean() * o.lambda_SR_content
        # the reference re-binds real_hr / fake_hr to their imagenet-normalised versions here (SRGAN.py:444-445);
        # everything after (D step included) sees the normalised tensors
        self.real_hr = convert_to_imagenet(self.real_hr)
        self.fake_hr = convert_to_imagenet(self.fake_hr)
        pred_fake, self.d_taps = self.netD(self.fake_hr)
        self.loss_G_GAN = gan_loss(o.gan_mode, pred_fake, True) * o.lambda_SR_adversarial
        fake_vgg = vgg_forward(self.vgg, self.fake_hr)
        real_vgg = vgg_forward(self.vgg, self.real_hr).detach()
        self.loss_perceptual = ((fake_vgg - real_vgg) ** 2).mean() * o.lambda_SR_perceptual
        self.loss_G = self.loss_content + self.loss_G_GAN + self.loss_perceptual
        if o.online_distillation:
            T = self.teacher
            feats = list(self.g_taps)
            _, t_d = T.netD(self.fake_hr)            # teacher D on the student fake: gradient reaches the student G
            feats += list(t_d)
            self.loss_content = 0.0
            self.loss_gram = 0.0
            for i, f in enumerate(feats):
                if i < 4:
                    f = F.conv2d(f, self.transform[i])
                tgt = self.target_features[i].detach()
                self.loss_gram = self.loss_gram + torch.sqrt(((gram(f) - gram(tgt)) ** 2).mean())
                self.loss_content = self.loss_content + torch.sqrt(((f - tgt) ** 2).mean())
            self.loss_gram = o.lambda_gram * self.loss_gram
            self.loss_content = o.lambda_content * self.loss_content
            self.loss_L1 = o.lambda_L1 * (self.fake_hr - T.fake_hr.detach()).abs().mean()
            self.loss_G = self.loss_G + self.loss_gram + self.loss_content + self.loss_L1
        self.loss_G.backward()

    def optimize_parameters(self):
        if self.opt.online_distillation:
            T = self.teacher
            T.set_input(*self._raw)
            T.optimize_parameters()
            self.target_features = [f.detach().clone() for f in (T.g_taps + T.d_taps)]
        self.forward()
        self._set_grad(self.d_weights + self.d_alphas, False)
        self.opt_G.zero_grad()
        self.backward_G()
        self.opt_G.step()
        self._set_grad(self.d_weights, True)
        self._set_grad(self.d_alphas, False)
        self.opt_D.zero_grad()
        self.backward_D()
        self.opt_D.step()

    def get_D_arch_diff(self, is_teacher):
        m = self.opt.gan_mode
        self.real_hr = convert_to_imagenet(self.real_hr)
        self.fake_hr = convert_to_imagenet(self.fake_hr)
        if is_teacher:
            self._set_grad(self.d_weights + self.d_alphas, False)
        pred_fake, _ = self.netD(self.fake_hr.detach())
        self.loss_D_arch_fake = gan_loss(m, pred_fake, False, True)
        self.loss_D_arch_fake_real = gan_loss(m, pred_fake, True, False)
        pred_real, _ = self.netD(self.real_hr)
        self.loss_D_arch_real = gan_loss(m, pred_real, True, True)
        diff = (self.loss_D_arch_fake_real - self.loss_D_arch_fake).abs()
        if is_teacher and not (isinstance(self.current_D_arch_diff_loss, float) and self.current_D_arch_diff_loss == 0.0):
            b = self.opt.ema_beta
            self.current_D_arch_diff_loss = b * diff + (1.0 - b) * self.current_D_arch_diff_loss
        else:
            self.current_D_arch_diff_loss = diff
        return self.current_D_arch_diff_loss

    def clipping_mask_alpha(self):
        with torch.no_grad():
            for a in self.d_alphas:
                a.clamp_(0, 1)

    def optimizer_netD_arch(self):
        T = self.teacher
        self.forward()
        T.set_input(*self._raw)
        T.forward()
        self._set_grad(self.d_weights + self.d_alphas, True)
        self._set_grad(self.d_weights, False)
        self.opt_arch.zero_grad()
        self.loss_teacher_D_arch_diff = T.get_D_arch_diff(True)
        self.loss_D_arch_diff = self.get_D_arch_diff(False)
        self.loss_D_arch = (self.loss_D_arch_diff - self.loss_teacher_D_arch_diff).abs()
        self.loss_D_arch = self.loss_D_arch + self.loss_D_arch_real + self.loss_D_arch_fake
        self.loss_D_arch.backward()
        self.opt_arch.step()

    def current_masks(self):
        return [((torch.sign(a.detach() - self.opt.threshold) + 1) / 2) for a in self.d_alphas]

    LOSSES = ["G_GAN", "D_real", "D_fake", "content", "perceptual", "D_arch_diff", "D_arch", "teacher_D_arch_diff", "gram",
              "L1"]

    def get_current_losses(self):
        return OrderedDict((n, float(getattr(self, "loss_" + n))) for n in self.LOSSES if hasattr(self, "loss_" + n))


def build_sr_pair(opt, filter_cfgs=None):
    """Student (gated D, distilled) + teacher (plain D, teacher widths) as train.py:84-105 builds them."""
    import copy
    student = SRGANOracle(opt, "S", filter_cfgs)
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    teacher = SRGANOracle(topt, "T")
    student.teacher = teacher
    return student, teacher


# ---------------------------------------------------------------------------- prune index selection (SRGAN.py:773-837)
def sr_norm_prune_cfg(G, threshold, n_blocks=16):
    """norm_prune: ||w||_1 over (1,2,3) > threshold of every residual_blocks.i.conv_block1 conv (the only prunable
    convs besides conv_block3, which the reference's list forgets to exclude: SRGAN.py:779-795)."""
    cfg = []
    for k, v in G.items():
        if v.dim() != 4:
            continue
        name = k[:-len(".weight")]
        unprunable = ["conv_block1.conv_block.0", "conv_block2.conv_block.0", "subpixel_convolutional_blocks.0.conv",
                      "subpixel_convolutional_blocks.1.conv"] + \
            ["residual_blocks.%d.conv_block2.conv_block.0" % i for i in range(n_blocks)]
        if name in unprunable:
            continue
        cfg.append(int((v.detach().abs().sum(dim=(1, 2, 3)) > threshold).sum()))
    return cfg


def sr_scale_prune_cfg(G, threshold, n_blocks=16):
    """scale_prune: gamma > threshold of every residual_blocks.i.conv_block1 BatchNorm (SRGAN.py:808-826)."""
    cfg = []
    unprunable = ["conv_block2.conv_block.1"] + ["residual_blocks.%d.conv_block2.conv_block.1" % i for i in range(n_blocks)]
    for k, v in G.items():
        if not k.endswith(".weight") or v.dim() != 1 or (k[:-6] + "running_mean") not in G:
            continue
        if k[:-len(".weight")] in unprunable:
            continue
        cfg.append(int((v.detach() > threshold).sum()))
    return cfg
