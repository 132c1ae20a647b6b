"""CPU oracle for the GCC SAGAN cooperative-compression step.

TEST INFRASTRUCTURE ONLY (same rules as gcc_oracle.py).  Functional fp32 torch-CPU restatement of

  SpectralNorm ............... /root/reference/models/SAGAN.py:14-71   (one power iteration per forward, in place)
  Self_Attn .................. /root/reference/models/SAGAN.py:73-107
  Generator / Discriminators . /root/reference/models/SAGAN.py:109-277 (image size 64 only, as the reference)
  training step .............. /root/reference/models/SAGAN.py:364-533

Reference behaviours that matter for parity and are reproduced here:
  * ``set_requires_grad(netD, True)`` also switches on the power-iteration vectors weight_u / weight_v of the
    discriminator; they then receive a gradient through sigma = u . (W v) and are stepped by optimizer_D.
  * the student's optimizer_G / optimizer_D parameter lists hold the spectral-normed convs' and the attention
    convs' parameters TWICE (wrapper module + inner module both match, SAGAN.py:318-326,339-344): torch's
    single-tensor Adam then steps each of them twice per ``step()`` with the same gradient.
Pinned by ``tests/golden/sagan_tiny.pt`` (oracle/make_golden.py runs the unmodified reference on CPU).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .gcc_oracle import Opt, batch_norm, det_normal, det_uniform, gan_loss, gate_mask, gram

LRELU = 0.1


class SAOpt(Opt):
    """options.parse() for --model sagan (options.py:206-224): hinge loss, lr 1e-4 (D: 4x), betas (0, 0.9)."""

    def __init__(self, **kw):
        super().__init__(gan_mode="hinge", lr=1e-4, ngf=16, ndf=16, teacher_ngf=32, teacher_ndf=32, lambda_content=1e-3,
                         lambda_gram=1e-1, lambda_L1=0.0)
        self.z_dim = 128
        for k, v in kw.items():
            setattr(self, k, v)


class DupAdam:
    """torch.optim.Adam (single-tensor loop) over a parameter LIST that may hold the same tensor several times:
    state is keyed by the tensor, every occurrence performs a full update with the same gradient."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.state = {}

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self):
        with torch.no_grad():
            for p in self.params:
                if p.grad is None:
                    continue
                st = self.state.setdefault(id(p), {"t": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
                g = p.grad
                st["t"] += 1
                t = st["t"]
                st["m"].mul_(self.b1).add_(g, alpha=1 - self.b1)
                st["v"].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                bc1 = 1 - self.b1 ** t
                bc2 = 1 - self.b2 ** t
                denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(self.eps)
                p.addcdiv_(st["m"], denom, value=-self.lr / bc1)


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


def spectral_weight(P, name):
    """SpectralNorm._update_u_v (SAGAN.py:25-38): one power iteration on .data, sigma differentiable in w_bar (and in
    u, v when they require grad)."""
    u, v, w = P[name + ".weight_u"], P[name + ".weight_v"], P[name + ".weight_bar"]
    h = w.shape[0]
    # `.data =` exactly as the reference: no version bump, and autograd's saved references to the leaves u / v
    # see the NEW values when an earlier forward of the same net is back-propagated later (backward_D runs
    # netD(real) and netD(fake) before one backward: the first pass's gradient uses the second pass's u, v)
    v.data = l2normalize(torch.mv(w.reshape(h, -1).data.t(), u.data))
    u.data = l2normalize(torch.mv(w.reshape(h, -1).data, v.data))
    sigma = u.dot(w.reshape(h, -1).mv(v))
    return w / sigma


def self_attn(P, pre, x):
    b, c, wd, ht = x.shape
    q = F.conv2d(x, P[pre + ".query_conv.weight"], P[pre + ".query_conv.bias"]).reshape(b, -1, wd * ht).permute(0, 2, 1)
    k = F.conv2d(x, P[pre + ".key_conv.weight"], P[pre + ".key_conv.bias"]).reshape(b, -1, wd * ht)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    val = F.conv2d(x, P[pre + ".value_conv.weight"], P[pre + ".value_conv.bias"]).reshape(b, -1, wd * ht)
    out = torch.bmm(val, att.permute(0, 2, 1)).reshape(b, c, wd, ht)
    return P[pre + ".gamma"] * out + x


def _attn_shapes(s, pre, c):
    s[pre + ".gamma"] = (1,)
    for n, co in (("query_conv", c // 8), ("key_conv", c // 8), ("value_conv", c)):
        s["%s.%s.weight" % (pre, n)] = (co, c, 1, 1)
        s["%s.%s.bias" % (pre, n)] = (co,)


def generator_channels(ngf, f=None):
    return [ngf * 8, ngf * 4, ngf * 2, ngf] if f is None else list(f)


def generator_shapes(ngf, z_dim=128, f=None):
    c = generator_channels(ngf, f)
    s = OrderedDict()
    cin = z_dim
    for i, co in enumerate(c):
        m = "l%d.0.module" % (i + 1)
        s[m + ".bias"] = (co,)
        s[m + ".weight_u"] = (cin,)
        s[m + ".weight_v"] = (co * 16,)
        s[m + ".weight_bar"] = (cin, co, 4, 4)
        b = "l%d.1" % (i + 1)
        s[b + ".weight"] = (co,)
        s[b + ".bias"] = (co,)
        s[b + ".running_mean"] = (co,)
        s[b + ".running_var"] = (co,)
        s[b + ".num_batches_tracked"] = ()
        cin = co
    s["last.0.weight"] = (c[3], 3, 4, 4)
    s["last.0.bias"] = (3,)
    _attn_shapes(s, "attn1", c[2])
    _attn_shapes(s, "attn2", c[3])
    return s


def disc_shapes(ndf, gated=False):
    s = OrderedDict()
    cin = 3
    for i in range(4):
        co = ndf * (2 ** i)
        m = "l%d.0.module" % (i + 1)
        s[m + ".bias"] = (co,)
        s[m + ".weight_u"] = (co,)
        s[m + ".weight_v"] = (cin * 16,)
        s[m + ".weight_bar"] = (co, cin, 4, 4)
        if gated:
            s["l%d.1.alpha" % (i + 1)] = (co,)
        cin = co
    s["last.0.weight"] = (1, cin, 4, 4)
    s["last.0.bias"] = (1,)
    _attn_shapes(s, "attn1", ndf * 4)
    _attn_shapes(s, "attn2", ndf * 8)
    return s


def make_params(shapes, prefix):
    """Name-seeded deterministic parameters (same convention as gcc_oracle.init_like_reference; u / v normalised
    N(0,1) vectors, attention gamma 0.1 instead of the reference's 0 so that the attention path carries signal)."""
    P = OrderedDict()
    for k, shape in shapes.items():
        full = prefix + k
        if k.endswith("num_batches_tracked"):
            P[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            P[k] = torch.zeros(shape)
        elif k.endswith("running_var"):
            P[k] = torch.ones(shape)
        elif k.endswith("alpha"):
            P[k] = torch.ones(shape)
        elif k.endswith("weight_u") or k.endswith("weight_v"):
            P[k] = l2normalize(det_normal(full, shape))
        elif k.endswith("gamma"):
            P[k] = torch.full(shape, 0.1)
        elif len(shape) == 4:
            P[k] = det_normal(full, shape, 0.0, 0.05)
        elif k.endswith(".weight"):
            P[k] = det_normal(full, shape, 1.0, 0.02)          # BatchNorm scale
        elif k.endswith(".bias"):
            is_bn = (k[:-4] + "running_mean") in shapes
            P[k] = det_normal(full, shape, 0.0, 1.0 if is_bn else 0.02)
        else:
            raise KeyError(k)
    for k, v in P.items():
        if v.dtype == torch.float32 and not ("running" in k or k.endswith("weight_u") or k.endswith("weight_v")):
            v.requires_grad_(True)
    return P


def generator_forward(P, z, train=True):
    out = z.reshape(z.shape[0], z.shape[1], 1, 1)
    taps = []
    for i in range(4):
        m = "l%d.0.module" % (i + 1)
        w = spectral_weight(P, m)
        out = F.conv_transpose2d(out, w, P[m + ".bias"], stride=1 if i == 0 else 2, padding=0 if i == 0 else 1)
        out = F.relu(batch_norm(out, P, "l%d.1" % (i + 1), train))
        if i == 1:
            taps.append(out)
        if i == 2:
            out = self_attn(P, "attn1", out)
    out = self_attn(P, "attn2", out)
    taps.append(out)
    out = torch.tanh(F.conv_transpose2d(out, P["last.0.weight"], P["last.0.bias"], stride=2, padding=1))
    return out, taps


def disc_forward(P, x, gated=False, threshold=0.5):
    out, taps = x, []
    for i in range(4):
        m = "l%d.0.module" % (i + 1)
        w = spectral_weight(P, m)
        out = F.conv2d(out, w, P[m + ".bias"], stride=2, padding=1)
        if gated:
            out = out * gate_mask(P["l%d.1.alpha" % (i + 1)], threshold)[None, :, None, None]
        out = F.leaky_relu(out, LRELU)
        if i == 1:
            taps.append(out)
        if i == 2:
            out = self_attn(P, "attn1", out)
    out = self_attn(P, "attn2", out)
    taps.append(out)
    out = F.conv2d(out, P["last.0.weight"], P["last.0.bias"])
    return out.squeeze(), taps


def _dup_list(P, with_uv_dups=True):
    """Parameter list in the order (and multiplicity) of the reference's collection loop over named_modules():
    SpectralNorm wrapper + its inner conv, Self_Attn + its three inner convs (SAGAN.py:318-326,339-344)."""
    lst = []
    names = list(P.keys())
    mods = OrderedDict()
    for k in names:
        if k.endswith("alpha"):
            continue
        if ".module." in k:
            mods.setdefault(k.split(".module.")[0], []).append(k)
    # module order of the reference: l4, l1, l2, l3, last, attn1, attn2 (l4 is assigned first, SAGAN.py:139-147)
    for l in ("l4", "l1", "l2", "l3"):
        sn = [k for k in names if k.startswith(l + ".0.module.")]
        lst += [P[k] for k in sn]          # SpectralNorm wrapper
        lst += [P[k] for k in sn]          # inner ConvTranspose2d / Conv2d
        lst += [P[k] for k in names if k.startswith(l + ".1.") and (k.endswith(".weight") or k.endswith(".bias"))]
    lst += [P["last.0.weight"], P["last.0.bias"]]
    for a in ("attn1", "attn2"):
        lst.append(P[a + ".gamma"])
        convs = [P["%s.%s.%s" % (a, c, wb)] for c in ("query_conv", "key_conv", "value_conv") for wb in ("weight", "bias")]
        lst += convs + convs
    return lst


class SAGANOracle:
    def __init__(self, opt, tag, filter_cfgs=None):
        self.opt, self.tag, self.filter_cfgs = opt, tag, filter_cfgs
        self.G = make_params(generator_shapes(opt.ngf, opt.z_dim, filter_cfgs), tag + ".netG.")
        self.gated = bool(opt.darts_discriminator)
        self.D = make_params(disc_shapes(opt.ndf, self.gated), tag + ".netD.")
        self.transform = []
        if opt.online_distillation:
            t_ch = [opt.teacher_ngf * 4, opt.teacher_ngf]
            s_ch = [opt.ngf * 4, opt.ngf] if filter_cfgs is None else [filter_cfgs[1], filter_cfgs[3]]
            for i in range(2):
                w = det_uniform("%s.transform.%d" % (tag, i), (t_ch[i], s_ch[i], 1, 1), 1.0 / math.sqrt(s_ch[i]))
                w.requires_grad_(True)
                self.transform.append(w)
            # collection order: transform conv i is appended when the loop meets the tapped module ('l2', 'attn2')
            g_list = _dup_list(self.G)
            self.opt_G = DupAdam(self.transform + g_list, opt.lr, (0.0, 0.9))
        else:
            self.opt_G = DupAdam([v for k, v in self.G.items() if v.dtype == torch.float32 and "running" not in k],
                                 opt.lr, (0.0, 0.9))
        self.d_alphas = [v for k, v in self.D.items() if k.endswith("alpha")]
        self.d_all = [v for k, v in self.D.items() if v.dtype == torch.float32]
        # set_netD_weight_grad only reaches Conv2d / BatchNorm2d modules: everything except the attention gammas and alphas
        self.d_conv = [v for k, v in self.D.items() if v.dtype == torch.float32 and not k.endswith("alpha")
                       and not k.endswith("gamma")]
        if self.gated:
            self.opt_D = DupAdam(_dup_list(self.D), opt.lr * 4, (0.0, 0.9))
            self.opt_arch = DupAdam(self.d_alphas, opt.arch_lr)
        else:
            self.opt_D = DupAdam(self.d_all, opt.lr * 4, (0.0, 0.9))
            self.opt_arch = None
        self.teacher = None
        self.current_D_arch_diff_loss = 0.0
        self.train_mode = True

    def netG(self, z):
        return generator_forward(self.G, z, self.train_mode)

    def netD(self, x):
        return disc_forward(self.D, x, self.gated, self.opt.threshold)

    def _set_grad(self, tensors, flag):
        for t in tensors:
            t.requires_grad_(flag)

    def set_input(self, z, real):
        self.z, self.real_img = z, real
        self._raw = (z, real)

    def forward(self):
        self.fake_img, self.g_taps = self.netG(self.z)

    def backward_D(self):
        m = self.opt.gan_mode
        pred_real, _ = self.netD(self.real_img)
        self.loss_D_real = gan_loss(m, pred_real, True, True)
        pred_fake, self.d_taps = self.netD(self.fake_img.detach())
        self.loss_D_fake = gan_loss(m, pred_fake, False, True)
        self.loss_D = self.loss_D_fake + self.loss_D_real
        self.loss_D.backward()

    def backward_G(self):
        o = self.opt
        pred_fake, self.d_taps = self.netD(self.fake_img)
        self.loss_G_GAN = gan_loss(o.gan_mode, pred_fake, True, False)
        self.loss_G = self.loss_G_GAN
        if o.online_distillation:
            T = self.teacher
            feats = list(self.g_taps)
            _, t_d = T.netD(self.fake_img)
            feats += list(t_d)
            self.loss_content = 0.0
            self.loss_gram = 0.0
            for i, f in enumerate(feats):
                if i < 2:
                    f = F.conv2d(f, self.transform[i])
                tgt = self.target_features[i].detach()
                self.loss_gram = self.loss_gram + torch.sqrt(((gram(f) - gram(tgt)) ** 2).mean())
                self.loss_content = self.loss_content + torch.sqrt(((f - tgt) ** 2).mean())
            self.loss_gram = o.lambda_gram * self.loss_gram
            self.loss_content = o.lambda_content * self.loss_content
            self.loss_L1 = o.lambda_L1 * (self.fake_img - T.fake_img.detach()).abs().mean()
            self.loss_G = self.loss_G + self.loss_gram + self.loss_content + self.loss_L1
            # the reference aliases loss_G to loss_G_GAN and accumulates IN PLACE (SAGAN.py:463,483-485): the value it
            # reports as G_GAN is the whole generator loss
            self.loss_G_GAN = self.loss_G
        self.loss_G.backward()

    def optimize_parameters(self):
        if self.opt.online_distillation:
            T = self.teacher
            T.set_input(*self._raw)
            T.optimize_parameters()
            self.target_features = [f.detach().clone() for f in (T.g_taps + T.d_taps)]
        self.forward()
        self._set_grad(self.d_all, True)            # set_requires_grad(netD, True): u and v included
        self._set_grad(self.d_alphas, False)
        self.opt_D.zero_grad()
        self.backward_D()
        self.opt_D.step()
        self._set_grad(self.d_all, False)
        self.opt_G.zero_grad()
        self.backward_G()
        self.opt_G.step()

    def get_D_arch_diff(self, is_teacher):
        m = self.opt.gan_mode
        if is_teacher:
            self._set_grad(self.d_all, False)
        pred_fake, _ = self.netD(self.fake_img.detach())
        self.loss_D_arch_fake = gan_loss(m, pred_fake, False, True)
        self.loss_D_arch_fake_real = gan_loss(m, pred_fake, True, False)
        pred_real, _ = self.netD(self.real_img)
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
        self._set_grad(self.d_all, True)
        self._set_grad(self.d_conv, False)          # set_netD_weight_grad(False): the attention gammas stay on
        self.opt_arch.zero_grad()
        self.loss_teacher_D_arch_diff = T.get_D_arch_diff(True)
        self.loss_D_arch_diff = self.get_D_arch_diff(False)
        self.loss_D_arch = (self.loss_D_arch_diff - self.loss_teacher_D_arch_diff).abs()
        self.loss_D_arch = self.loss_D_arch + self.loss_D_arch_real + self.loss_D_arch_fake
        self.loss_D_arch.backward()
        self.opt_arch.step()

    def current_masks(self):
        return [((torch.sign(a.detach() - self.opt.threshold) + 1) / 2) for a in self.d_alphas]

    LOSSES = ["G_GAN", "D_real", "D_fake", "D_arch_diff", "D_arch", "teacher_D_arch_diff", "content", "gram", "L1"]

    def get_current_losses(self):
        return OrderedDict((n, float(getattr(self, "loss_" + n))) for n in self.LOSSES if hasattr(self, "loss_" + n))


def build_sa_pair(opt, filter_cfgs=None):
    import copy
    student = SAGANOracle(opt, "S", filter_cfgs)
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    teacher = SAGANOracle(topt, "T")
    student.teacher = teacher
    return student, teacher


def scale_prune_cfg(G, threshold):
    """SAGANModel.scale_prune (SAGAN.py:741-764): count of BatchNorm scales above the threshold per level l1..l4."""
    return [int((G["l%d.1.weight" % i].detach() > threshold).sum()) for i in (1, 2, 3, 4)]
