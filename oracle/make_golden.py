"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py

Shims (none touches arithmetic): stub modules for the missing `thop` and
`skimage.metrics`; parameters are overwritten with the name-seeded deterministic values of
oracle.gcc_oracle.init_like_reference so that fixtures do not depend on torch's RNG stream.
The fixtures hold summary statistics (sum, |sum|, sum of squares, 24 strided samples) of every
tensor of interest plus exact masks / channel counts, not the tensors themselves.
"""
import copy
import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

from oracle import gcc_oracle as O  # noqa: E402


def import_reference():
    thop = types.ModuleType("thop")
    thop.profile = lambda *a, **k: (0.0, 0.0)
    sys.modules["thop"] = thop
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.metrics")
    skm.peak_signal_noise_ratio = lambda *a, **k: 0.0
    skm.structural_similarity = lambda *a, **k: 0.0
    sk.metrics = skm
    sys.modules["skimage"] = sk
    sys.modules["skimage.metrics"] = skm
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    from options import options  # noqa
    if not any(a.dest == "generator_only" for a in options.parser._actions):
        options.parser.add_argument("--generator_only", action="store_true")
    import models.Pix2Pix as P2P  # noqa
    os.chdir(cwd)
    return options, P2P


def stats(t, nsamp=24):
    t = t.detach().float().reshape(-1)
    n = t.numel()
    idx = torch.linspace(0, n - 1, min(nsamp, n)).long()
    return {"n": n, "sum": float(t.double().sum()), "abs": float(t.double().abs().sum()),
            "sq": float((t.double() ** 2).sum()), "samples": t[idx].tolist(), "idx": idx.tolist()}


def make_opt(options, argv):
    old = sys.argv
    sys.argv = ["train.py"] + argv
    try:
        opt = options.parse()
    finally:
        sys.argv = old
    opt.isTrain = True
    return opt


def set_det_params(model, tag):
    sd = model.netG.state_dict()
    O.init_like_reference(sd, tag + ".netG.")
    model.netG.load_state_dict(sd)
    sd = model.netD.state_dict()
    O.init_like_reference(sd, tag + ".netD.")
    model.netD.load_state_dict(sd)
    if hasattr(model, "transform_convs"):
        import math
        for i, conv in enumerate(model.transform_convs):
            cin = conv.weight.shape[1]
            conv.weight.data.copy_(O.det_uniform("%s.transform.%d" % (tag, i), conv.weight.shape, 1.0 / math.sqrt(cin)))


def build_reference_pair(options, P2P, argv, small, cfgs=(None, None)):
    opt = make_opt(options, argv)
    for k, v in small.items():
        setattr(opt, k, v)
    model = P2P.Pix2PixModel(opt, filter_cfgs=cfgs[0], channel_cfgs=cfgs[1])
    teacher_opt = copy.deepcopy(opt)  # train.py:92-105
    teacher_opt.ngf = opt.teacher_ngf
    teacher_opt.ndf = opt.teacher_ndf
    teacher_opt.darts_discriminator = False
    teacher_opt.online_distillation = False
    teacher_opt.generator_only = False
    teacher = P2P.Pix2PixModel(teacher_opt)
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    set_det_params(model, "S")
    set_det_params(teacher, "T")
    model.model_train()
    return opt, model, teacher


def record_model(rec, prefix, model):
    for k, v in model.netG.state_dict().items():
        rec[prefix + ".G." + k] = stats(v)
    for k, v in model.netD.state_dict().items():
        rec[prefix + ".D." + k] = stats(v)
    for k, v in model.netG.named_parameters():
        if v.grad is not None:
            rec[prefix + ".G.grad." + k] = stats(v.grad)
    for k, v in model.netD.named_parameters():
        if v.grad is not None:
            rec[prefix + ".D.grad." + k] = stats(v.grad)
    if hasattr(model, "transform_convs"):
        for i, c in enumerate(model.transform_convs):
            rec[prefix + ".transform.%d" % i] = stats(c.weight)
            if c.weight.grad is not None:
                rec[prefix + ".transform.grad.%d" % i] = stats(c.weight.grad)


def run_step_case(options, P2P, name, argv, small, batch, iters, cfgs=(None, None)):
    torch.manual_seed(0)
    opt, model, teacher = build_reference_pair(options, P2P, argv, small, cfgs)
    out = {"config": {"argv": argv, "small": small, "batch": batch, "iters": iters, "cfgs": cfgs,
                      "direction": opt.direction}, "iters": []}
    for it in range(iters):
        rec = {}
        A = O.det_image("%s.A.%d" % (name, it), batch, 3, 256, 256)
        B = O.det_image("%s.B.%d" % (name, it), batch, 3, 256, 256)
        model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
        model.optimize_parameters()
        rec["fake_B"] = stats(model.fake_B)
        rec["Tfake_B"] = stats(teacher.fake_B)
        for i, f in enumerate(model.target_distillation_features):
            rec["target_feature.%d" % i] = stats(f)
        for i, f in enumerate(model.get_distillation_features()):
            rec["student_feature.%d" % i] = stats(f)
        record_model(rec, "S", model)
        record_model(rec, "T", teacher)
        for n in ("G_GAN", "G_L1", "D_real", "D_fake", "content", "gram"):
            rec["loss.S." + n] = float(getattr(model, "loss_" + n))
        for n in ("G_GAN", "G_L1", "D_real", "D_fake"):
            rec["loss.T." + n] = float(getattr(teacher, "loss_" + n))
        # arch step on a validation batch (train.py:147-151)
        vA = O.det_image("%s.vA.%d" % (name, it), batch, 3, 256, 256)
        vB = O.det_image("%s.vB.%d" % (name, it), batch, 3, 256, 256)
        model.set_input({"A": vA, "B": vB, "A_paths": "", "B_paths": ""})
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
        for n in ("D_arch_diff", "D_arch", "teacher_D_arch_diff"):
            rec["loss.S." + n] = float(getattr(model, "loss_" + n))
        for k, v in model.netD.named_parameters():
            if k.endswith("alpha"):
                rec["arch.alpha." + k] = stats(v)
                rec["arch.alpha_grad." + k] = stats(v.grad)
        for k, v in model.netD.state_dict().items():
            if "running" in k:
                rec["arch.S.D." + k] = stats(v)
        for k, v in teacher.netD.state_dict().items():
            if "running" in k:
                rec["arch.T.D." + k] = stats(v)
        rec["losses"] = dict(model.get_current_losses())
        out["iters"].append(rec)
        print(name, "iter", it, {k: round(v, 5) for k, v in rec["losses"].items()}, flush=True)
    return out


def run_cycle_case(options, name, small, batch, iters, size, cfgs=(None, None)):
    """MobileCycleGANModel (models/CycleGAN.py) driven exactly like train.py drives it."""
    import math
    import models.CycleGAN as CG
    torch.manual_seed(0)
    opt = make_opt(options, ["--dataroot", "x/horse2zebra", "--model", "cyclegan", "--darts_discriminator",
                             "--online_distillation", "--lambda_content", "0.01", "--lambda_gram", "10", "--gpu_ids", "-1"])
    for k, v in small.items():
        setattr(opt, k, v)
    model = CG.MobileCycleGANModel(opt, cfg_AtoB=cfgs[0], cfg_BtoA=cfgs[1])
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    topt.generator_only = False
    teacher = CG.MobileCycleGANModel(topt)
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    for m, tag in ((model, "S"), (teacher, "T")):
        for k in "AB":
            for kind in ("G", "D"):
                net = getattr(m, "net%s_%s" % (kind, k))
                sd = net.state_dict()
                O.init_like_reference(sd, "%s.net%s_%s." % (tag, kind, k))
                net.load_state_dict(sd)
            for i, conv in enumerate(getattr(m, "transform_%s_convs" % k, [])):
                cin = conv.weight.shape[1]
                conv.weight.data.copy_(O.det_uniform("%s.transform_%s.%d" % (tag, k, i), conv.weight.shape, 1.0 / math.sqrt(cin)))
    model.model_train()
    out = {"config": {"small": small, "batch": batch, "iters": iters, "size": size, "cfgs": cfgs,
                      "direction": opt.direction}, "iters": []}
    for it in range(iters):
        rec = {}
        A = O.det_image("%s.A.%d" % (name, it), batch, 3, size, size)
        B = O.det_image("%s.B.%d" % (name, it), batch, 3, size, size)
        model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
        model.optimize_parameters()
        for n in ("fake_A", "fake_B", "rec_A", "rec_B", "idt_A", "idt_B"):
            rec[n] = stats(getattr(model, n))
        rec["Tfake_A"], rec["Tfake_B"] = stats(teacher.fake_A), stats(teacher.fake_B)
        for k in "AB":
            for i, f in enumerate(getattr(model, "target_distillation_%s_features" % k)):
                rec["target_%s.%d" % (k, i)] = stats(f)
            for kind in ("G", "D"):
                for tag, m in (("S", model), ("T", teacher)):
                    net = getattr(m, "net%s_%s" % (kind, k))
                    for kk, v in net.state_dict().items():
                        rec["%s.%s_%s.%s" % (tag, kind, k, kk)] = stats(v)
                    for kk, v in net.named_parameters():
                        if v.grad is not None:
                            rec["%s.%s_%s.grad.%s" % (tag, kind, k, kk)] = stats(v.grad)
            for i, c in enumerate(getattr(model, "transform_%s_convs" % k)):
                rec["S.transform_%s.%d" % (k, i)] = stats(c.weight)
                rec["S.transform_%s.grad.%d" % (k, i)] = stats(c.weight.grad)
        vA = O.det_image("%s.vA.%d" % (name, it), batch, 3, size, size)
        vB = O.det_image("%s.vB.%d" % (name, it), batch, 3, size, size)
        model.set_input({"A": vA, "B": vB, "A_paths": "", "B_paths": ""})
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
        for k in "AB":
            for kk, v in getattr(model, "netD_" + k).named_parameters():
                if kk.endswith("alpha"):
                    rec["arch.alpha_%s.%s" % (k, kk)] = stats(v)
                    rec["arch.alpha_grad_%s.%s" % (k, kk)] = stats(v.grad)
        rec["losses"] = {k: float(v) for k, v in model.get_current_losses().items()}
        out["iters"].append(rec)
        print(name, "iter", it, {k: round(v, 5) for k, v in rec["losses"].items()}, flush=True)
    return out


def run_sr_case(options, name, small, batch, iters, lr_size):
    """SRGAN (models/SRGAN.py) driven exactly like train.py drives it.  The one value-touching shim: torchvision's
    vgg19 is built with weights=None (no network for the pretrained file) and its conv parameters are overwritten
    with oracle.srgan_oracle.make_vgg_params() on both models."""
    import math
    import torchvision
    from oracle import srgan_oracle as S
    import models.GANLoss as GL
    real_vgg19 = torchvision.models.vgg19
    GL.vgg19 = lambda pretrained=True: real_vgg19(weights=None)
    import models.SRGAN as SR
    torch.manual_seed(0)
    opt = make_opt(options, ["--dataroot", "x/sr", "--model", "srgan", "--darts_discriminator", "--online_distillation",
                             "--lambda_content", "1e-3", "--lambda_gram", "1e-1", "--gpu_ids", "-1"])
    for k, v in small.items():
        setattr(opt, k, v)
    opt.generator_only = False
    model = SR.SRGAN(opt)
    topt = copy.deepcopy(opt)
    topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
    topt.darts_discriminator = False
    topt.online_distillation = False
    topt.generator_only = False
    teacher = SR.SRGAN(topt)
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    vggP = S.make_vgg_params()
    for m, tag in ((model, "S"), (teacher, "T")):
        sd = m.netG.state_dict()
        O.init_like_reference(sd, tag + ".netG.")
        S.init_sr_params(sd, tag + ".netG.")
        m.netG.load_state_dict(sd)
        sd = m.netD.state_dict()
        O.init_like_reference(sd, tag + ".netD.")
        m.netD.load_state_dict(sd)
        m.truncated_vgg19.load_state_dict(vggP)
        for i, conv in enumerate(getattr(m, "transform_convs", [])):
            cin = conv.weight.shape[1]
            conv.weight.data.copy_(O.det_uniform("%s.transform.%d" % (tag, i), conv.weight.shape, 1.0 / math.sqrt(cin)))
    model.model_train()
    out = {"config": {"small": small, "batch": batch, "iters": iters, "lr_size": lr_size,
                      "lambdas": {k: getattr(opt, k) for k in ("lambda_content", "lambda_gram", "lambda_L1",
                                                               "lambda_SR_adversarial", "lambda_SR_content",
                                                               "lambda_SR_perceptual", "lr", "arch_lr", "gan_mode")}},
           "iters": []}
    for it in range(iters):
        rec = {}
        lr = S.convert_to_imagenet(O.det_image("%s.lr.%d" % (name, it), batch, 3, lr_size, lr_size))
        hr = O.det_image("%s.hr.%d" % (name, it), batch, 3, 4 * lr_size, 4 * lr_size)
        model.set_input({"lr": lr, "hr": hr, "lr_names": "", "hr_names": ""})
        model.optimize_parameters()
        rec["fake_hr"], rec["Tfake_hr"] = stats(model.fake_hr), stats(teacher.fake_hr)
        for i, f in enumerate(model.target_distillation_features):
            rec["target.%d" % i] = stats(f)
        record_model(rec, "S", model)
        record_model(rec, "T", teacher)
        vlr = S.convert_to_imagenet(O.det_image("%s.vlr.%d" % (name, it), batch, 3, lr_size, lr_size))
        vhr = O.det_image("%s.vhr.%d" % (name, it), batch, 3, 4 * lr_size, 4 * lr_size)
        model.set_input({"lr": vlr, "hr": vhr, "lr_names": "", "hr_names": ""})
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
        for kk, v in model.netD.named_parameters():
            if kk.endswith("alpha"):
                rec["arch.alpha." + kk] = stats(v)
                rec["arch.alpha_grad." + kk] = stats(v.grad)
        rec["losses"] = {k: float(v) for k, v in model.get_current_losses().items()}
        out["iters"].append(rec)
        print(name, "iter", it, {k: round(v, 5) for k, v in rec["losses"].items()}, flush=True)
    # prune index selection on the final student weights (SRGAN.py:773-837) and on spread-out deterministic ones
    sd = {k: O.det_normal("srprune." + k, v.shape, 0.5, 0.3) if v.dim() == 1 and v.numel() > 1 else v
          for k, v in model.netG.state_dict().items()}
    model.netG.load_state_dict(sd)
    prunes = {}
    for mode, thr in (("scale", 0.45), ("scale", 0.8), ("norm", None)):
        o2 = copy.deepcopy(opt)
        o2.scale_prune, o2.norm_prune = mode == "scale", mode == "norm"
        model.opt = o2
        if thr is None:
            norms = [float(torch.sum(torch.abs(m.weight.data), (1, 2, 3)).median()) for n, m in model.netG.named_modules()
                     if isinstance(m, torch.nn.Conv2d) and "conv_block1.conv_block.0" in n and n.startswith("residual")]
            thr = sum(norms) / len(norms)
        pruned = model.prune(thr)
        prunes["%s@%.6f" % (mode, thr)] = {"thr": thr, "cfg": list(pruned.filter_cfgs)}
    model.opt = opt
    out["prune"] = prunes
    out["prune_state"] = {k: stats(v) for k, v in model.netG.state_dict().items() if v.dim() >= 1}
    return out


def run_sa_case(options, name, small, batch, iters):
    """SAGANModel (models/SAGAN.py) driven exactly like train.py drives it (image size 64, the only one it supports).
    Shim: the integer Adam betas (0, 0.9) of SAGAN.py:302,327,346,356 are coerced to floats."""
    import math
    from oracle import sagan_oracle as SA
    import models.SAGAN as SG
    torch.manual_seed(0)
    opt = make_opt(options, ["--dataroot", "x/celeb", "--model", "sagan", "--darts_discriminator", "--online_distillation",
                             "--lambda_content", "1e-3", "--lambda_gram", "1e-1", "--gpu_ids", "-1"])
    for k, v in small.items():
        setattr(opt, k, v)
    # shim: this torch rejects integer betas (0, 0.9): coerce to floats, nothing else changes
    real_adam = torch.optim.Adam

    class _FloatBetasAdam(real_adam):
        def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), **kw):
            super().__init__(params, lr=lr, betas=(float(betas[0]), float(betas[1])), **kw)
    torch.optim.Adam = _FloatBetasAdam
    try:
        model = SG.SAGANModel(opt)
        topt = copy.deepcopy(opt)
        topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
        topt.darts_discriminator = False
        topt.online_distillation = False
        topt.generator_only = False
        teacher = SG.SAGANModel(topt)
    finally:
        torch.optim.Adam = real_adam
    teacher.model_train()
    setattr(model, "teacher_model", teacher)
    model.init_distillation()
    teacher.init_distillation()
    sopt = SA.SAOpt(**small)
    S, T = SA.build_sa_pair(sopt)
    for m, orc, tag in ((model, S, "S"), (teacher, T, "T")):
        m.netG.load_state_dict({k: v.detach().clone() for k, v in orc.G.items()})
        m.netD.load_state_dict({k: v.detach().clone() for k, v in orc.D.items()})
        for i, conv in enumerate(getattr(m, "transform_convs", [])):
            conv.weight.data.copy_(orc.transform[i].detach())
    model.model_train()
    out = {"config": {"small": small, "batch": batch, "iters": iters,
                      "opts": {k: getattr(opt, k) for k in ("lambda_content", "lambda_gram", "lambda_L1", "lr", "arch_lr",
                                                            "gan_mode", "z_dim", "crop_size")}}, "iters": []}
    for it in range(iters):
        rec = {}
        z = O.det_normal("%s.z.%d" % (name, it), (batch, opt.z_dim))
        real = O.det_image("%s.real.%d" % (name, it), batch, 3, 64, 64)
        model.set_input({"z": z, "real_img": real, "img_path": ""})
        model.optimize_parameters()
        rec["fake_img"], rec["Tfake_img"] = stats(model.fake_img), stats(teacher.fake_img)
        for i, f in enumerate(model.target_distillation_features):
            rec["target.%d" % i] = stats(f)
        record_model(rec, "S", model)
        record_model(rec, "T", teacher)
        vz = O.det_normal("%s.vz.%d" % (name, it), (batch, opt.z_dim))
        vreal = O.det_image("%s.vreal.%d" % (name, it), batch, 3, 64, 64)
        model.set_input({"z": vz, "real_img": vreal, "img_path": ""})
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
        for kk, v in model.netD.named_parameters():
            if kk.endswith("alpha"):
                rec["arch.alpha." + kk] = stats(v)
                rec["arch.alpha_grad." + kk] = stats(v.grad)
        rec["losses"] = {k: float(v) for k, v in model.get_current_losses().items()}
        out["iters"].append(rec)
        print(name, "iter", it, {k: round(v, 5) for k, v in rec["losses"].items()}, flush=True)
    o2 = copy.deepcopy(opt)
    o2.scale_prune = True
    model.opt = o2
    torch.optim.Adam = _FloatBetasAdam
    try:
        out["prune"] = {"%g" % thr: list(model.prune(thr).filter_cfgs) for thr in (0.98, 1.0, 1.02)}
    finally:
        torch.optim.Adam = real_adam
    out["prune_state"] = {k: v.clone() for k, v in model.netG.state_dict().items() if k.endswith(".1.weight")}
    return out


def run_cycle_prune_case(options):
    """get_prunenet_cfg / max_min_conv_norm of MobileCycleGANModel (CycleGAN.py:803-885) on deterministic weights."""
    import models.CycleGAN as CG
    opt = make_opt(options, ["--dataroot", "x/horse2zebra", "--model", "cyclegan", "--ngf", "16", "--ndf", "16",
                             "--gpu_ids", "-1", "--norm_prune"])
    m = CG.MobileCycleGANModel(opt)
    sd = m.netG_A.state_dict()
    O.init_like_reference(sd, "P.netG.")
    m.netG_A.load_state_dict(sd)
    out = {"cfg@%g" % thr: m.get_prunenet_cfg(m.netG_A, thr) for thr in (0.8, 0.95, 1.05, 1.15)}
    mx, mn = m.max_min_conv_norm(m.netG_A)
    out["maxmin"] = (float(mx), float(mn))
    return out


def run_prune_case(options, P2P):
    out = {}
    # U-Net scale / norm prune at fixed thresholds on deterministic weights
    for mode, thr_list in (("scale_prune", [0.98, 1.0, 1.02]), ("norm_prune", [2.0, 6.0, 10.0])):
        opt = make_opt(options, ["--dataroot", "x/cityscapes", "--model", "pix2pix", "--ngf", "16", "--ndf", "16",
                                 "--gpu_ids", "-1", "--" + mode, "--no_dropout"])
        model = P2P.Pix2PixModel(opt)
        set_det_params(model, "P")
        for thr in thr_list:
            pruned = getattr(model, mode)(thr)
            out["%s@%g" % (mode, thr)] = (list(pruned.filter_cfgs), list(pruned.channel_cfgs))
        if mode == "scale_prune":
            mx, mn = model.max_min_bn_scale()
        else:
            mx, mn = model.max_min_conv_norm()
        out[mode + ".maxmin"] = (float(mx), float(mn))
    opt = make_opt(options, ["--dataroot", "x/cityscapes", "--model", "pix2pix", "--ngf", "16", "--ndf", "16",
                             "--gpu_ids", "-1", "--norm_prune", "--backbone", "resnet"])
    model = P2P.Pix2PixModel(opt)
    set_det_params(model, "P")
    for thr in (0.5, 2.3, 2.6):
        pruned = model.resnet_prune(thr)
        out["resnet_prune@%g" % thr] = list(pruned.filter_cfgs)
    mx, mn = model.max_min_conv_norm()
    out["resnet.maxmin"] = (float(mx), float(mn))
    return out


def run_gate_case(P2P):
    from models.DifferentiableOp import DifferentiableOP
    op = DifferentiableOP(6, 0.5)
    op.alpha.data.copy_(torch.tensor([0.7, 0.5, 0.2, 1.0, 0.0, 0.5000001]))
    x = O.det_normal("gate.x", (2, 6, 3, 3))
    x.requires_grad_(True)
    y = op(x)
    gy = O.det_normal("gate.gy", (2, 6, 3, 3))
    y.backward(gy)
    return {"alpha": op.alpha.data.tolist(), "mask": op.get_current_mask().tolist(), "y": y.detach().clone(),
            "dx": x.grad.clone(), "dalpha": op.alpha.grad.clone()}


def run_ganloss_case():
    from models.GANLoss import GANLoss
    pred = O.det_normal("ganloss.pred", (3, 1, 30, 30))
    out = {}
    for mode in ("hinge", "lsgan", "vanilla", "wgangp"):
        crit = GANLoss(mode)
        for real in (True, False):
            out["%s.D.%s" % (mode, real)] = float(crit(pred, real, for_discriminator=True))
        out["%s.G" % mode] = float(crit(pred, True, for_discriminator=False))
    return out


def main():
    options, P2P = import_reference()
    gold = os.path.join(REPO, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    sr_small = {"ngf": 8, "teacher_ngf": 16, "ndf": 8, "teacher_ndf": 16}
    sa_small = {"ngf": 16, "teacher_ngf": 32, "ndf": 16, "teacher_ndf": 32}
    if len(sys.argv) > 1 and sys.argv[1] == "sagan":     # regenerate only the SAGAN fixture
        torch.save(run_sa_case(options, "sa_tiny", sa_small, batch=4, iters=2), os.path.join(gold, "sagan_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "srgan":     # regenerate only the SRGAN fixture
        torch.save(run_sr_case(options, "sr_tiny", sr_small, batch=2, iters=2, lr_size=12),
                   os.path.join(gold, "srgan_tiny.pt"))
        return
    base = ["--dataroot", "x/cityscapes", "--model", "pix2pix", "--darts_discriminator", "--online_distillation",
            "--lambda_content", "50", "--lambda_gram", "1e4", "--gpu_ids", "-1", "--no_dropout"]
    tiny = {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}
    torch.save(run_step_case(options, P2P, "unet_tiny", base, tiny, batch=2, iters=2),
               os.path.join(gold, "pix2pix_unet_tiny.pt"))
    fc = [8, 13, 30, 61, 64, 59, 40, 64, 37, 50, 64, 48, 27, 14, 5]
    cc = [8, 13, 30, 61, 64, 59, 40, 64, 37 + 40, 50 + 59, 64 + 64, 48 + 61, 27 + 30, 14 + 13, 5 + 8]
    torch.save(run_step_case(options, P2P, "unet_pruned", base, tiny, batch=1, iters=1, cfgs=(fc, cc)),
               os.path.join(gold, "pix2pix_unet_pruned.pt"))
    rcfg = [8, 16, 29, 21, 29, 17, 29, 30, 29, 11, 29, 25, 29, 32, 29, 9, 29, 27, 29, 19, 29, 13, 7]
    torch.save(run_step_case(options, P2P, "resnet_tiny", base + ["--backbone", "resnet"], tiny, batch=1, iters=1,
                             cfgs=(rcfg, None)), os.path.join(gold, "pix2pix_resnet_tiny.pt"))
    ccfg = [8, 16, 29, 21, 29, 17, 29, 30, 29, 11, 29, 25, 29, 32, 29, 9, 29, 27, 29, 19, 29, 13, 7]
    torch.save(run_cycle_case(options, "cycle_tiny", tiny, batch=1, iters=2, size=128, cfgs=(ccfg, None)),
               os.path.join(gold, "cyclegan_tiny.pt"))
    torch.save(run_cycle_prune_case(options), os.path.join(gold, "cyclegan_prune.pt"))
    torch.save({"prune": run_prune_case(options, P2P), "gate": run_gate_case(P2P), "ganloss": run_ganloss_case()},
               os.path.join(gold, "pix2pix_small_ops.pt"))
    torch.save(run_sr_case(options, "sr_tiny", sr_small, batch=2, iters=2, lr_size=12), os.path.join(gold, "srgan_tiny.pt"))
    torch.save(run_sa_case(options, "sa_tiny", sa_small, batch=4, iters=2), os.path.join(gold, "sagan_tiny.pt"))
    print("golden fixtures written to", gold)


if __name__ == "__main__":
    main()
