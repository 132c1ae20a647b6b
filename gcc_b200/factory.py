"""train.py:84-105 as a function: options -> model class -> (student, teacher) wired for online distillation, plus
synthetic batches of the shapes the reference's data loaders produce (SURVEY.md 8b "input dict contract").  Used by
bench.py, the examples and the tests so that all of them drive the models through the same calls train.py makes."""
import copy

import torch

from . import options

DATAROOTS = {"pix2pix": "./database/cityscapes/", "cyclegan": "./database/horse2zebra", "srgan": "./database/sr",
             "sagan": "./database/celeb"}
# loss weights of the reference's training scripts (scripts/*/train.sh)
DISTILL_WEIGHTS = {"pix2pix": ("50", "1e4"), "cyclegan": ("1e-2", "1e-1"), "srgan": ("1e-2", "1e-1"),
                   "sagan": ("1e-2", "1e-1")}


def make_opt(model="pix2pix", device_index=0, argv=(), **overrides):
    lc, lg = DISTILL_WEIGHTS[model]
    base = ["--dataroot", DATAROOTS[model], "--model", model, "--darts_discriminator", "--online_distillation",
            "--lambda_content", lc, "--lambda_gram", lg, "--gpu_ids", str(device_index)]
    opt = options.parse(base + list(argv))
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


def build_pair(opt, cfgs=(None, None)):
    """model = model_class(opt[, cfgs]); teacher = same class at the teacher widths with a plain discriminator and no
    distillation (train.py:92-105)."""
    cls = options.get_model_class(opt)
    model = cls(opt, cfgs[0], cfgs[1])
    teacher = None
    if opt.online_distillation:
        topt = copy.deepcopy(opt)
        topt.ngf, topt.ndf = opt.teacher_ngf, opt.teacher_ndf
        topt.darts_discriminator = False
        topt.online_distillation = False
        topt.generator_only = False
        teacher = cls(topt)
        if opt.model == "pix2pix":
            teacher._defer_G_step = True   # data parallel: the teacher generator's exchange overlaps the student's steps
        teacher.model_train()
        setattr(model, "teacher_model", teacher)
        model.init_distillation()
        teacher.init_distillation()
    model.model_train()
    return model, teacher


def synthetic_batch(model, b, size=None, generator=None, pin=False, z_dim=128):
    """One batch dict as the reference's dataset classes yield it (values uniform in [-1, 1]; z normal)."""
    def u(*shape):
        t = torch.rand(*shape, generator=generator).mul_(2).sub_(1)
        return t.pin_memory() if pin else t

    def n(*shape):
        t = torch.randn(*shape, generator=generator)
        return t.pin_memory() if pin else t

    if model == "srgan":
        hr = size or 96
        return {"lr": n(b, 3, hr // 4, hr // 4), "hr": u(b, 3, hr, hr), "lr_names": "", "hr_names": ""}
    if model == "sagan":
        return {"z": n(b, z_dim), "real_img": u(b, 3, 64, 64), "img_path": ""}
    s = size or 256
    return {"A": u(b, 3, s, s), "B": u(b, 3, s, s), "A_paths": "", "B_paths": ""}


def run_iteration(model, train, val):
    """One pass of train.py:144-151."""
    model.set_input(train)
    model.optimize_parameters()
    if model.opt.darts_discriminator and model.teacher_model is not None:
        model.set_input(val)
        model.clipping_mask_alpha()
        model.optimizer_netD_arch()
