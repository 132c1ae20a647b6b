"""Option namespace for the GCC step: the flags of /root/reference/options/options.py:6-126 that the
pix2pix model reads, with the same names, defaults and the per-model overrides of ``parse()``
(options.py:164-191, 226-228).  ``parse(argv)`` returns the same kind of Namespace train.py passes
to the model classes."""
import argparse

_FLAGS = [
    # name, type, default
    ("dataroot", str, None), ("name", str, "default"), ("gpu_ids", str, "0"), ("checkpoints_dir", str, "./experiments"),
    ("phase", str, "train"), ("load_path", str, None), ("pretrain_path", str, None), ("model", str, "pix2pix"),
    ("input_nc", int, 3), ("output_nc", int, 3), ("ngf", int, 64), ("pretrain_ngf", int, 64), ("ndf", int, 128),
    ("backbone", str, "unet"), ("num_downs", int, 8), ("dataset_mode", str, "aligned"), ("direction", str, "AtoB"),
    ("num_threads", int, 8), ("batch_size", int, 1), ("load_size", int, 286), ("crop_size", int, 256),
    ("max_dataset_size", float, float("inf")), ("preprocess", str, "resize_and_crop"), ("print_freq", int, 500),
    ("save_epoch_freq", int, 1), ("epoch_count", int, 1), ("n_epochs", int, 100), ("n_epochs_decay", int, 150),
    ("lr", float, 0.0002), ("gan_mode", str, "hinge"), ("pool_size", int, 100), ("lr_policy", str, "linear"),
    ("lr_decay_iters", int, 50), ("lambda_A", float, 10.0), ("lambda_B", float, 10.0), ("lambda_identity", float, 0.5),
    ("lambda_L1", float, 0.0), ("ntest", float, float("inf")), ("aspect_ratio", float, 1.0),
    ("lambda_weight", float, 0.0), ("lambda_scale", float, 0.0), ("target_budget", float, None),
    ("target_budget_B", float, None), ("lottery_path", str, None), ("arch_lr", float, 1e-4),
    ("lambda_alpha", float, 0.01), ("ema_beta", float, 1.0), ("threshold", float, 0.5), ("distillation_path", str, None),
    ("lambda_content", float, 0.0), ("lambda_gram", float, 0.0), ("teacher_ngf", int, 64), ("teacher_ndf", int, 64),
    ("initial_path", str, None), ("teacher_initial_path", str, None), ("z_dim", int, 128),
    ("lambda_SR_adversarial", float, 1e-3), ("lambda_SR_content", float, 0.0), ("lambda_SR_perceptual", float, 1.0),
    ("image_size", int, 96),
]
_SWITCHES = ["no_dropout", "serial_batches", "no_flip", "split_dataset", "scale_prune", "norm_prune",
             "darts_discriminator", "arch_lr_step", "adaptive_ema", "regular", "arch_base_loss", "only_arch_base",
             "normalize_arch", "clear_arch", "online_distillation", "normal_distillation", "center_crop",
             "generator_only"]
# gcc_b200 extensions (not reference flags): --sync_bn = global-batch parity under data parallel (synchronised
# BatchNorm statistics and loss partial sums all-reduced before sqrt / abs)
_EXT_SWITCHES = ["sync_bn"]


def make_parser():
    p = argparse.ArgumentParser("GAN-Compression (gcc_b200)")
    for name, typ, default in _FLAGS:
        p.add_argument("--" + name, type=typ, default=default)
    for name in _SWITCHES + _EXT_SWITCHES:
        p.add_argument("--" + name, action="store_true")
    return p


def parse(argv=None):
    opt = make_parser().parse_args(argv)
    opt.gpu_ids = [int(s) for s in str(opt.gpu_ids).split(",") if int(s) >= 0]
    root = opt.dataroot or ""
    if opt.model in ("pix2pix", "newpix2pix"):
        opt.norm = "batch"
        opt.dataset_mode = "aligned"
        opt.no_flip = True
        opt.load_size = 256
        opt.pool_size = 0
        opt.teacher_ndf = 128
        opt.lambda_L1 = 100.0
        if "cityscapes" in root:
            opt.direction, opt.save_epoch_freq, opt.n_epochs, opt.n_epochs_decay, opt.print_freq = "BtoA", 5, 100, 150, 100
        if "edges2shoes" in root:
            opt.batch_size, opt.n_epochs, opt.n_epochs_decay = 4, 10, 30
        if "maps" in root:
            opt.n_epochs, opt.direction, opt.no_flip, opt.load_size = 100, "BtoA", False, 286
            opt.n_epochs_decay, opt.save_epoch_freq, opt.print_freq, opt.lambda_L1 = 200, 5, 100, 10.0
    elif "cyclegan" in opt.model:
        opt.dataset_mode = "unaligned"
        opt.gan_mode = "lsgan"
        opt.n_epochs, opt.n_epochs_decay, opt.print_freq = 100, 100, 100
    elif opt.model == "srgan":                      # options.py:192-205
        opt.dataset_mode = "sr"
        opt.gan_mode = "vanilla"
        opt.lr = 1e-4
        opt.batch_size = 16
        opt.n_epochs_decay = 0
        if opt.generator_only:
            opt.n_epochs = 130
        else:
            opt.n_epochs = 30
            opt.lr_policy = "step"
            opt.lr_decay_iters = opt.n_epochs // 2
    elif opt.model == "sagan":                      # options.py:206-218
        opt.dataset_mode = "sa"
        opt.crop_size = 64
        opt.batch_size = 64
        opt.lr = 1e-4
        opt.n_epochs_decay = 0
        opt.save_epoch_freq = 5
        if "church" in root:
            opt.n_epochs, opt.center_crop = 300, False
        else:
            opt.n_epochs, opt.center_crop = 100, True
    else:
        raise NotImplementedError("%s not implemented" % opt.model)
    if opt.lambda_weight > 0 or opt.lambda_scale > 0:
        opt.n_epochs //= 10
        opt.n_epochs_decay //= 10
    return opt


def get_model_class(opt):
    """models/__init__.py:3-14."""
    if opt.model == "pix2pix":
        from .pix2pix import Pix2PixModel
        return Pix2PixModel
    if opt.model == "cyclegan":
        from .cyclegan import MobileCycleGANModel
        return MobileCycleGANModel
    if opt.model == "srgan":
        from .srgan import SRGAN
        return SRGAN
    if opt.model == "sagan":
        from .sagan import SAGANModel
        return SAGANModel
    raise NotImplementedError("%s not implemented" % opt.model)
