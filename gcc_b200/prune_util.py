"""utils/prune_util.py of the reference (get_flops_parms :6-18, binarysearch_threshold :20-47, prune :49-63,
cyclegan_binarysearch_cfg :65-108, cyclegan_prune :110-126) for the gcc_b200 models: same function names, argument
meaning and exceptions, with `gcc_b200.macs` in place of the unvendored `thop` (MAC budgets are in thop units)."""
from . import macs
from .options import get_model_class


def get_flops_parms(model, device=None, opt=None, verbose=False):
    """(GMACs in thop units, M parameters) of a generator, on the input the reference profiles it with."""
    return macs.get_flops_parms(model, opt, convention="thop")


def _tolerance(opt):
    root = str(opt.dataroot)
    if "sr" in root:
        return 0.01
    if "celeb" in root or "church" in root:
        return 0.001
    return 0.1


def binarysearch_threshold(model, target_budget):
    if model.opt.scale_prune:
        max_scale, min_scale = model.max_min_bn_scale()
    else:
        max_scale, min_scale = model.max_min_conv_norm()
    tolerance = _tolerance(model.opt)
    while max_scale > min_scale:
        mid_scale = (max_scale + min_scale) / 2
        pruned_model = model.prune(mid_scale)
        budget, _ = get_flops_parms(pruned_model.netG, pruned_model.device, pruned_model.opt)
        if abs(target_budget - budget) <= tolerance:
            return mid_scale
        elif target_budget - budget > tolerance:
            max_scale = mid_scale
        else:
            min_scale = mid_scale
    raise NotImplementedError("No appropriate threshold found")


def prune(model, opt, logger):
    if opt.target_budget is None:
        raise NotImplementedError("the target budget must be exist!!!")
    if opt.pretrain_path is None:
        raise NotImplementedError("the pretrain path must be exist!!!")
    model.load_models(opt.pretrain_path, load_discriminator=False)
    threshold = binarysearch_threshold(model, opt.target_budget)
    pruned_model = model.prune(threshold, lottery_path=opt.lottery_path)
    filter_cfg, channel_cfg = pruned_model.get_cfg()
    g, params = get_flops_parms(pruned_model.netG, pruned_model.device, pruned_model.opt)
    logger.info(filter_cfg)
    logger.info(channel_cfg)
    logger.info("MACs:%.7f G  |  Params:%.4f M" % (g, params))
    return pruned_model


def cyclegan_binarysearch_cfg(model, target_budget, target_budget_B):
    tolerance = 0.05
    model_class = get_model_class(model.opt)
    final = {}
    for key, net, target, kw in (("A", model.netG_A, target_budget, "cfg_AtoB"), ("B", model.netG_B, target_budget_B, "cfg_BtoA")):
        max_scale, min_scale = model.max_min_conv_norm(net)
        while max_scale > min_scale:
            mid_scale = (max_scale + min_scale) / 2
            cfg = model.get_prunenet_cfg(net, mid_scale)
            pruned_model = model_class(model.opt, **{kw: cfg})
            budget, _ = get_flops_parms(getattr(pruned_model, "netG_" + key), pruned_model.device, pruned_model.opt)
            if abs(target - budget) <= tolerance:
                final[key] = cfg
                break
            elif target - budget > tolerance:
                max_scale = mid_scale
            else:
                min_scale = mid_scale
    if "A" not in final or "B" not in final:
        raise NotImplementedError("No appropriate threshold found")
    return final["A"], final["B"]


def cyclegan_prune(model, opt, logger):
    """The reference overrides the searched cfgs with two hard-coded lists (utils/prune_util.py:120-121); kept."""
    if opt.target_budget is None or opt.target_budget_B is None:
        raise NotImplementedError("the target budget must be exist!!!")
    if opt.pretrain_path is None:
        raise NotImplementedError("the pretrain path must be exist!!!")
    model.load_models(opt.pretrain_path, load_discriminator=False)
    model_class = get_model_class(model.opt)
    cyclegan_binarysearch_cfg(model, opt.target_budget, opt.target_budget_B)
    pruned_model = model_class(model.opt, cfg_AtoB=list(CYCLEGAN_CFG_ATOB), cfg_BtoA=list(CYCLEGAN_CFG_BTOA))
    cfg_AtoB, cfg_BtoA = pruned_model.get_cfg()
    logger.info(cfg_AtoB)
    logger.info(cfg_BtoA)
    for tag, net in (("AtoB", pruned_model.netG_A), ("BtoA", pruned_model.netG_B)):
        g, params = get_flops_parms(net, pruned_model.device, pruned_model.opt)
        logger.info("%s MACs:%.7f G  |  Params:%.4f M" % (tag, g, params))
    return pruned_model


# the channel lists the reference trains its CycleGAN students with (utils/prune_util.py:120-121)
CYCLEGAN_CFG_ATOB = (24, 48, 86, 72, 86, 47, 86, 44, 86, 43, 86, 43, 86, 29, 86, 30, 86, 37, 86, 36, 86, 48, 24)
CYCLEGAN_CFG_BTOA = (24, 48, 96, 91, 96, 73, 96, 62, 96, 61, 96, 74, 96, 54, 96, 51, 96, 58, 96, 81, 96, 48, 24)
