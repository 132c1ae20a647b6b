"""Prune index selection for the pix2pix generators: which channels survive a threshold.

Reference: Pix2PixModel.scale_prune / norm_prune / resnet_prune / max_min_* and
utils/prune_util.binarysearch_threshold (/root/reference/models/Pix2Pix.py:754-952,
/root/reference/utils/prune_util.py:20-63).  Only the surviving-channel COUNTS matter (the pruned
model is re-initialised), so this is host logic over the state-dict; the comparisons
(`gamma > thr`, `sum|w| > thr`) are evaluated in fp32 on the CPU copy so that the selected indices are
bit-identical to the reference's.
"""
import torch

from .nets import unet_level_prefixes


def _cpu(sd):
    return {k: v.detach().float().cpu() for k, v in sd.items()}


def _l1(w, transposed):
    return torch.sum(torch.abs(w), (0, 2, 3) if transposed else (1, 2, 3))


def _unet_bn_order():
    pre = unet_level_prefixes()
    return [pre[i] + ".2" for i in range(1, 7)] + [pre[7] + ".4"] + [pre[i] + ".6" for i in range(6, 0, -1)]


def _unet_conv_order():
    pre = unet_level_prefixes()
    names = [(pre[0] + ".0", False)] + [(pre[i] + ".1", False) for i in range(1, 8)]
    names += [(pre[7] + ".3", True)] + [(pre[i] + ".5", True) for i in range(6, 0, -1)] + [(pre[0] + ".3", True)]
    return names


def unet_scale_prune_cfg(sd, ngf, threshold):
    sd = _cpu(sd)
    pre = unet_level_prefixes()
    filter_cfgs, channel_cfgs = [ngf], [ngf]
    in_up, n_up = False, 0
    for name in _unet_bn_order():
        if name + ".weight" not in sd:
            continue
        keep = int((sd[name + ".weight"] > threshold).sum())
        filter_cfgs.append(keep)
        if name == pre[7] + ".4":
            in_up = True
            if keep == 0:
                filter_cfgs[-2] = 0
        if in_up:
            n_up += 1
            if filter_cfgs[-2 * n_up] == 0:
                filter_cfgs[-1] = 0
                keep = 0
            channel_cfgs.append(keep + filter_cfgs[-1 - 2 * n_up])
        else:
            channel_cfgs.append(keep)
        if name == pre[6] + ".2":  # the innermost down conv has no BN: keep it at full width unless its parent died
            full = 0 if filter_cfgs[-1] == 0 else ngf * 8
            filter_cfgs.append(full)
            channel_cfgs.append(full)
    return filter_cfgs, channel_cfgs


def unet_norm_prune_cfg(sd, ngf, threshold):
    sd = _cpu(sd)
    filter_cfgs, channel_cfgs = [], []
    n_up = 0
    for name, tr in _unet_conv_order():
        if name + ".weight" not in sd:
            continue
        keep = int((_l1(sd[name + ".weight"], tr) > threshold).sum())
        filter_cfgs.append(keep)
        if tr:
            n_up += 1
            if name != "model.model.3":
                channel_cfgs.append(keep + filter_cfgs[-1 - 2 * n_up])
        else:
            channel_cfgs.append(keep)
    if filter_cfgs[0] == 0:
        filter_cfgs[0] = ngf
        channel_cfgs[0] = ngf
        channel_cfgs[-1] += ngf
    return filter_cfgs, channel_cfgs


_RES_UNPRUNABLE = ["model.26"] + ["model.%d.conv_block.%d.conv.0" % (i, j) for i in range(10, 19) for j in (1, 6)]
_RES_RESIDUAL = ["model.7"] + ["model.%d.conv_block.6.conv.2" % i for i in range(10, 19)]


_RES_CONVT = ("model.19", "model.22")     # the two ConvTranspose2d of a full 9-block MobileResNet generator


def convt_names(net):
    """Names of a net's transposed convs by layer KIND (the reference tests isinstance(m, nn.ConvTranspose2d),
    Pix2Pix.py:791,936): in a generator whose blocks were dropped by an earlier pruning the module indices shift."""
    return tuple(l.tname for l in getattr(net, "_layers", []) if getattr(l, "kind", None) == "convT")


def resnet_prune_cfg(sd, threshold, convt=_RES_CONVT):
    sd = _cpu(sd)
    alive = None
    for n in _RES_RESIDUAL:
        m = _l1(sd[n + ".weight"], False) > threshold
        alive = m if alive is None else (alive | m)
    cfg = []
    for k, v in sd.items():
        if not k.endswith(".weight") or v.dim() != 4:
            continue
        name = k[:-len(".weight")]
        if name in _RES_UNPRUNABLE:
            continue
        if name in _RES_RESIDUAL:
            cfg.append(int(alive.sum()))
        else:
            cfg.append(int((_l1(v, name in convt) > threshold).sum()))
    return cfg


def unet_max_min_bn_scale(sd):
    sd = _cpu(sd)
    pre = unet_level_prefixes()
    prunable = {pre[5] + ".2", pre[6] + ".2", pre[7] + ".4", pre[6] + ".6", pre[5] + ".6"}
    un_max, pr_max, lo = float("inf"), -float("inf"), float("inf")
    for name in _unet_bn_order():
        w = sd[name + ".weight"]
        if name in prunable:
            pr_max = max(float(w.max()), pr_max)
        else:
            un_max = min(float(w.max()), un_max)
        lo = min(float(w.min()), lo)
    return min(pr_max, un_max), lo


def max_min_conv_norm(sd, backbone, convt=_RES_CONVT):
    sd = _cpu(sd)
    pre = unet_level_prefixes()
    prunable = {pre[5] + ".1", pre[6] + ".1", pre[7] + ".1", pre[7] + ".3", pre[6] + ".5", pre[5] + ".5"}
    un_max, pr_max, lo = float("inf"), -float("inf"), float("inf")
    for k, v in sd.items():
        if not k.endswith(".weight") or v.dim() != 4:
            continue
        name = k[:-len(".weight")]
        if name in _RES_UNPRUNABLE:
            continue
        if backbone == "resnet":
            tr = name in convt
        else:
            tr = dict(_unet_conv_order()).get(name, False)
        nrm = _l1(v, tr)
        if name in prunable:
            pr_max = max(float(nrm.max()), pr_max)
        else:
            un_max = min(float(nrm.max()), un_max)
        lo = min(float(nrm.min()), lo)
    if backbone == "resnet":
        return un_max, lo
    return min(pr_max, un_max), lo


def binarysearch_threshold(model, target_budget, count_macs, tolerance=0.1):
    """prune_util.binarysearch_threshold (utils/prune_util.py:20-47) with an injected MAC counter
    (the reference's `thop` is an unvendored, unversioned dependency: parity unpinned)."""
    hi, lo = model.max_min_bn_scale() if model.opt.scale_prune else model.max_min_conv_norm()
    while hi > lo:
        mid = (hi + lo) / 2
        pruned = model.prune(mid)
        budget = count_macs(pruned)
        if abs(target_budget - budget) <= tolerance:
            return mid
        if target_budget - budget > tolerance:
            hi = mid
        else:
            lo = mid
    raise NotImplementedError("No appropriate threshold found")


# ---- CycleGAN (models/CycleGAN.py:803-885): the residual stream is pruned by the MEAN norm over its ten convs
def _cyclegan_residual_mean(sd):
    total = None
    for n in _RES_RESIDUAL:
        v = _l1(sd[n + ".weight"], False)
        total = v.clone() if total is None else total + v
    return total / len(_RES_RESIDUAL)


def cyclegan_prunenet_cfg(sd, threshold):
    sd = _cpu(sd)
    alive = _cyclegan_residual_mean(sd) > threshold
    cfg = []
    for k, v in sd.items():
        if not k.endswith(".weight") or v.dim() != 4:
            continue
        name = k[:-len(".weight")]
        if name in _RES_UNPRUNABLE:
            continue
        if name in _RES_RESIDUAL:
            cfg.append(int(alive.sum()))
        else:
            cfg.append(int((_l1(v, name in ("model.19", "model.22")) > threshold).sum()))
    return cfg


def cyclegan_max_min_conv_norm(sd):
    sd = _cpu(sd)
    mean = _cyclegan_residual_mean(sd)
    un_max, lo = float("inf"), float("inf")
    for k, v in sd.items():
        if not k.endswith(".weight") or v.dim() != 4:
            continue
        name = k[:-len(".weight")]
        if name in _RES_UNPRUNABLE:
            continue
        nrm = mean if name in _RES_RESIDUAL else _l1(v, name in ("model.19", "model.22"))
        un_max = min(float(nrm.max()), un_max)
        lo = min(float(nrm.min()), lo)
    return un_max, lo
