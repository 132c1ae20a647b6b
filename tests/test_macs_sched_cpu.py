"""CPU checks of the host-side pieces added for the pruning front-end and the schedulers:
  * gcc_b200.macs reproduces SURVEY.md 8d's per-net forward GMACs (true multiply-accumulates) and, in thop units, the
    MAC budgets the reference's training scripts quote for its shipped student configurations;
  * the 'plateau' LR policy (utils/util.py:299-300) and the SAGAN arch scheduler (SAGAN.py:348-353)."""
import argparse

import pytest


def test_true_macs_match_survey():
    from gcc_b200 import macs, nets
    g = lambda net, hw=None: macs.count_macs(net, hw) / 1e9
    assert g(nets.UnetGenertor(ngf=64, device="cpu")) == pytest.approx(6.05, abs=0.01)
    assert g(nets.UnetGenertor(ngf=32, device="cpu")) == pytest.approx(1.55, abs=0.01)
    assert g(nets.MobileResnetGenerator(ngf=64, device="cpu")) == pytest.approx(11.07, abs=0.01)
    assert g(nets.MobileResnetGenerator(ngf=32, device="cpu")) == pytest.approx(3.12, abs=0.01)
    assert g(nets.NLayerDiscriminator(input_nc=6, ndf=128, device="cpu")) == pytest.approx(12.57, abs=0.01)


def test_thop_units_match_reference_budgets():
    """scripts/cyclegan/train.sh: --target_budget 2.4 / --target_budget_B 2.7 for the channel lists of
    utils/prune_util.py:120-121 (search tolerance 0.05); scripts/sagan/train.sh: 0.016 for ngf 48 (tolerance 0.001)."""
    from gcc_b200 import macs, nets, prune_util, sagan
    t = lambda net, hw=None: macs.count_macs(net, hw, "thop") / 1e9
    assert abs(t(nets.MobileResnetGenerator(ngf=24, cfg=list(prune_util.CYCLEGAN_CFG_ATOB), device="cpu")) - 2.4) <= 0.05
    assert abs(t(nets.MobileResnetGenerator(ngf=24, cfg=list(prune_util.CYCLEGAN_CFG_BTOA), device="cpu")) - 2.7) <= 0.05
    assert abs(t(sagan.Generator(ngf=48, device="cpu")) - 0.016) <= 0.001
    # a stride-2 transposed conv counts 4x its real work in thop units; U-Net ngf 32: 0.52 (downs) + 4 x 1.03 (ups)
    u = nets.UnetGenertor(ngf=32, device="cpu")
    assert t(u) == pytest.approx(4.65, abs=0.01) and macs.count_macs(u) / 1e9 == pytest.approx(1.55, abs=0.01)


def test_get_flops_parms_input_sizes():
    from gcc_b200 import macs, srgan
    opt = argparse.Namespace(dataroot="./database/sr", image_size=96, load_size=256)
    g, params = macs.get_flops_parms(srgan.Generator(n_channels=24, device="cpu"), opt)
    assert g == pytest.approx(0.2154, abs=1e-3)       # scripts/srgan/train.sh searches for 0.20 +- 0.01 after pruning
    assert params > 0


class _Opt:
    def __init__(self, lr):
        self.param_groups = [{"lr": lr}]


def test_plateau_policy():
    from gcc_b200.pix2pix import get_scheduler
    o = _Opt(1.0)
    s = get_scheduler(o, argparse.Namespace(lr_policy="plateau"))
    with pytest.raises(TypeError):
        s.step()                                       # the reference's update_learning_rate calls it like this
    import torch
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, mode="min", factor=0.2, threshold=0.01, patience=5)
    for m in [1.0, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.499, 0.499]:
        s.step(m)
        ref.step(m)
        assert o.param_groups[0]["lr"] == pytest.approx(ref_opt.param_groups[0]["lr"])
    assert o.param_groups[0]["lr"] < 1.0


def test_unknown_policy_raises():
    from gcc_b200.pix2pix import get_scheduler
    with pytest.raises(NotImplementedError):
        get_scheduler(_Opt(1.0), argparse.Namespace(lr_policy="nope"))
