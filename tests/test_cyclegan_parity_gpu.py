"""CycleGAN GCC iteration on the B200 (gcc_b200.cyclegan.MobileCycleGANModel) against the CPU oracle
(oracle.gcc_oracle.CycleGANOracle, pinned to the reference by tests/golden/cyclegan_tiny.pt).

Stated tolerances (bf16 activations and activation-gradients vs the fp32 oracle):
  losses ................... 5 % (+1e-2 abs)            measured <= 0.4 %
  single-pass images ....... rel L2 <= 5e-2             measured 2.2-2.9 % (41 bf16-rounded InstanceNorm stages)
  cycle reconstructions .... rel L2 <= 0.15             measured 9-11 % (two generator passes chained)
  discriminator gradients .. rel L2 <= 0.15, cos >= 0.99
  generator gradients ...... rel L2 <= 0.5,  cos >= 0.9  measured 0.38 / 0.925
  gate masks ............... bit exact
The generator-gradient tolerance is wide on purpose and is a property of bf16 storage, not of the kernels: at
initialisation the lsgan discriminator output is almost constant, so the gradient entering every InstanceNorm
backward is a large common-mode value plus a small signal, and `dy - mean(dy)` cancels most of the 8-bit mantissa
(scripts/debug_cycle_grads.py: the identity-loss term alone, one generator pass and no discriminator, is at
rel 0.087 / cos 0.996; the GAN term through the InstanceNorm discriminator is at 0.39 / 0.925)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}
CFG = [8, 16, 29, 21, 29, 17, 29, 30, 29, 11, 29, 25, 29, 32, 29, 9, 29, 27, 29, 19, 29, 13, 7]


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


C3 = {"ngf": 24, "teacher_ngf": 64, "ndf": 64, "teacher_ndf": 64}
C3_A = [24, 48, 86, 72, 86, 47, 86, 44, 86, 43, 86, 43, 86, 29, 86, 30, 86, 37, 86, 36, 86, 48, 24]
C3_B = [24, 48, 96, 91, 96, 73, 96, 62, 96, 61, 96, 74, 96, 54, 96, 51, 96, 58, 96, 81, 96, 48, 24]


@pytest.mark.parametrize("case", ["tiny", "c3_widths"])
def test_cyclegan_iteration_matches_oracle(case):
    """tiny: widths 8 / 16, one pruned generator (the bounds of the module docstring).  c3_widths: BASELINE configs[2]'s
    real widths (student ngf 24 with the channel lists of utils/prune_util.py:120-121, teacher ngf 64, ndf 64) where the
    near-cancellation of the tiny nets is gone: generators rel-L2 <= 5e-2 / cos >= 0.998, discriminators <= 3e-2 /
    0.9995, reconstructions <= 5e-2."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    global TINY, CFG
    widths, cfg_a, cfg_b = (TINY, CFG, None) if case == "tiny" else (C3, C3_A, C3_B)
    from gcc_b200 import options
    from gcc_b200.cyclegan import MobileCycleGANModel, build_cycle_teacher
    from oracle import gcc_oracle as O
    opt = options.parse(["--dataroot", "x/horse2zebra", "--model", "cyclegan", "--darts_discriminator",
                         "--online_distillation", "--lambda_content", "0.01", "--lambda_gram", "10", "--gpu_ids", "0"])
    assert opt.gan_mode == "lsgan" and opt.lambda_L1 == 0.0
    for k, v in widths.items():
        setattr(opt, k, v)
    model = MobileCycleGANModel(opt, cfg_AtoB=cfg_a, cfg_BtoA=cfg_b)
    teacher = build_cycle_teacher(model, opt)
    S, T = O.build_cycle_pair(O.CycleOpt(direction=opt.direction, **widths), cfg_a, cfg_b)
    for mine, orc in ((model, S), (teacher, T)):
        for k in "AB":
            getattr(mine, "netG_" + k).load_state_dict({n: v.detach() for n, v in orc.G[k].items()})
            getattr(mine, "netD_" + k).load_state_dict({n: v.detach() for n, v in orc.D[k].items()})
            with torch.no_grad():
                for i, w in enumerate(orc.transform[k]):
                    getattr(mine, "transform_%s_convs" % k)[i].weight.copy_(w.detach())
        mine.sync_weights()
        mine.model_train()
    size = 128
    A, B = O.det_image("cyc.A", 1, 3, size, size), O.det_image("cyc.B", 1, 3, size, size)
    S.set_input(A, B)
    S.optimize_parameters()
    model.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
    model.optimize_parameters()
    torch.cuda.synchronize()
    rep = {}
    for n in ("fake_A", "fake_B", "rec_A", "rec_B", "idt_A", "idt_B"):
        rep[n] = _rel(getattr(model, n).cpu(), getattr(S, n).detach())
    rep["Tfake_B"] = _rel(teacher.fake_B.cpu(), T.fake_B.detach())

    worst = {}

    def grads(arena, named, prefix, tag=None):
        a, b, per = [], [], []
        for n, v in named.items():
            if v.dtype == torch.float32 and v.grad is not None and prefix + n in arena.grads and not n.endswith(".bias"):
                ga, gb = arena.grads[prefix + n].detach().float().cpu().flatten(), v.grad.flatten()
                a.append(ga)
                b.append(gb)
                per.append((round(_rel(ga, gb), 4), round(_cos(ga, gb), 4), n, float(gb.norm())))
        if tag:
            worst[tag] = sorted(per, reverse=True)[:8]
        return torch.cat(a), torch.cat(b)

    for tag, mine, orc in (("S", model, S), ("T", teacher, T)):
        for k in "AB":
            a, b = grads(mine.arena_G, orc.G[k], k + ".", "%s.G_%s" % (tag, k))
            rep["%s.G_%s.grad.rel" % (tag, k)], rep["%s.G_%s.grad.cos" % (tag, k)] = _rel(a, b), _cos(a, b)
            a, b = grads(mine.arena_D, {n: v for n, v in orc.D[k].items() if not n.endswith("alpha")}, k + ".")
            rep["%s.D_%s.grad.rel" % (tag, k)], rep["%s.D_%s.grad.cos" % (tag, k)] = _rel(a, b), _cos(a, b)
    losses = {}
    mine_l = {n: float(getattr(model, "loss_" + n).detach()) for n in ("D_A", "G_A", "cycle_A", "idt_A", "D_B", "G_B",
                                                                       "cycle_B", "idt_B", "content_A", "content_B",
                                                                       "gram_A", "gram_B")}
    for n, v in mine_l.items():
        losses[n] = (v, float(getattr(S, "loss_" + n)))
    vA, vB = O.det_image("cyc.vA", 1, 3, size, size), O.det_image("cyc.vB", 1, 3, size, size)
    S.set_input(vA, vB)
    S.clipping_mask_alpha()
    S.optimizer_netD_arch()
    model.set_input({"A": vA, "B": vB, "A_paths": "", "B_paths": ""})
    model.clipping_mask_alpha()
    model.optimizer_netD_arch()
    torch.cuda.synchronize()
    got = model.get_current_losses()
    exp = S.get_current_losses()
    assert list(got) == list(exp)
    for n in ("D_arch_A", "D_arch_B", "teacher_netD_A_arch_diff", "teacher_netD_B_arch_diff"):
        losses[n] = (got[n], exp[n])
    masks_ok = all(torch.equal(m.cpu(), om) for m, om in
                   zip(model.netD_A.get_current_masks() + model.netD_B.get_current_masks(), S.current_masks()))
    rep["losses"] = {k: {"b200": a, "oracle": b} for k, (a, b) in losses.items()}
    print("WORST", json.dumps(worst, indent=0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "step_parity_cyclegan%s.json" % ("" if case == "tiny" else "_c3")), "w"),
              indent=1)
    print(json.dumps(rep, indent=1))
    bad = []
    for k, v in rep.items():
        if k == "losses":
            continue
        if case == "tiny":
            lim_cos = 0.9 if ".G_" in k else 0.99
            lim_rel = 0.5 if ".G_" in k else 0.15
            lim_img = 0.15 if k.startswith("rec_") else 5e-2
        else:
            lim_cos = 0.998 if ".G_" in k else 0.9995
            lim_rel = 5e-2 if ".G_" in k else 3e-2
            lim_img = 5e-2
        if k.endswith(".cos"):
            if v < lim_cos:
                bad.append((k, v))
        elif k.endswith(".rel"):
            if v > lim_rel:
                bad.append((k, v))
        elif v > lim_img:
            bad.append((k, v))
    for k, (a, b) in losses.items():
        if abs(a - b) > 5e-2 * abs(b) + 1e-2:
            bad.append(("loss." + k, a, b))
    assert masks_ok, "gate masks differ from the oracle"
    assert not bad, bad
