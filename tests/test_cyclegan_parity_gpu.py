"""CycleGAN GCC iteration on the B200 (gcc_b200.cyclegan.MobileCycleGANModel) against the CPU oracle
(oracle.gcc_oracle.CycleGANOracle, pinned to the reference by tests/golden/cyclegan_tiny.pt).

Stated tolerances (bf16 activations and activation-gradients vs the fp32 oracle):
  losses ................... 5 % (+1e-2 abs)            measured <= 0.9 %
  single-pass images ....... rel L2 <= 5e-2             measured 2.2-2.9 % (41 bf16-rounded InstanceNorm stages)
  cycle reconstructions .... rel L2 <= 0.15             measured 9-11 % (two generator passes chained)
  discriminator gradients .. rel L2 <= 0.15, cos >= 0.99  (measured 1.2-13 %; real/fake cancellation as in pix2pix)
  generator gradients ...... NO fixed bound: calibrated against the reference arithmetic under bf16 storage (below)
  gate masks ............... bit exact

Generator gradients.  The CycleGAN generator loss is dominated by L1 terms (cycle, identity: lambda 10 / 5), whose
gradient is sign(rec - real) / n.  A forward pass stored in bf16 moves rec by 2-10 % (41 InstanceNorm stages per
generator, two generators chained), which flips the sign wherever |rec - real| is below that error: a few per cent of
the pixels, each flip changing the upstream gradient by 2 / n.  sqrt(4 * 3 %) ~ 35 % of the gradient norm is therefore
the sensitivity of the REFERENCE'S OWN arithmetic to bf16 storage -- measured on the CPU oracle with its storage rounded
at the points where the B200 path rounds (oracle/bf16_emulation.py): weights only 30 % / cos 0.955, forward
activations only 34-38 %, backward gradients only 0.9 %, all three 38-39 % / 0.925; the B200 path measures 33-40 %.
It is not a backward-precision effect (an fp32 gradient path would not change it) and it is the same at the real
widths of BASELINE configs[2] (case c3_widths).  The test therefore requires the B200 generator gradients to be no
further from the fp32 oracle than 1.3x the bf16-emulated oracle is (+ 3 % absolute), with a cosine no more than 0.03
below it -- a wrong or mis-scaled loss term (e.g. a missing lambda) would exceed that by far -- and keeps every loss,
image and discriminator-gradient bound fixed."""
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
    """tiny: widths 8 / 16, one pruned generator.  c3_widths: BASELINE configs[2]'s real widths (student ngf 24 with the
    channel lists of utils/prune_util.py:120-121, teacher ngf 64, ndf 64).  Same bounds (module docstring)."""
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
    # the reference arithmetic with bf16 storage emulated: the calibration of the generator-gradient bound
    from oracle import bf16_emulation as E
    with E.emulating(O):
        S16, T16 = O.build_cycle_pair(O.CycleOpt(direction=opt.direction, **widths), cfg_a, cfg_b)
        S16.set_input(E.bf(A), E.bf(B))
        S16.optimize_parameters()
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

    emu = {}
    for tag, mine, orc, o16 in (("S", model, S, S16), ("T", teacher, T, T16)):
        for k in "AB":
            a, b = grads(mine.arena_G, orc.G[k], k + ".", "%s.G_%s" % (tag, k))
            rep["%s.G_%s.grad.rel" % (tag, k)], rep["%s.G_%s.grad.cos" % (tag, k)] = _rel(a, b), _cos(a, b)
            e = torch.cat([v.grad.flatten() for n, v in o16.G[k].items() if v.dtype == torch.float32 and v.grad is not None
                           and k + "." + n in mine.arena_G.grads and not n.endswith(".bias")])
            emu["%s.G_%s" % (tag, k)] = (_rel(e, b), _cos(e, b))
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
    rep["bf16_oracle_vs_fp32_oracle"] = {k: {"rel": v[0], "cos": v[1]} for k, v in emu.items()}
    print("WORST", json.dumps(worst, indent=0))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "step_parity_cyclegan%s.json" % ("" if case == "tiny" else "_c3")), "w"),
              indent=1)
    print(json.dumps(rep, indent=1))
    bad = []
    for k, v in rep.items():
        if k in ("losses", "bf16_oracle_vs_fp32_oracle"):
            continue
        lim_cos, lim_rel = 0.99, 0.15
        lim_img = 0.15 if k.startswith("rec_") else 5e-2
        if ".G_" in k:       # calibrated: see the module docstring
            e_rel, e_cos = emu[k.split(".grad.")[0]]
            lim_rel, lim_cos = 1.3 * e_rel + 0.03, e_cos - 0.03
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
