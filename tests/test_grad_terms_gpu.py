"""Tight gradient parity for the discriminators (VERDICT r1 "a bound that wide cannot detect a wrong term").

The discriminator gradient of a GAN step is the sum of a real-batch and a fake-batch term that nearly cancel at the
same weights (measured on the tiny pix2pix teacher: |g_fake| 13.2, |g_real| 14.2, cosine -0.996, |g_fake + g_real|
1.6), so the 0.4-0.7 % bf16 error of each term becomes ~5 % of their sum -- in the reference's own fp32 arithmetic as
soon as weights and forward activations are rounded to bf16 (scripts/exp_bf16_emulation.py; backward-only rounding:
0.35 %).  Hence two checks instead of one wide bound on the sum:

  1. every TERM against the fp32 oracle: rel-L2 <= 2e-2, cosine >= 0.9995 (measured 0.5 %);
  2. the SUM (the gradient the optimizer sees): its error is bounded by the terms' error times the condition number
     kappa = (|g_fake| + |g_real|) / |g_fake + g_real| (22 at the tiny widths, 2.3 at BASELINE's C2 widths where the sum
     itself is at 0.8 %), and it may not exceed 1.5x the deviation the REFERENCE arithmetic itself shows when its
     storage is rounded to bf16 at the same points (oracle/bf16_emulation.py: 6.5 % tiny / 1.1 % C2; the B200 path:
     7.0 % / 0.8 %).  The emulated oracle is a different realisation of the same rounding noise, so B200-vs-emulated
     is reported, not asserted tightly.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}
C2 = {"ngf": 32, "teacher_ngf": 64, "ndf": 128, "teacher_ndf": 128}


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))


def _build(widths):
    from gcc_b200 import factory
    from oracle import gcc_oracle as O
    opt = factory.make_opt("pix2pix", 0, ["--no_dropout"], **widths)
    model, teacher = factory.build_pair(opt)
    S, T = O.build_pair(O.Opt(backbone="unet", direction=opt.direction, **widths))
    for mine, orc in ((model, S), (teacher, T)):
        mine.netG.load_state_dict({k: v.detach() for k, v in orc.G.items()})
        mine.netD.load_state_dict({k: v.detach() for k, v in orc.D.items()})
        with torch.no_grad():
            for i, w in enumerate(orc.transform):
                mine.transform_convs[i].weight.copy_(w.detach())
        mine.sync_weights()
        mine.model_train()
    return model, teacher, S, T, O


def _oracle_terms(O, T):
    """fake-only / real-only / summed discriminator gradients of the (plain-D) teacher at its current weights."""
    out = {}
    names = [k for k, v in T.D.items() if v.dtype == torch.float32 and "running" not in k]
    for which in ("fake", "real", "both"):
        for k in names:
            T.D[k].grad = None
        T._set_grad([T.D[k] for k in names], True)
        loss = 0
        if which in ("fake", "both"):
            pf, _ = T.netD(torch.cat((T.real_A, T.fake_B.detach()), 1))
            loss = loss + O.gan_loss(T.opt.gan_mode, pf, False, True)
        if which in ("real", "both"):
            pr, _ = T.netD(torch.cat((T.real_A, T.real_B), 1))
            loss = loss + O.gan_loss(T.opt.gan_mode, pr, True, True)
        (loss * 0.5).backward()
        out[which] = {k: T.D[k].grad.detach().clone() for k in names if T.D[k].grad is not None}
    return out


def _mine_terms(teacher):
    out = {}
    for which in ("fake", "real", "both"):
        teacher.set_requires_grad(teacher.netD, True)
        teacher.optimizer_D.zero_grad()
        loss = 0
        if which in ("fake", "both"):
            pf, _ = teacher._D(teacher.netD, teacher._fake_AB(teacher.fake_B_nhwc, True))
            loss = loss + teacher._gan(pf, 1)
        if which in ("real", "both"):
            pr, _ = teacher._D(teacher.netD, teacher.real_AB)
            loss = loss + teacher._gan(pr, 0)
        (loss * 0.5).backward()
        torch.cuda.synchronize()
        out[which] = {n: teacher.arena_D.grads[n].detach().float().cpu().clone() for n, _, _ in teacher.arena_D.specs}
        teacher._release_graphs()
    return out


def _flat(d, keys):
    return torch.cat([d[k].flatten() for k in keys])


@pytest.mark.parametrize("widths,batch", [(TINY, 2), (C2, 2)], ids=["tiny", "c2_widths"])
def test_discriminator_gradient_terms(widths, batch):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import bf16_emulation as E
    model, teacher, S, T, O = _build(widths)
    A, B = O.det_image("terms.A", batch, 3, 256, 256), O.det_image("terms.B", batch, 3, 256, 256)
    teacher.set_input({"A": A, "B": B, "A_paths": "", "B_paths": ""})
    teacher.forward()
    mine = _mine_terms(teacher)
    T.set_input(A, B)
    T.forward()
    ref = _oracle_terms(O, T)
    keys = [k for k in ref["both"] if k in mine["both"]]
    rep = {}
    for which in ("fake", "real", "both"):
        a, b = _flat(mine[which], keys), _flat(ref[which], keys)
        rep[which] = (_rel(a, b), _cos(a, b), float(b.norm()))
    # the same with the B200 storage emulated on the oracle side
    with E.emulating(O):
        T2 = O.build_pair(O.Opt(backbone="unet", direction=T.opt.direction, **widths))[1]
        T2.set_input(E.bf(A), E.bf(B))
        T2.forward()
        emu = _oracle_terms(O, T2)
    a, b = _flat(mine["both"], keys), _flat(emu["both"], keys)
    rep["both_vs_bf16_oracle"] = (_rel(a, b), _cos(a, b), float(b.norm()))
    e, b32 = _flat(emu["both"], keys), _flat(ref["both"], keys)
    rep["bf16_oracle_vs_fp32_oracle"] = (_rel(e, b32), _cos(e, b32), float(b32.norm()))
    print({k: tuple(round(x, 5) for x in v) for k, v in rep.items()})
    kappa = (rep["fake"][2] + rep["real"][2]) / rep["both"][2]
    print("condition number (|g_fake| + |g_real|) / |g_fake + g_real| = %.1f" % kappa)
    for which in ("fake", "real"):
        assert rep[which][0] <= 2e-2 and rep[which][1] >= 0.9995, (which, rep[which])
    # the sum against the fp32 oracle: bounded by the terms' error times the condition number ...
    assert rep["both"][0] <= 1.5e-2 * kappa, (rep["both"], kappa)
    # ... and no worse than 1.5x what the reference's own arithmetic does under bf16 storage
    assert rep["both"][0] <= max(2e-2, 1.5 * rep["bf16_oracle_vs_fp32_oracle"][0]), rep
    if kappa < 4:      # BASELINE's widths: the sum itself meets the survey's bf16 bound
        assert rep["both"][0] <= 2e-2 and rep["both"][1] >= 0.9995, rep
