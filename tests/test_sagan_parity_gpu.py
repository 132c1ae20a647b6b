"""SAGAN GCC iteration on the B200 (gcc_b200.sagan.SAGANModel) against the CPU oracle (oracle.sagan_oracle, pinned to
the reference by tests/golden/sagan_tiny.pt), plus the SAGAN-only kernels against torch fp32.

Stated tolerances (bf16 activations and activation-gradients vs the fp32 oracle):
  losses ................................ 5 % (+2e-3 abs)
  fake images ........................... rel L2 <= 5e-2
  discriminator gradients ............... global rel L2 <= 0.15, cosine >= 0.99   (measured 5-6 % / 0.998)
  generator gradients ................... global rel L2 <= 0.25, cosine >= 0.97   (measured student 0.156 / 0.988,
                                          teacher 0.063 / 0.998; exactly-zero-gradient parameters excluded: conv biases
                                          in front of a BatchNorm, attention key biases)
      The generator gradient runs through the whole discriminator (4 LeakyReLU(0.1) stages) and the generator
      (4 ReLU stages).  bf16 storage flips the branch of the ~0.4 % of pre-activations that lie within rounding of
      zero; each flip changes that element's gradient by 10x (LeakyReLU 0.1) or switches it off (ReLU), i.e. ~6 %
      relative error per stage, adding incoherently to ~15 % over eight stages; scripts/debug_sagan_grads.py shows the
      same figure for the GAN term alone and for every tensor (no single layer stands out, cosines 0.98-0.995).
  spectral-norm vectors u, v after the step ... rel L2 <= 2e-2
  gate masks ............................ bit exact
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = {"ngf": 16, "teacher_ngf": 32, "ndf": 16, "teacher_ndf": 32}


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda")


def test_spectral_norm_and_attention_kernels(cuda):
    from gcc_b200 import ops
    from gcc_b200._lib import call
    from gcc_b200.sagan import AttnFn
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(0)
    # power iteration + sigma + backward against autograd on the same formula
    h, wd = 24, 80
    W = torch.randn(h, wd)
    u0 = torch.randn(h); u0 /= u0.norm()
    v0 = torch.randn(wd); v0 /= v0.norm()
    Wd, u, v = W.cuda(), u0.cuda().clone(), v0.cuda().clone()
    t = torch.empty(h, device=cuda); sigma = torch.empty(1, device=cuda); scr = torch.empty(2, device=cuda)
    call("gcc_spectral_norm_fwd", Wd.data_ptr(), u.data_ptr(), v.data_ptr(), h, wd, t.data_ptr(), sigma.data_ptr(),
         scr.data_ptr(), st)
    v1 = W.t() @ u0; v1 = v1 / (v1.norm() + 1e-12)
    u1 = W @ v1; u1 = u1 / (u1.norm() + 1e-12)
    Wr = W.clone().requires_grad_(True); ur = u1.clone().requires_grad_(True); vr = v1.clone().requires_grad_(True)
    sig = ur.dot(Wr.mv(vr))
    assert torch.allclose(u.cpu(), u1, atol=1e-5) and torch.allclose(v.cpu(), v1, atol=1e-5)
    assert float(sigma) == pytest.approx(float(sig), rel=1e-5)
    dWeff = torch.randn(h, wd)
    ((Wr / sig) * dWeff).sum().backward()
    dW = torch.zeros(h, wd, device=cuda); du = torch.zeros(h, device=cuda); dv = torch.zeros(wd, device=cuda)
    call("gcc_spectral_norm_bwd", dWeff.cuda().data_ptr(), Wd.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(),
         t.data_ptr(), h, wd, dW.data_ptr(), du.data_ptr(), dv.data_ptr(), scr.data_ptr(), st)
    assert _rel(dW.cpu(), Wr.grad) < 1e-4 and _rel(du.cpu(), ur.grad) < 1e-4 and _rel(dv.cpu(), vr.grad) < 1e-4
    # attention core forward / backward (batched tcgen05 GEMMs + row softmax; L must be a multiple of 8: the
    # reference's attention maps are 16 / 64 / 256 / 1024 positions).  Cases: L < one 128-pixel tile, and L spanning tiles.
    for (n, hh, ww, c, d) in ((2, 6, 4, 24, 3), (3, 16, 24, 40, 5)):
        _attn_case(cuda, AttnFn, n, hh, ww, c, d)


def _attn_case(cuda, AttnFn, n, hh, ww, c, d):
    L = hh * ww
    q = torch.randn(n, hh, ww, 8, device=cuda).to(torch.bfloat16); q[..., d:] = 0
    k = torch.randn(n, hh, ww, 8, device=cuda).to(torch.bfloat16); k[..., d:] = 0
    vv = torch.randn(n, hh, ww, c, device=cuda).to(torch.bfloat16)
    qr, kr, vr = (x.float().cpu().reshape(n, L, -1).requires_grad_(True) for x in (q, k, vv))
    q.requires_grad_(True); k.requires_grad_(True); vv.requires_grad_(True)
    out = AttnFn.apply(q, k, vv, d, c)
    att = torch.softmax(qr[..., :d] @ kr[..., :d].transpose(1, 2), dim=-1)
    ref = att @ vr
    assert _rel(out.float().cpu().reshape(n, L, c), ref.detach()) < 1e-2
    g = torch.randn(n, hh, ww, c, device=cuda).to(torch.bfloat16)
    out.backward(g)
    ref.backward(g.float().cpu().reshape(n, L, c))
    assert _rel(vv.grad.float().cpu().reshape(n, L, c), vr.grad) < 2e-2
    assert _rel(q.grad.float().cpu().reshape(n, L, 8)[..., :d], qr.grad[..., :d]) < 3e-2
    assert _rel(k.grad.float().cpu().reshape(n, L, 8)[..., :d], kr.grad[..., :d]) < 3e-2


def test_sagan_nets_match_oracle(cuda):
    from gcc_b200 import ops
    from gcc_b200.sagan import Discriminator, Generator
    from oracle import gcc_oracle as O
    from oracle import sagan_oracle as SA
    PG = SA.make_params(SA.generator_shapes(16), "t.netG.")
    G = Generator(ngf=16, device=cuda)
    G.load_state_dict({k: v.detach().clone() for k, v in PG.items()})
    G.train()
    z = O.det_normal("t.z", (4, 128))
    out = G(ops.to_nhwc(z.cuda().reshape(4, 128, 1, 1)))
    ref, taps = SA.generator_forward(PG, z)
    assert _rel(ops.to_nchw(out.detach(), 3).cpu(), ref.detach()) < 5e-2
    sd = G.state_dict()
    for k in PG:
        if k.endswith("weight_u") or k.endswith("weight_v"):
            assert _rel(sd[k].cpu(), PG[k].detach()) < 1e-3, k           # the power iteration ran in place, same result
    x = O.det_image("t.x", 4, 3, 64, 64)
    PD2 = SA.make_params(SA.disc_shapes(16, False), "t.netD.")
    D2 = Discriminator(ndf=16, device=cuda)
    D2.load_state_dict({k: v.detach().clone() for k, v in PD2.items()})
    pred = D2(ops.to_nhwc(x.cuda()))
    ref, _ = SA.disc_forward(PD2, x)
    got = pred.detach().float().cpu().reshape(4, 8)[:, 0]
    assert torch.allclose(got, ref.detach(), rtol=5e-2, atol=5e-3), (got, ref)


def test_sagan_iteration_matches_oracle(cuda):
    from gcc_b200 import options
    from gcc_b200.sagan import SAGANModel, build_sa_teacher
    from oracle import gcc_oracle as O
    from oracle import sagan_oracle as SA
    opt = options.parse(["--dataroot", "x/celeb", "--model", "sagan", "--darts_discriminator", "--online_distillation",
                         "--lambda_content", "1e-3", "--lambda_gram", "1e-1", "--gpu_ids", "0"])
    assert opt.gan_mode == "hinge" and opt.lr == 1e-4 and opt.crop_size == 64
    for k, v in TINY.items():
        setattr(opt, k, v)
    model = SAGANModel(opt)
    teacher = build_sa_teacher(model, opt)
    S, T = SA.build_sa_pair(SA.SAOpt(**TINY))
    for mine, orc in ((model, S), (teacher, T)):
        mine.netG.load_state_dict({n: v.detach().clone() for n, v in orc.G.items()})
        mine.netD.load_state_dict({n: v.detach().clone() for n, v in orc.D.items()})
        with torch.no_grad():
            for i, w in enumerate(orc.transform):
                mine.transform_convs[i].weight.copy_(w.detach())
        mine.sync_weights()
        mine.model_train()
    b = 4
    z, real = O.det_normal("sap.z", (b, 128)), O.det_image("sap.real", b, 3, 64, 64)
    S.set_input(z, real)
    S.optimize_parameters()
    model.set_input({"z": z, "real_img": real, "img_path": ""})
    model.optimize_parameters()
    torch.cuda.synchronize()
    rep = {"fake_img": _rel(model.fake_img.cpu(), S.fake_img.detach()),
           "Tfake_img": _rel(teacher.fake_img.cpu(), T.fake_img.detach())}

    worst = {}

    def grads(mine_net_arenas, named, tag=None):
        a, bb, per = [], [], []
        for n, v in named.items():
            if v.dtype != torch.float32 or v.grad is None or n.endswith("alpha"):
                continue
            if float(v.grad.norm()) / max(v.grad.numel(), 1) ** 0.5 < 1e-7:
                continue                              # exactly-zero true gradient (fp32 noise in the oracle)
            if n.endswith("weight_v") or n.endswith("weight_u"):
                continue                              # checked through the stepped values below
            for ar in mine_net_arenas:
                if n in ar.grads:
                    a.append(ar.grads[n].detach().float().cpu().flatten())
                    bb.append(v.grad.flatten())
                    per.append((round(_rel(a[-1], bb[-1]), 4), round(_cos(a[-1], bb[-1]), 4), n, float(bb[-1].norm())))
                    break
        if tag:
            worst[tag] = sorted(per, reverse=True)[:10]
        return torch.cat(a), torch.cat(bb)

    for tag, mine, orc in (("S", model, S), ("T", teacher, T)):
        a, bb = grads([mine.arena_G, mine.arena_Gs], orc.G, tag + ".G")
        rep[tag + ".G.grad.rel"], rep[tag + ".G.grad.cos"] = _rel(a, bb), _cos(a, bb)
        a, bb = grads([mine.arena_D, mine.arena_Ds], orc.D, tag + ".D")
        rep[tag + ".D.grad.rel"], rep[tag + ".D.grad.cos"] = _rel(a, bb), _cos(a, bb)
        sdD, sdG = mine.netD.state_dict(), mine.netG.state_dict()
        for net, sd, P in (("D", sdD, orc.D), ("G", sdG, orc.G)):
            uv = [(sd[k].cpu().flatten(), P[k].detach().flatten()) for k in P if k.endswith("weight_u") or k.endswith("weight_v")]
            rep["%s.%s.uv.rel" % (tag, net)] = _rel(torch.cat([x for x, _ in uv]), torch.cat([y for _, y in uv]))
    losses = {n: (float(getattr(model, "loss_" + n).detach()), float(getattr(S, "loss_" + n)))
              for n in ("G_GAN", "D_real", "D_fake", "content", "gram")}
    vz, vreal = O.det_normal("sap.vz", (b, 128)), O.det_image("sap.vreal", b, 3, 64, 64)
    S.set_input(vz, vreal)
    S.clipping_mask_alpha()
    S.optimizer_netD_arch()
    model.set_input({"z": vz, "real_img": vreal, "img_path": ""})
    model.clipping_mask_alpha()
    model.optimizer_netD_arch()
    torch.cuda.synchronize()
    got, exp = model.get_current_losses(), S.get_current_losses()
    for n in ("D_arch_diff", "D_arch", "teacher_D_arch_diff"):
        losses[n] = (got[n], exp[n])
    a = torch.cat([model.arena_A.grads[n].float().cpu().flatten() for n in model.arena_A.grads])
    bb = torch.cat([v.grad.flatten() for n, v in S.D.items() if n.endswith("alpha")])
    rep["alpha.grad.rel"], rep["alpha.grad.cos"] = _rel(a, bb), _cos(a, bb)
    masks_ok = all(torch.equal(m.cpu(), om) for m, om in zip(model.netD.get_current_masks(), S.current_masks()))
    rep["losses"] = {k: {"b200": x, "oracle": y} for k, (x, y) in losses.items()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "step_parity_sagan.json"), "w"), indent=1)
    rep["worst"] = worst
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "step_parity_sagan.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
    bad = []
    for k, v in rep.items():
        if k in ("losses", "worst"):
            continue
        if k.endswith(".cos"):
            if v < (0.97 if ".G." in k else 0.99):
                bad.append((k, v))
        elif k.endswith("uv.rel"):
            if v > 2e-2:
                bad.append((k, v))
        elif k.endswith(".rel"):
            if v > (0.25 if ".G." in k else 0.15):
                bad.append((k, v))
        elif v > 5e-2:
            bad.append((k, v))
    for k, (x, y) in losses.items():
        if abs(x - y) > 5e-2 * abs(y) + 2e-3:
            bad.append(("loss." + k, x, y))
    assert masks_ok, "gate masks differ from the oracle"
    assert not bad, bad
