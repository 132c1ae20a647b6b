"""Bisect the CycleGAN generator-gradient mismatch: per-loss-term gradients, B200 vs oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import options, ops
from gcc_b200.cyclegan import MobileCycleGANModel
from oracle import gcc_oracle as O

TINY = {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}
opt = options.parse(["--dataroot", "x/horse2zebra", "--model", "cyclegan", "--gpu_ids", "0"])
for k, v in TINY.items():
    setattr(opt, k, v)
M = MobileCycleGANModel(opt)
S = O.CycleGANOracle(O.CycleOpt(direction=opt.direction, online_distillation=False, darts_discriminator=False, **TINY), "S")
for k in "AB":
    getattr(M, "netG_" + k).load_state_dict({n: v.detach() for n, v in S.G[k].items()})
    getattr(M, "netD_" + k).load_state_dict({n: v.detach() for n, v in S.D[k].items()})
M.sync_weights(); M.model_train()
A, B = O.det_image("cyc.A", 1, 3, 128, 128), O.det_image("cyc.B", 1, 3, 128, 128)

def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a @ b) / (a.norm() * b.norm() + 1e-30)), float(a.norm() / (b.norm() + 1e-30))

def compare(tag):
    for k in "AB":
        a, b = [], []
        for n, v in S.G[k].items():
            if v.dtype == torch.float32 and v.grad is not None and not n.endswith(".bias"):
                a.append(M.arena_G.grads[k + "." + n].detach().float().cpu().flatten()); b.append(v.grad.flatten())
        if b:
            print("%-22s G_%s rel %.4f cos %.4f norm-ratio %.3f" % ((tag, k) + rel(torch.cat(a), torch.cat(b))), flush=True)

def reset():
    M.arena_G.zero_grad()
    for k in "AB":
        for v in S.G[k].values():
            if v.dtype == torch.float32: v.grad = None

terms = ["idt_A", "idt_B", "cycle_A", "cycle_B", "G_A", "G_B", "all"]
for term in terms:
    reset()
    M.set_input({"A": A, "B": B}); M.forward(); M.set_requires_grad([M.netD_A, M.netD_B], False)
    S.set_input(A, B); S.forward(); S._set_grad(S.d_weights, False)
    l1 = lambda a, b: ops.DiffLossFn.apply(a, b, 3, 0)
    mine = {"idt_A": lambda: l1(M.idt_A_nhwc, M.real_B_nhwc), "idt_B": lambda: l1(M.idt_B_nhwc, M.real_A_nhwc),
            "cycle_A": lambda: l1(M.rec_A_nhwc, M.real_A_nhwc), "cycle_B": lambda: l1(M.rec_B_nhwc, M.real_B_nhwc),
            "G_A": lambda: M._gan(M._D(M.netD_A, M.fake_B_nhwc)[0], 0), "G_B": lambda: M._gan(M._D(M.netD_B, M.fake_A_nhwc)[0], 0)}
    orc = {"idt_A": lambda: (S.idt_A - S.real_B).abs().mean(), "idt_B": lambda: (S.idt_B - S.real_A).abs().mean(),
           "cycle_A": lambda: (S.rec_A - S.real_A).abs().mean(), "cycle_B": lambda: (S.rec_B - S.real_B).abs().mean(),
           "G_A": lambda: O.gan_loss("lsgan", S.netD("A", S.fake_B)[0], True), "G_B": lambda: O.gan_loss("lsgan", S.netD("B", S.fake_A)[0], True)}
    if term == "all":
        lm = sum(f() for f in mine.values()); lo = sum(f() for f in orc.values())
    else:
        lm, lo = mine[term](), orc[term]()
    lm.backward(); lo.backward()
    print("loss %s: b200 %.5f oracle %.5f" % (term, float(lm), float(lo)))
    compare(term)
