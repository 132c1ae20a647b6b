"""Bisect the SAGAN student-generator gradient error: GAN term only vs distillation terms only (norm ratio, rel, cos)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import options
from gcc_b200.sagan import SAGANModel, build_sa_teacher
from oracle import gcc_oracle as O, sagan_oracle as SA
TINY = {"ngf": 16, "teacher_ngf": 32, "ndf": 16, "teacher_ndf": 32}

def run(lc, lg, gan_scale_zero=False):
    opt = options.parse(["--dataroot", "x/celeb", "--model", "sagan", "--darts_discriminator", "--online_distillation",
                         "--lambda_content", str(lc), "--lambda_gram", str(lg), "--gpu_ids", "0"])
    for k, v in TINY.items(): setattr(opt, k, v)
    model = SAGANModel(opt); teacher = build_sa_teacher(model, opt)
    S, T = SA.build_sa_pair(SA.SAOpt(lambda_content=lc, lambda_gram=lg, **TINY))
    for mine, orc in ((model, S), (teacher, T)):
        mine.netG.load_state_dict({n: v.detach().clone() for n, v in orc.G.items()})
        mine.netD.load_state_dict({n: v.detach().clone() for n, v in orc.D.items()})
        with torch.no_grad():
            for i, w in enumerate(orc.transform): mine.transform_convs[i].weight.copy_(w.detach())
        mine.sync_weights(); mine.model_train()
    z, real = O.det_normal("sap.z", (4, 128)), O.det_image("sap.real", 4, 3, 64, 64)
    S.set_input(z, real); S.optimize_parameters()
    model.set_input({"z": z, "real_img": real, "img_path": ""}); model.optimize_parameters()
    torch.cuda.synchronize()
    print("lambda_content %g lambda_gram %g" % (lc, lg))
    for n in ("last.0.bias", "last.0.weight", "l4.0.module.weight_bar", "l4.1.weight", "attn2.value_conv.weight", "l2.1.weight", "l1.0.module.weight_bar"):
        for ar in (model.arena_G, model.arena_Gs):
            if n in ar.grads:
                a = ar.grads[n].float().cpu().flatten().double(); b = S.G[n].grad.flatten().double()
                print("   %-28s ratio %.4f rel %.4f cos %.4f  |ref| %.3e" % (n, float(a.norm() / b.norm()), float((a - b).norm() / b.norm()),
                      float(a @ b / (a.norm() * b.norm())), float(b.norm())))
run(0.0, 0.0)
run(1e-3, 0.0)
run(0.0, 1e-1)
