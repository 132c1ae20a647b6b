"""Bisect which part of the iteration breaks CUDA-graph capture."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gcc_b200 import options, ops
from gcc_b200.pix2pix import Pix2PixModel, build_teacher

opt = options.parse(["--dataroot", "x/cityscapes", "--darts_discriminator", "--online_distillation", "--lambda_content", "50",
                     "--lambda_gram", "1e4", "--gpu_ids", "0", "--no_dropout"])
for k, v in {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}.items():
    setattr(opt, k, v)
model = Pix2PixModel(opt)
T = build_teacher(model, opt)
model.model_train()
B = 2
data = {k: torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1 for k in ("A", "B")}


def attempt(name, fn, warm=2):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warm):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print("CAPTURE OK  :", name, flush=True)
    except Exception as e:
        print("CAPTURE FAIL:", name, "->", str(e).split("\n")[0][:150], flush=True)
        torch.cuda.synchronize()


def f_forward():
    T.set_input({"A": data["A"], "B": data["B"]})
    T.forward()

def f_dfwd():
    f_forward()
    pred, _ = T._D(T.netD, T._fake_AB(T.fake_B_nhwc, True))
    T.loss_x = T._gan(pred, 1)

def f_loss_bwd_only():
    x = torch.zeros(B, 30, 30, 8, dtype=torch.bfloat16, device="cuda").requires_grad_(True)
    l = ops.GanLossFn.apply(x, 1, 0, 1)
    l.backward()

def f_torch_bwd_only():
    x = torch.zeros(B, 30, device="cuda").requires_grad_(True)
    (x * 2).sum().backward()

def f_act_bwd():
    x = torch.zeros(B, 30, 30, 8, dtype=torch.bfloat16, device="cuda").requires_grad_(True)
    y = ops.ActFn.apply(x, 1, 0.2)
    y.backward(torch.ones_like(y))

def f_dbwd():
    f_dfwd()
    T.set_requires_grad(T.netD, True)
    T.optimizer_D.zero_grad()
    T.loss_x.backward()

def f_tstep():
    T.set_input({"A": data["A"], "B": data["B"]})
    T.optimize_parameters()

def f_full():
    model.set_input({"A": data["A"], "B": data["B"]})
    model.optimize_parameters()
    model.set_input({"A": data["B"], "B": data["A"]})
    model.clipping_mask_alpha()
    model.optimizer_netD_arch()

attempt("torch-only backward", f_torch_bwd_only)
attempt("teacher G forward", f_forward)
attempt("teacher G + D forward + loss", f_dfwd)
attempt("GanLossFn fwd+bwd", f_loss_bwd_only)
attempt("ActFn fwd+bwd", f_act_bwd)
T.loss_x = None
attempt("teacher optimize_parameters", f_tstep)
attempt("full iteration", f_full)
l0 = model.get_current_losses()
print("losses after replay:", {k: round(v, 4) for k, v in l0.items()})
