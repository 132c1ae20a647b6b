"""Which bf16 rounding point explains the gradient error of the B200 path against the fp32 oracle?

Runs the CPU oracle's pix2pix step (tiny widths) in fp32 and again with bf16 rounding EMULATED at chosen points
(weights / forward activations / backward activation-gradients), and prints the per-network gradient error of each
variant against the fp32 run (global rel-L2, cosine).  CPU only.

    python scripts/exp_bf16_emulation.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF

from oracle import gcc_oracle as O

FLAGS = {"w": False, "fwd": False, "bwd": False, "bwd_scope": "all"}


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


class Rnd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return bf(x) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (bf(g) if ctx.bwd else g), None, None


def rnd(x):
    return Rnd.apply(x, FLAGS["fwd"], FLAGS["bwd"])


class FProxy:
    """torch.nn.functional with rounding after every op the B200 path stores as bf16."""

    def __getattr__(self, name):
        return getattr(TF, name)

    def _w(self, w):
        return Rnd.apply(w, True, False) if (FLAGS["w"] and w is not None) else w

    def conv2d(self, x, w, b=None, **kw):
        return rnd(TF.conv2d(x, self._w(w), b, **kw))

    def conv_transpose2d(self, x, w, b=None, **kw):
        return rnd(TF.conv_transpose2d(x, self._w(w), b, **kw))

    def leaky_relu(self, x, s=0.01, *a):
        return rnd(TF.leaky_relu(x, s))

    def relu(self, x, *a):
        return rnd(TF.relu(x))


def run(tag, flags, tiny, batch):
    FLAGS.update(flags)
    O.F = FProxy()
    bn0, in0 = O.batch_norm, O.instance_norm
    if flags["fwd"] or flags["bwd"]:
        O_bn, O_in = bn0, in0
    S, T = O.build_pair(O.Opt(backbone="unet", direction="BtoA", **tiny))
    A, B = O.det_image("emu.A", batch, 3, 256, 256), O.det_image("emu.B", batch, 3, 256, 256)
    if flags["fwd"]:
        A, B = bf(A), bf(B)
    S.set_input(A, B)
    S.optimize_parameters()
    out = {}
    for name, net in (("S.G", S.G), ("S.D", S.D), ("T.G", T.G), ("T.D", T.D)):
        out[name] = torch.cat([v.grad.flatten() for k, v in net.items() if v.dtype == torch.float32 and v.grad is not None
                               and not k.endswith("alpha")])
    out["losses"] = {n: float(getattr(S, "loss_" + n)) for n in ("G_GAN", "G_L1", "D_real", "D_fake")}
    O.F = TF
    return out


def main():
    tiny = {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}
    batch = 2
    ref = run("fp32", {"w": False, "fwd": False, "bwd": False}, tiny, batch)
    variants = [("weights only", {"w": True, "fwd": False, "bwd": False}),
                ("forward activations only", {"w": False, "fwd": True, "bwd": False}),
                ("backward gradients only", {"w": False, "fwd": False, "bwd": True}),
                ("weights + forward", {"w": True, "fwd": True, "bwd": False}),
                ("all three (the B200 storage)", {"w": True, "fwd": True, "bwd": True})]
    for tag, fl in variants:
        got = run(tag, fl, tiny, batch)
        line = []
        for k in ("S.G", "S.D", "T.G", "T.D"):
            a, b = got[k].double(), ref[k].double()
            line.append("%s %.4f/%.5f" % (k, float((a - b).norm() / b.norm()), float(a @ b / (a.norm() * b.norm()))))
        print("%-32s" % tag, "  ".join(line), flush=True)


if __name__ == "__main__":
    main()
