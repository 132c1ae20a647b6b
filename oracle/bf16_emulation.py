"""TEST INFRASTRUCTURE (like everything under oracle/): the CPU oracle with the B200 path's bf16 STORAGE emulated.

The B200 kernels accumulate in fp32 but store weights (GEMM operand packs), activations and activation-gradients as
bf16.  `emulating(module)` swaps the oracle module's `torch.nn.functional` handle for a proxy that rounds to bf16 at
the same points: conv / transposed-conv weights, conv outputs, activation outputs (norm -> gate -> activation is ONE
fused kernel on the B200 path, so the norm output itself is not rounded) and, in backward, the gradients with respect
to those tensors.  The arithmetic between the rounding points stays the oracle's fp32 restatement of the reference.

Why it exists (scripts/exp_bf16_emulation.py, measured on the tiny pix2pix step): the discriminator-gradient error of
the B200 path against the fp32 oracle (4.6-8.6 % rel-L2) is NOT a backward-precision effect -- rounding only the
backward gradients gives 0.35 % -- it is the forward storage: rounding only the weights gives 3.4-4.7 %, only the
forward activations 4.6-5.9 %, all three 5.0-5.4 %.  The discriminator gradient is the sum of a real-batch and a
fake-batch term that nearly cancel at the same weights, so a 2^-9 relative perturbation of the forward pass moves
their difference by a few per cent.  Parity against THIS oracle is the tight check that no term is wrong
(tests/test_step_parity_gpu.py); parity against the fp32 oracle carries the stated bf16 bound.
"""
import contextlib

import torch
import torch.nn.functional as TF


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return bf(x) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (bf(g) if ctx.bwd else g), None, None


class FunctionalBf16:
    """`torch.nn.functional` with bf16 rounding where the B200 path stores bf16."""

    def __init__(self, weights=True, forward=True, backward=True):
        self.w, self.f, self.b = weights, forward, backward

    def __getattr__(self, name):
        return getattr(TF, name)

    def _act(self, x):
        return _Round.apply(x, self.f, self.b)

    def _w(self, w):
        return _Round.apply(w, True, False) if (self.w and w is not None) else w

    def conv2d(self, x, w, b=None, **kw):
        return self._act(TF.conv2d(x, self._w(w), b, **kw))

    def conv_transpose2d(self, x, w, b=None, **kw):
        return self._act(TF.conv_transpose2d(x, self._w(w), b, **kw))

    def leaky_relu(self, x, negative_slope=0.01, *a, **k):
        return self._act(TF.leaky_relu(x, negative_slope))

    def relu(self, x, *a, **k):
        return self._act(TF.relu(x))


@contextlib.contextmanager
def emulating(*modules, weights=True, forward=True, backward=True):
    """with emulating(gcc_oracle): ...   (every oracle module given has its `F` handle swapped for the block)"""
    saved = [(m, m.F) for m in modules]
    proxy = FunctionalBf16(weights, forward, backward)
    for m in modules:
        m.F = proxy
    try:
        yield proxy
    finally:
        for m, f in saved:
            m.F = f
