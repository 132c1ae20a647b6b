"""Flat parameter arenas: one contiguous fp32 buffer each for the parameters, gradients and Adam
moments of an optimizer group, plus the bf16 GEMM-operand packs of its convolution weights.

B200-first replacement for per-tensor ``torch.optim.Adam`` (reference: models/Pix2Pix.py:382-440):
one fused Adam launch per group, one memset for ``zero_grad``, one table-driven re-pack launch, and
the flat gradient buffer is exactly what the data-parallel all-reduce sends over NVLink.

Convolution weights keep the reference's logical shapes (OIHW / IOHW, so state-dicts round-trip)
but are stored channels-last ([D0][KH][KW][D1]), which is the layout the wgrad kernel writes and
the packer reads.
"""
import math

import torch
import torch.nn as nn

from . import _lib


def rp8(c):
    """Physical channel count of an NHWC bf16 activation with c logical channels: 8 for the <= 8-channel images (one
    16-byte vector per pixel: image-mode / row-window GEMMs), otherwise the next multiple of 16, so that every pixel
    row starts on a 32-byte sector.  (Round 1 padded to 8: the 24 / 40 / 88-channel tensors of the pruned students then
    have 48 / 80 / 176-byte rows, and every TMA box row and every vector store straddles sectors -- measured on SRGAN's
    24-channel 3x3 convs: 95 us against 50 us for the SAME tile count at 64 channels.)"""
    c = int(c)
    return 8 if c <= 8 else (c + 15) // 16 * 16


class ConvPacks:
    """bf16 operand packs of one conv weight: direct [D0][T][D1p] and transposed [D1][T][D0p]."""

    def __init__(self, d0, t, d1, device):
        self.d0, self.t, self.d1 = d0, t, d1
        self.d0p, self.d1p = rp8(d0), rp8(d1)
        self.direct = torch.zeros(d0, t, self.d1p, dtype=torch.bfloat16, device=device)
        self.transposed = torch.zeros(d1, t, self.d0p, dtype=torch.bfloat16, device=device)


class ParamArena:
    def __init__(self, device, lr=2e-4, betas=(0.9, 0.999), eps=1e-8):
        self.device = torch.device(device)
        self.specs = []
        self.lr, self.betas, self.eps = lr, betas, eps
        self.finalized = False
        self.dirty = True
        self.version = 0        # bumped by every re-pack: derived operand caches (ops.StemConvFn ...) key on it

    def add(self, name, shape, kind="vec"):
        """kind: 'conv' (O,I,KH,KW), 'convT' (I,O,KH,KW) or 'vec' (anything, stored contiguous); 'conv_nopack' /
        'convT_nopack' store like conv / convT but keep no bf16 operand packs (spectral-normed weights are packed
        per forward, scaled by 1 / sigma)."""
        assert not self.finalized
        self.specs.append((name, tuple(int(s) for s in shape), kind))
        return len(self.specs) - 1

    def finalize(self):
        off = 0
        self.offsets = []
        for _, shape, _ in self.specs:
            n = int(math.prod(shape)) if len(shape) else 1
            self.offsets.append((off, n))
            off += (n + 7) // 8 * 8
        self.numel = max(off, 8)
        dev = self.device
        self.P = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.G = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.M = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.V = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(5, dtype=torch.float32, device=dev)
        self._write_hyper()
        self.params, self.grads, self.packs, self.flat_grad = {}, {}, {}, {}
        table = []
        for (name, shape, kind), (o, n) in zip(self.specs, self.offsets):
            pv = self._view(self.P, o, n, shape, kind)
            gv = self._view(self.G, o, n, shape, kind)
            p = nn.Parameter(pv, requires_grad=True)
            p.grad = gv
            p._gcc_arena = self
            p._gcc_name = name
            self.params[name] = p
            self.grads[name] = gv
            self.flat_grad[name] = self.G[o:o + n]
            if kind in ("conv", "convT"):
                d0, d1, kh, kw = shape
                pk = ConvPacks(d0, kh * kw, d1, dev)
                self.packs[name] = pk
                if dev.type == "cuda":
                    table.append([self.P[o:o + n].data_ptr(), pk.direct.data_ptr(), pk.transposed.data_ptr(),
                                  d0, kh * kw, d1, pk.d1p, pk.d0p])
        self.table_count = len(table)
        self.table = torch.tensor(table if table else [[0] * 8], dtype=torch.int64, device=dev)
        self.finalized = True
        self.dirty = True
        return self

    @staticmethod
    def _view(flat, o, n, shape, kind):
        seg = flat[o:o + n]
        if kind.startswith("conv"):
            d0, d1, kh, kw = shape
            return seg.view(d0, kh, kw, d1).permute(0, 3, 1, 2)
        return seg.view(shape)

    def _write_hyper(self):
        h = torch.tensor([self.lr, self.betas[0], self.betas[1], self.eps], dtype=torch.float32)
        self.hyper[:4].copy_(h)

    def set_lr(self, lr):
        if lr != self.lr:
            self.lr = lr
            self.hyper[0:1].fill_(lr)

    # -- optimizer interface (mirrors the torch.optim.Optimizer calls the reference makes)
    def zero_grad(self):
        self.G.zero_()

    def step(self):
        _lib.call("gcc_adam_step_f32", self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(),
                  self.numel, self.hyper.data_ptr(), _lib.current_stream())
        self.dirty = True

    def mark_dirty(self):
        self.dirty = True

    def ensure_packed(self):
        if self.dirty:
            if self.table_count:
                _lib.call("gcc_pack_weights_table", self.table.data_ptr(), self.table_count, _lib.current_stream())
            self.dirty = False
            self.version += 1

    def state(self):
        return {"M": self.M, "V": self.V, "hyper": self.hyper}
