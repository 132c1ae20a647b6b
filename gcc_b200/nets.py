"""Generators and discriminators of the GCC pix2pix step on the B200 kernels.

Class names, constructor arguments and state-dict keys mirror the reference
(/root/reference/models/Pix2Pix.py:79-130 UnetGenertor, :199-265 MobileResnetGenerator,
:267-305 NLayerDiscriminator, :307-348 MaskNLayerDiscriminator,
/root/reference/models/DifferentiableOp.py:34-59 DifferentiableOP) so checkpoints round-trip and
``train.py`` / ``test.py`` style drivers work unchanged, but a forward is a flat sequence of fused
kernel launches over NHWC bf16 activations instead of an nn.Sequential of ATen ops.

Every net's ``forward`` takes and returns NHWC bf16 tensors (see ``ops.to_nhwc`` / ``ops.to_nchw``
for the NCHW fp32 boundary) and leaves the distillation taps the reference's forward hooks would
hold (models/Pix2Pix.py:702-731) in ``self.taps`` as (tensor, logical_channels) pairs.
"""
import torch
import torch.nn as nn

from . import ops
from .arena import ParamArena, rp8
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH


class _Tree(nn.Module):
    """Holder that registers parameters / buffers under dotted reference names."""

    def put(self, dotted, value, buffer=False):
        head, _, rest = dotted.partition(".")
        if rest:
            if head not in self._modules:
                self.add_module(head, _Tree())
            self._modules[head].put(rest, value, buffer)
        elif buffer:
            self.register_buffer(head, value)
        else:
            self.register_parameter(head, value)


class ConvLayer:
    """`name` is the reference state-dict name; `prefix` only disambiguates the arena key when several nets share
    one optimizer arena (CycleGAN's netG_A / netG_B)."""

    def __init__(self, arena, name, kind, cin, cout, k, stride, pad, outpad=0, bias=False, prefix=""):
        self.arena, self.kind = arena, kind
        self.cin, self.cout, self.k, self.stride, self.pad, self.outpad = cin, cout, k, stride, pad, outpad
        self.tname = name
        self.wname = prefix + name + ".weight"
        self.bname = prefix + name + ".bias" if bias else None
        shape = (cout, cin, k, k) if kind == "conv" else (cin, cout, k, k)
        arena.add(self.wname, shape, kind)
        if bias:
            arena.add(self.bname, (cout,), "vec")
        # k4 s2 p1 layers whose image side has <= 8 channels: image-mode GEMMs / col2im (ops.ColConvFn)
        self.colpath = (k == 4 and stride == 2 and pad == 1 and outpad == 0 and
                        ((kind == "conv" and cin <= 8) or (kind == "convT" and cout <= 8)))
        # k4 s1 p1 convs with <= 8 output channels (PatchGAN logits head) run as a 1x1 GEMM + fold (ops.HeadConvFn)
        self.headpath = (kind == "conv" and k == 4 and stride == 1 and pad == 1 and cout <= 8 and cin >= 64)
        # stride-1 convs with a large kernel and <= 8 channels on one side (k7 / k9 stem and head of the MobileResNet /
        # SRResNet generators, VGG's first conv): row-window GEMM (ops.StemConvFn) / 1x1 GEMM + fold (ops.FoldConvFn)
        self.stempath = (kind == "conv" and stride == 1 and 3 <= k <= 16 and cin <= 8 and cout >= 16 and outpad == 0)
        self.foldpath = (kind == "conv" and stride == 1 and 3 <= k <= 9 and cout <= 8 and cin >= 16 and not self.headpath)

    def bind(self):
        self.weight = self.arena.params[self.wname]
        self.bias = self.arena.params[self.bname] if self.bname else None
        self.packs = self.arena.packs[self.wname]

    def __call__(self, x, act=ACT_NONE, slope=0.2):
        fn = ops.ColConvFn if self.colpath else (ops.HeadConvFn if self.headpath else (
            ops.StemConvFn if self.stempath else (ops.FoldConvFn if self.foldpath else ops.ConvFn)))
        return fn.apply(x, self.weight, self.bias, self, act, slope)

    def with_stats(self, x):
        """Conv whose epilogue also accumulates the per-channel sum / sum of squares a following BatchNorm needs.
        Returns (y, sums) with sums = None when the layer shape does not take the fused path."""
        n, h, w, _ = x.shape
        oh, ow = ops.conv_out_hw(h, w, self.k, self.stride, self.pad, self.kind == "convT", self.outpad)
        if self.colpath or self.headpath or self.stempath or self.foldpath or n * oh * ow <= 16384 or not x.is_cuda:
            return self(x), None
        sums = ops.zero_pool.take(2 * rp8(self.cout), x.device)
        return ops.ConvFn.apply(x, self.weight, self.bias, self, ACT_NONE, 0.2, sums), sums


class DwConvLayer:
    def __init__(self, arena, name, c, prefix=""):
        self.arena, self.c = arena, c
        self.tname = name
        self.wname, self.bname = prefix + name + ".weight", prefix + name + ".bias"
        arena.add(self.wname, (c, 1, 3, 3), "vec")
        arena.add(self.bname, (c,), "vec")

    def bind(self):
        self.weight = self.arena.params[self.wname]
        self.bias = self.arena.params[self.bname]

    def __call__(self, x):
        return ops.DwConvFn.apply(x, self.weight, self.bias, self)


class NormLayer:
    """mode 'bn' (affine, running stats), 'in' (InstanceNorm2d affine=False) or 'id'; optional gate."""

    def __init__(self, arena, name, c, mode, device, gate_arena=None, gate_name=None, thr=0.5, slope=0.2, prefix=""):
        self.c, self.mode, self.thr, self.slope = c, mode, float(thr), slope
        self.tname, self.tgate = name, gate_name
        name = prefix + name
        gate_name = prefix + gate_name if gate_name is not None else None
        self.name, self.arena = name, arena
        self.training = True
        self.stats_hook = None
        self.num_batches = 0
        self.running_mean = self.running_var = None
        if mode == "bn":
            arena.add(name + ".weight", (c,), "vec")
            arena.add(name + ".bias", (c,), "vec")
            self.running_mean = torch.zeros(c, dtype=torch.float32, device=device)
            self.running_var = torch.ones(c, dtype=torch.float32, device=device)
        self.gate_arena, self.gate_name = gate_arena, gate_name
        if gate_name is not None:
            gate_arena.add(gate_name + ".alpha", (c,), "vec")

    def bind(self):
        self.gamma = self.arena.params[self.name + ".weight"] if self.mode == "bn" else None
        self.beta = self.arena.params[self.name + ".bias"] if self.mode == "bn" else None
        self.alpha = self.gate_arena.params[self.gate_name + ".alpha"] if self.gate_name else None

    def __call__(self, x, act=ACT_NONE, act2=None, sums=None, y_into=None, y2_into=None):
        return ops.NormActFn.apply(x, self.gamma, self.beta, self.alpha, self, act, act2, sums, y_into, y2_into)


class _Net(nn.Module):
    """Common plumbing: arenas -> registered parameters/buffers with reference names."""

    def _register(self, layers):
        tree = _Tree()
        self._norms = []
        for l in layers:
            l.bind()
            if isinstance(l, ConvLayer):
                tree.put(l.tname + ".weight", l.weight)
                if l.bias is not None:
                    tree.put(l.tname + ".bias", l.bias)
            elif isinstance(l, DwConvLayer):
                tree.put(l.tname + ".weight", l.weight)
                tree.put(l.tname + ".bias", l.bias)
            elif isinstance(l, NormLayer):
                self._norms.append(l)
                if l.mode == "bn":
                    tree.put(l.tname + ".weight", l.gamma)
                    tree.put(l.tname + ".bias", l.beta)
                    tree.put(l.tname + ".running_mean", l.running_mean, buffer=True)
                    tree.put(l.tname + ".running_var", l.running_var, buffer=True)
                    l.nbt = torch.zeros((), dtype=torch.long, device=l.running_mean.device)
                    tree.put(l.tname + ".num_batches_tracked", l.nbt, buffer=True)
                if l.alpha is not None:
                    tree.put(l.tgate + ".alpha", l.alpha)
        # expose the reference's top-level attribute name ("model")
        for k, m in tree._modules.items():
            self.add_module(k, m)

    def train(self, mode=True):
        super().train(mode)
        for l in getattr(self, "_norms", []):
            l.training = mode
        return self

    def sync_counters(self):
        for l in self._norms:
            if l.mode == "bn":
                l.nbt.fill_(l.num_batches)

    def state_dict(self, *a, **k):
        self.sync_counters()
        sd = super().state_dict(*a, **k)
        return type(sd)((key, v.contiguous() if torch.is_tensor(v) else v) for key, v in sd.items())

    def load_state_dict(self, sd, strict=True):
        own = dict(self.named_parameters())
        own.update(dict(self.named_buffers()))
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError("state_dict mismatch: missing %s unexpected %s" % (missing, unexpected))
        with torch.no_grad():
            for k, v in sd.items():
                if k in own:
                    own[k].copy_(v.to(own[k].device))
        for l in self._norms:
            if l.mode == "bn":
                l.num_batches = int(l.nbt)
        for a in self._arenas:
            a.mark_dirty()


# --------------------------------------------------------------------------------------- U-Net
def unet_level_prefixes(num_downs=8):
    pre = ["model.model"]
    for i in range(1, num_downs):
        pre.append(pre[-1] + (".1" if i == 1 else ".3") + ".model")
    return pre


def unet_channels(ngf, f=None, c=None, input_nc=3, output_nc=3):
    """(down_in, down_out, up_in, up_out) per level, 0 = outermost (models/Pix2Pix.py:87-127)."""
    lv = [None] * 8
    lv[0] = (input_nc, ngf if f is None else f[0], ngf * 2 if c is None else c[14], output_nc)
    lv[1] = (ngf if c is None else c[0], ngf * 2 if f is None else f[1], ngf * 4 if c is None else c[13],
             ngf if f is None else f[14])
    lv[2] = (ngf * 2 if c is None else c[1], ngf * 4 if f is None else f[2], ngf * 8 if c is None else c[12],
             ngf * 2 if f is None else f[13])
    lv[3] = (ngf * 4 if c is None else c[2], ngf * 8 if f is None else f[3], ngf * 16 if c is None else c[11],
             ngf * 4 if f is None else f[12])
    for i in range(3):
        lv[6 - i] = (ngf * 8 if c is None else c[5 - i], ngf * 8 if f is None else f[6 - i],
                     ngf * 16 if c is None else c[8 + i], ngf * 8 if f is None else f[9 + i])
    lv[7] = (ngf * 8 if c is None else c[6], ngf * 8 if f is None else f[7], ngf * 8 if c is None else c[7],
             ngf * 8 if f is None else f[8])
    present = [True] * 8
    if f is not None:
        if f[7] == 0 or f[8] == 0:
            present[7] = False
        for i in range(3):
            if f[6 - i] == 0 or f[9 + i] == 0:
                present[6 - i] = False
    return lv, present


class UnetGenertor(_Net):
    """U-Net-256 generator (spelling of the class name follows the reference)."""

    def __init__(self, input_nc=3, output_nc=3, num_downs=8, ngf=64, use_dropout=False, filter_cfgs=None,
                 channel_cfgs=None, arena=None, device="cuda"):
        super().__init__()
        assert num_downs == 8, "the reference only builds the 8-level U-Net"
        self.arena = arena if arena is not None else ParamArena(device, betas=(0.5, 0.999))
        self._arenas = [self.arena]
        self.ngf, self.use_dropout = ngf, use_dropout
        self.lv, self.present = unet_channels(ngf, filter_cfgs, channel_cfgs, input_nc, output_nc)
        # Levels whose cfg entry is zero are left out and the next deeper PRESENT level becomes the submodule
        # (models/Pix2Pix.py:87-102); module names follow the nesting actually built.  The deepest present level gets
        # an Identity submodule when it is not the innermost block (Pix2Pix.py:62-69).
        self.levels = [i for i in range(8) if self.present[i]]
        self.next_level = {a: b for a, b in zip(self.levels, self.levels[1:])}
        pre_list = unet_level_prefixes(len(self.levels))
        pre = {lvl: pre_list[k] for k, lvl in enumerate(self.levels)}
        for a, b in self.next_level.items():
            if self.lv[b][0] != self.lv[a][1] or self.lv[a][2] != self.lv[a][1] + self.lv[b][3]:
                raise ValueError("U-Net cfg: level %d (out %d, up-in %d) does not chain to level %d (in %d, up-out %d)" % (
                    a, self.lv[a][1], self.lv[a][2], b, self.lv[b][0], self.lv[b][3]))
        A = self.arena
        self.down, self.dnorm, self.up, self.unorm = {}, {}, {}, {}
        layers = []
        for i in self.levels:
            di, do, ui, uo = self.lv[i]
            p = pre[i]
            if i == 0:
                self.down[0] = ConvLayer(A, p + ".0", "conv", di, do, 4, 2, 1)
                self.dnorm[0] = NormLayer(A, p + ".act0", do, "id", device)
                self.up[0] = ConvLayer(A, p + ".3", "convT", ui, uo, 4, 2, 1, bias=True)
            elif i == 7:
                self.down[7] = ConvLayer(A, p + ".1", "conv", di, do, 4, 2, 1)
                self.up[7] = ConvLayer(A, p + ".3", "convT", ui, uo, 4, 2, 1)
                self.unorm[7] = NormLayer(A, p + ".4", uo, "bn", device)
            else:
                self.down[i] = ConvLayer(A, p + ".1", "conv", di, do, 4, 2, 1)
                self.dnorm[i] = NormLayer(A, p + ".2", do, "bn", device)
                self.up[i] = ConvLayer(A, p + ".5", "convT", ui, uo, 4, 2, 1)
                self.unorm[i] = NormLayer(A, p + ".6", uo, "bn", device)
        # registration order = reference module order (down modules, submodule, up modules)
        def order(i):
            layers.append(self.down[i])
            if i in self.dnorm:
                layers.append(self.dnorm[i])
            if i in self.next_level:
                order(self.next_level[i])
            layers.append(self.up[i])
            if i in self.unorm:
                layers.append(self.unorm[i])
        order(0)
        self._layers = layers
        if arena is None:
            self.finalize()
        self.seed = None
        self.taps = []

    def finalize(self):
        if not self.arena.finalized:
            self.arena.finalize()
        self._register(self._layers)
        self.seed = torch.zeros(1, dtype=torch.int64, device=self.arena.device)

    def _skip_buffer(self, i, like):
        """Buffer of the concatenation cat[skip of level i, up output of the next level] (Pix2Pix.py:77) when both halves
        are whole 8-channel groups and the child's up path has no dropout in between: the producers then write their
        halves in place (zero-copy concat).  None = assemble with a copy (ops.CatFn)."""
        j = self.next_level.get(i)
        if j is None or not like.is_cuda:
            return None
        ca, uo = self.lv[i][1], self.lv[j][3]
        if rp8(ca) != ca or rp8(uo) != uo or (self.use_dropout and self.training and j in (4, 5, 6)):
            return None
        n, h, w, _ = like.shape
        return torch.empty(n, h, w, ca + uo, dtype=torch.bfloat16, device=like.device)

    def _block(self, i, a, a_relu, ca, catbuf=None):
        """a = lrelu(parent activation) with ca logical channels, a_relu = relu of the same.
        Returns relu(cat[a, up_i]) and its logical channel count."""
        di, do, ui, uo = self.lv[i]
        if i == 7:
            d = self.down[i](a)
            rc, crc = ops.ActFn.apply(d, ACT_RELU, 0.0), do
        else:
            d, dsums = self.down[i].with_stats(a)
            buf = self._skip_buffer(i, d)
            y, y2 = self.dnorm[i](d, ACT_LRELU, ACT_RELU, dsums, None, None if buf is None else (buf, 0))
            if i in self.next_level:
                rc, crc = self._block(self.next_level[i], y, y2, do, buf)
                feat = y
            else:                      # Identity submodule: the in-place uprelu also mutates the hooked tensor
                rc, crc, feat = y2, do, y2
            if i == 1:
                self.taps[0], self.taps[3] = (feat, do), (rc, crc)
            if i == 3:
                self.taps[1], self.taps[2] = (feat, do), (rc, crc)
        u, usums = self.up[i].with_stats(rc)
        if catbuf is not None:
            ub = self.unorm[i](u, ACT_RELU, None, usums, (catbuf, ca), None)
            return ops.CatViewFn.apply(a_relu, ub, catbuf), ca + uo
        ub = self.unorm[i](u, ACT_RELU, None, usums)
        if self.use_dropout and i in (4, 5, 6) and self.training:
            self.seed_salt += 1
            ub = ops.DropoutFn.apply(ub, 0.5, self.seed, self.seed_salt)
        return ops.CatFn.apply(a_relu, ub, ca, uo), ca + uo

    def forward(self, x):
        self.taps = [None] * 4
        self.seed_salt = 0
        if self.use_dropout and self.training:
            self.seed.add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)
        di, do, ui, uo = self.lv[0]
        d0 = self.down[0](x)
        buf = self._skip_buffer(0, d0)
        a, a_relu = self.dnorm[0](d0, ACT_LRELU, ACT_RELU, None, None, None if buf is None else (buf, 0))
        r, cr = self._block(self.next_level[0], a, a_relu, do, buf)
        return self.up[0](r, ACT_TANH)


# -------------------------------------------------------------------------------- MobileResNet
class MobileResnetGenerator(_Net):
    def __init__(self, input_nc=3, output_nc=3, ngf=64, n_blocks=9, cfg=None, arena=None, device="cuda", opt=None,
                 dropout_rate=0, padding_type="reflect", prefix=""):
        super().__init__()
        if padding_type != "reflect":
            raise NotImplementedError("padding [%s] is not implemented" % padding_type)
        self.arena = arena if arena is not None else ParamArena(device, betas=(0.5, 0.999))
        self._arenas = [self.arena]
        A = self.arena
        self.ngf, self.cfg = ngf, cfg
        layers = []
        seq = []  # (kind, ...)
        idx = 0
        c0 = ngf if cfg is None else cfg[0]
        conv = ConvLayer(A, "model.1", "conv", input_nc, c0, 7, 1, 0, bias=True, prefix=prefix)
        seq.append(("stem", conv, NormLayer(A, "model.2", c0, "in", device, prefix=prefix)))
        idx = 1
        for i in range(2):
            mult = 2 ** i
            cin = ngf * mult if cfg is None else cfg[idx - 1]
            cout = ngf * mult * 2 if cfg is None else cfg[idx]
            idx += 1
            name = "model.%d" % (4 + 3 * i)
            seq.append(("down", ConvLayer(A, name, "conv", cin, cout, 3, 2, 1, bias=True, prefix=prefix),
                        NormLayer(A, "model.%d" % (5 + 3 * i), cout, "in", device, prefix=prefix), name))
        mi = 10
        for i in range(n_blocks):
            c_in = ngf * 4 if cfg is None else cfg[idx - 1]
            c_mid = ngf * 4 if cfg is None else cfg[idx]
            idx += 1
            c_out = ngf * 4 if cfg is None else cfg[idx]
            idx += 1
            if c_mid == 0:
                continue
            name = "model.%d" % mi
            parts = []
            for j, (a, b) in ((1, (c_in, c_mid)), (6, (c_mid, c_out))):
                q = "%s.conv_block.%d.conv" % (name, j)
                parts.append((DwConvLayer(A, q + ".0", a, prefix=prefix), NormLayer(A, q + ".1", a, "in", device, prefix=prefix),
                              ConvLayer(A, q + ".2", "conv", a, b, 1, 1, 0, bias=True, prefix=prefix),
                              NormLayer(A, "%s.conv_block.%d" % (name, j + 1), b, "in", device, prefix=prefix)))
            seq.append(("block", parts, name, c_out))
            mi += 1
        out_ch = ngf
        for i in range(2):
            mult = 2 ** (2 - i)
            cin = ngf * mult if cfg is None else cfg[idx - 1]
            out_ch = int(ngf * mult / 2) if cfg is None else cfg[idx]
            idx += 1
            seq.append(("up", ConvLayer(A, "model.%d" % mi, "convT", cin, out_ch, 3, 2, 1, outpad=1, bias=True, prefix=prefix),
                        NormLayer(A, "model.%d" % (mi + 1), out_ch, "in", device, prefix=prefix)))
            mi += 3
        seq.append(("head", ConvLayer(A, "model.%d" % (mi + 1), "conv", out_ch, output_nc, 7, 1, 0, bias=True, prefix=prefix)))
        self.seq = seq
        for s in seq:
            if s[0] == "block":
                for dw, n1, pw, n2 in s[1]:
                    layers += [dw, n1, pw, n2]
            else:
                layers += [l for l in s[1:] if isinstance(l, (ConvLayer, NormLayer))]
        self._layers = layers
        if arena is None:
            self.finalize()
        self.taps = []

    def finalize(self):
        if not self.arena.finalized:
            self.arena.finalize()
        self._register(self._layers)

    def forward(self, x):
        self.taps = []
        h = x
        for s in self.seq:
            kind = s[0]
            if kind == "stem":
                h = s[2](s[1](ops.ReflectPadFn.apply(h, 3)), ACT_RELU)
            elif kind == "down":
                h = s[2](s[1](h), ACT_RELU)
                if s[3] == "model.7":
                    self.taps.append((h, s[1].cout))
            elif kind == "block":
                r = h
                if ops.slab_ok(h):
                    # 64 x 64 (and smaller) blocks: dw3x3 -> IN and IN -> ReLU / IN + skip as single launches (slab.cu)
                    for j, (dw, n1, pw, n2) in enumerate(s[1]):
                        t = pw(ops.DwInSlabFn.apply(r, dw.weight, dw.bias, dw))
                        r = ops.InActSlabFn.apply(t, None if j == 0 else h, pw.cout, 2 if j == 0 else 0, 0.0)
                    h = r
                else:
                    for j, (dw, n1, pw, n2) in enumerate(s[1]):
                        r = n2(pw(n1(dw(r), ACT_NONE)), ACT_RELU if j == 0 else ACT_NONE)
                    h = ops.AddFn.apply(h, r)
                if s[2] in ("model.12", "model.15", "model.18"):
                    self.taps.append((h, s[3]))
            elif kind == "up":
                h = s[2](s[1](h), ACT_RELU)
            else:
                h = s[1](ops.ReflectPadFn.apply(h, 3), ACT_TANH)
        return h


# ------------------------------------------------------------------------------------ PatchGAN
class DifferentiableOP(nn.Module):
    """Stand-alone channel gate with the reference's interface (models/DifferentiableOp.py:34-59)."""

    def __init__(self, output_channel, threshold=0.5, device="cuda"):
        super().__init__()
        self.arena = ParamArena(device)
        self.norm = NormLayer(self.arena, "gate", output_channel, "id", device, self.arena, "gate", threshold)
        self.arena.finalize()
        self.norm.bind()
        self.arena.P.fill_(1.0)
        self.alpha = self.norm.alpha
        self.threshold = threshold * torch.ones(1)

    def forward(self, x_nhwc):
        return self.norm(x_nhwc, ACT_NONE)

    def clip_alpha(self):
        ops.call("gcc_clamp_f32", self.arena.P.data_ptr(), self.arena.numel, 0.0, 1.0, ops._st())

    def get_current_mask(self):
        return (torch.sign(self.alpha.detach() - float(self.threshold)) + 1) / 2


class NLayerDiscriminator(_Net):
    """PatchGAN-70 (models/Pix2Pix.py:267-305); ``gated=True`` gives MaskNLayerDiscriminator (:307-348)."""
    gated = False

    def __init__(self, input_nc=3, ndf=64, n_layers=3, threshold=0.5, arena=None, gate_arena=None, device="cuda",
                 norm="bn", prefix=""):
        """norm='in': CycleGAN's plain discriminator (InstanceNorm2d, every conv has a bias; CycleGAN.py:140-178)."""
        super().__init__()
        assert n_layers == 3
        gated = self.gated
        self.arena = arena if arena is not None else ParamArena(device, betas=(0.5, 0.999))
        self.gate_arena = (gate_arena if gate_arena is not None else ParamArena(device)) if gated else None
        self._arenas = [self.arena] + ([self.gate_arena] if gated else [])
        A, GA = self.arena, self.gate_arena
        idx = ([(0, None, 2), (3, 4, 5), (7, 8, 9), (11, 12, 13), (15, None, None)] if gated else
               [(0, None, None), (2, 3, None), (5, 6, None), (8, 9, None), (11, None, None)])
        ch = [input_nc, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
        self.ch, self.threshold = ch, threshold
        self.convs, self.norms = [], []
        layers = []
        for li, (ci, bi, gi) in enumerate(idx):
            conv = ConvLayer(A, "model.%d" % ci, "conv", ch[li], ch[li + 1], 4, 2 if li < 3 else 1, 1,
                             bias=(li in (0, 4)) or norm == "in", prefix=prefix)
            self.convs.append(conv)
            layers.append(conv)
            if li == 4:
                break
            gname = "model.%d" % gi if gi is not None else None
            if li == 0:
                nl = NormLayer(A, "model.act0", ch[1], "id", device, GA, gname, threshold, prefix=prefix)
            else:
                nl = NormLayer(A, "model.%d" % bi, ch[li + 1], norm, device, GA, gname, threshold, prefix=prefix)
            self.norms.append(nl)
            layers.append(nl)
        self._layers = layers
        if arena is None:
            self.finalize()
        self.taps = []

    def finalize(self):
        for a in self._arenas:
            if not a.finalized:
                a.finalize()
        self._register(self._layers)
        if self.gated:
            with torch.no_grad():
                self.gate_arena.P.fill_(1.0)  # alpha init = 1 (DifferentiableOp.py:39)

    def forward(self, x):
        """x: [N, H, W, 8] (6 logical channels).  Returns the patch logits [N, H/8-2, W/8-2, 8] (1 logical)."""
        self.taps = []
        h = x
        for li in range(4):
            if li == 0 and self.norms[0].alpha is None:
                h = self.convs[0](h, ACT_LRELU, 0.2)        # plain D: LeakyReLU fused into the conv epilogue
            elif li == 0:
                h = self.norms[0](self.convs[0](h), ACT_LRELU)
            elif self.norms[li].mode == "in" and self.norms[li].alpha is None:
                c = self.convs[li](h)
                if ops.slab_ok(c):
                    h = ops.InActSlabFn.apply(c, None, self.ch[li + 1], 1, 0.2)      # InstanceNorm + LeakyReLU, one launch
                else:
                    h = self.norms[li](c, ACT_LRELU)
            else:
                c, csums = self.convs[li].with_stats(h)
                h = self.norms[li](c, ACT_LRELU, None, csums)
            if li in (1, 3) and not self.gated:
                self.taps.append((h, self.ch[li + 1]))
        return self.convs[4](h)

    def alphas(self):
        return [n.alpha for n in self.norms if n.alpha is not None]

    def get_current_masks(self):
        return [(torch.sign(a.detach() - self.threshold) + 1) / 2 for a in self.alphas()]


class MaskNLayerDiscriminator(NLayerDiscriminator):
    gated = True
