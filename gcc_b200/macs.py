"""Multiply-accumulate counter for the generators: the replacement of ``thop.profile`` in the pruning front-end
(utils/prune_util.py:6-18) and the source of the per-net GMACs behind bench.py's step-level FLOP accounting
(U-Net ngf 64 6.05, ngf 32 1.55, MobileResNet ngf 64 11.07, ngf 32 3.12, PatchGAN ndf 128 12.57 GMAC).

The count is analytic (a walk over the net's own layer objects with shape propagation; no forward pass, works on a
CPU-built net) with two conventions:

  "true"  every convolution's real multiply-accumulates: out_pixels * cout * cin/groups * k^2 for Conv2d,
          in_pixels * cin * cout * k^2 for ConvTranspose2d, plus the two attention matrix products of SAGAN.
  "thop"  the rule of thop's conv hook, which the reference's MAC budgets are quoted in (scripts/*/train.sh):
          out_elements * (cin/groups * k^2) for Conv2d AND ConvTranspose2d (a stride-2 transposed conv counts 4x its
          real work), only for modules whose forward hooks fire: SAGAN's spectral-normed convs are invoked through
          ``module.forward`` (SAGAN.py:69-71) and are invisible to it.  thop is an unvendored, unversioned dependency of
          the reference (requirements.txt:6) whose normalisation / activation rules changed between releases; only
          its convolution rule is reproduced, so thresholds searched against a budget are "parity unpinned"
           while the masks at a GIVEN threshold are bit-exact (gcc_b200/prune.py).
"""
from .ops import conv_out_hw


def _conv(layer, h, w, hooked=True):
    tr = layer.kind == "convT"
    oh, ow = conv_out_hw(h, w, layer.k, layer.stride, layer.pad, tr, getattr(layer, "outpad", 0))
    return {"name": layer.tname, "kind": layer.kind, "cin": layer.cin, "cout": layer.cout, "k": layer.k,
            "stride": layer.stride, "groups": 1, "in_hw": (h, w), "out_hw": (oh, ow), "hooked": hooked}, oh, ow


def _dw(layer, h, w):
    return {"name": layer.tname, "kind": "conv", "cin": layer.c, "cout": layer.c, "k": 3, "stride": 1, "groups": layer.c,
            "in_hw": (h, w), "out_hw": (h, w), "hooked": True}


def conv_list(net, input_hw=None):
    """Every convolution of a generator / discriminator with its input and output extent, in forward order."""
    from . import nets
    out = []
    cls = type(net).__name__
    if isinstance(net, nets.UnetGenertor):
        h, w = input_hw or (256, 256)
        sizes = {}
        reach = net.levels            # absent levels are skipped, the next present one is chained (nets.UnetGenertor)
        for i in reach:
            e, h, w = _conv(net.down[i], h, w)
            out.append(e)
            sizes[i] = (h, w)
        for i in reversed(reach):
            e, _, _ = _conv(net.up[i], *sizes[i])
            out.append(e)
        return out
    if isinstance(net, nets.MobileResnetGenerator):
        h, w = input_hw or (256, 256)
        for s in net.seq:
            kind = s[0]
            if kind in ("stem", "head"):
                e, _, _ = _conv(s[1], h + 6, w + 6)          # ReflectionPad2d(3) + k7 p0
                out.append(e)
            elif kind in ("down", "up"):
                e, h, w = _conv(s[1], h, w)
                out.append(e)
            elif kind == "block":
                for dw, _n1, pw, _n2 in s[1]:
                    out.append(_dw(dw, h, w))
                    e, _, _ = _conv(pw, h, w)
                    out.append(e)
        return out
    if isinstance(net, nets.NLayerDiscriminator):
        h, w = input_hw or (256, 256)
        for c in net.convs:
            e, h, w = _conv(c, h, w)
            out.append(e)
        return out
    if cls == "Generator" and hasattr(net, "blocks") and hasattr(net, "sub"):        # SRResNet (srgan.Generator)
        h, w = input_hw or (24, 24)
        out.append(_conv(net.c1, h, w)[0])
        for ca, _na, _pa, cb, _nb in net.blocks:
            out.append(_conv(ca, h, w)[0])
            out.append(_conv(cb, h, w)[0])
        out.append(_conv(net.c2, h, w)[0])
        for conv, _pr in net.sub:
            out.append(_conv(conv, h, w)[0])
            h, w = 2 * h, 2 * w
        out.append(_conv(net.c3, h, w)[0])
        return out
    if cls == "Generator" and hasattr(net, "at1"):                                    # sagan.Generator
        h = w = 1
        for i, sn in enumerate(net.sn):
            e, h, w = _conv(sn, h, w, hooked=False)
            out.append(e)
            if i == 2:
                out += _attn(net.at1, h, w)
        out += _attn(net.at2, h, w)
        out.append(_conv(net.last_conv, h, w)[0])
        return out
    raise NotImplementedError("no MAC walk for %s" % cls)


def _attn(at, h, w):
    res = [_conv(c, h, w)[0] for c in (at.q, at.k, at.v)]      # plain nn.Conv2d modules: their hooks fire
    L = h * w
    d, C = at.d, at.c
    res.append({"name": "attn.bmm", "kind": "bmm", "macs": L * L * d + L * L * C, "hooked": False})
    return res


def conv_macs(e, convention="true"):
    if e["kind"] == "bmm":
        return e["macs"] if convention == "true" else 0
    kk = e["k"] * e["k"] * (e["cin"] // e["groups"])
    oh, ow = e["out_hw"]
    if convention == "thop":
        return oh * ow * e["cout"] * kk if e["hooked"] else 0
    if e["kind"] == "convT":
        ih, iw = e["in_hw"]
        return ih * iw * e["cout"] * kk
    return oh * ow * e["cout"] * kk


def count_macs(net, input_hw=None, convention="true"):
    """Forward multiply-accumulates of one sample (an int)."""
    return sum(conv_macs(e, convention) for e in conv_list(net, input_hw))


def count_params(net):
    return sum(p.numel() for p in net.parameters())


def get_flops_parms(net, opt=None, input_hw=None, convention="thop"):
    """utils/prune_util.py:6-18: (GMACs, M parameters) of a generator on the input size the reference profiles:
    image_size / 4 for SRGAN, a latent vector for SAGAN, load_size otherwise."""
    if input_hw is None and opt is not None:
        root = str(getattr(opt, "dataroot", ""))
        if "sr" in root:
            input_hw = (opt.image_size // 4, opt.image_size // 4)
        elif "celeb" in root or "church" in root:
            input_hw = None
        else:
            input_hw = (opt.load_size, opt.load_size)
    return count_macs(net, input_hw, convention) / 1000.0 ** 3, count_params(net) / 1000.0 ** 2
