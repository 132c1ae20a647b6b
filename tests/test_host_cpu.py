"""CPU-side checks: the C-ABI library loads and exports every symbol include/gcc_b200.h declares, the host
modules mirror the reference's state-dict layout, options and prune index selection match the oracle."""
import ctypes
import os

import pytest
import torch

from oracle import gcc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gcc_b200 import _build, _lib
    path = _build.build()
    lib = ctypes.CDLL(path)
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(lib, name), "libgcc_b200.so lacks %s declared in include/gcc_b200.h" % name
    lib.gcc_abi_version.restype = ctypes.c_int
    assert lib.gcc_abi_version() == 1


def test_launch_planning_is_pure_host_arithmetic():
    """The two launch-planning decisions of the GEMM kernels (tail-wave split of the persistent conv kernel, split-K
    factor of the weight-gradient kernel) are exported as pure host functions: pinned here on the shapes of the c2
    iteration and against an independent restatement of the cost model."""
    from gcc_b200 import _build, _lib
    lib = ctypes.CDLL(_build.build())
    for name in ("gcc_plan_conv_tail", "gcc_plan_wgrad_splits"):
        getattr(lib, name).restype = ctypes.c_int
        getattr(lib, name).argtypes = [ctypes.c_int] * len(_lib.parse_header()[name][1])
    tail = lib.gcc_plan_conv_tail
    # 512 tiles on 148 SMs = 3.46 waves, 68 tiles in the last one
    assert tail(512, 256, 148, 128) == 2          # PatchGAN 1024 -> 512 data gradient (256 k-blocks per tile)
    assert tail(512, 64, 148, 128) == 0           # 256 -> 512 k4 s2 forward: below the default threshold (measured slower)
    assert tail(512, 64, 148, 32) == 2            # ... taken when the threshold is lowered (the unit test's setting)
    assert tail(512, 24, 148, 16) == 0            # never below 32 k-blocks: a part keeps >= 16
    assert tail(1024, 256, 148, 128) == 0         # 6.92 waves: the last wave is 92 % full
    assert tail(454, 256, 148, 128) == 4          # 10 tiles left: 14 parts would fit, at most 4 are taken
    assert tail(100, 256, 148, 128) == 0          # less than one wave: the classic split-K path decides
    assert tail(148 * 9 + 10, 256, 148, 128) == 0 # many waves: the tail does not matter
    assert tail(444, 256, 148, 128) == 0          # whole waves

    def model(base_ctas, total_pb, bn, mt, c8):
        t_kb, t_epi = 2.0 * bn * mt * 1.4, 40.0 * bn * mt + 6000.0
        slots = 148 * (2 if (bn <= 128 and mt == 1) else 1)
        cands = []
        for sp in range(1, (256 if c8 else 64) + 1):
            kb = (total_pb + sp - 1) // sp
            if sp > 1 and kb < 8:
                break
            waves = (base_ctas * sp + slots - 1) // slots
            cands.append((sp, waves * (kb * t_kb + t_epi)))
        best = min(c for _, c in cands)
        return next(sp for sp, c in cands if c <= best * 1.03)

    splits = lib.gcc_plan_wgrad_splits
    assert splits(128, 512, 256, 2, 0) == 1       # PatchGAN 512 -> 1024: 128 CTAs, one long K loop each
    assert splits(16, 2048, 128, 2, 0) == 9       # 128 -> 256 k4 s2: 16 x 9 = 144 CTAs fill the 148 SMs (8 left 20 idle)
    for base in (1, 4, 16, 32, 128, 512):
        for pb in (4, 16, 128, 512, 2048, 8192):
            for bn, mt in ((64, 1), (128, 1), (128, 2), (256, 1), (256, 2)):
                assert splits(base, pb, bn, mt, 0) == model(base, pb, bn, mt, 0), (base, pb, bn, mt)
    assert splits(1, 8192, 128, 1, 1) == model(1, 8192, 128, 1, 1)


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gcc_b200 import _lib, ops
    with pytest.raises(_lib.GccB200Error):
        ops.ActFn.apply(torch.zeros(1, 2, 2, 8, dtype=torch.bfloat16), 1, 0.2)
    from gcc_b200 import options
    from gcc_b200.pix2pix import Pix2PixModel
    opt = options.parse(["--dataroot", "x", "--gpu_ids", "-1"])
    with pytest.raises(RuntimeError):
        Pix2PixModel(opt)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gcc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def _check(net, shapes):
    sd = net.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k


FC = [8, 13, 30, 61, 64, 59, 40, 64, 37, 50, 64, 48, 27, 14, 5]
CC = [8, 13, 30, 61, 64, 59, 40, 64, 77, 109, 128, 109, 57, 27, 13]
RC = [8, 16, 29, 21, 29, 17, 29, 30, 29, 11, 29, 25, 29, 32, 29, 9, 29, 27, 29, 19, 29, 13, 7]


def test_state_dict_layout_matches_reference_names():
    from gcc_b200 import nets
    _check(nets.UnetGenertor(ngf=8, device="cpu"), O.unet_param_shapes(8))
    _check(nets.UnetGenertor(ngf=8, filter_cfgs=FC, channel_cfgs=CC, device="cpu"), O.unet_param_shapes(8, FC, CC))
    _check(nets.MobileResnetGenerator(ngf=8, device="cpu"), O.resnet_param_shapes(8))
    _check(nets.MobileResnetGenerator(ngf=8, cfg=RC, device="cpu"), O.resnet_param_shapes(8, RC))
    _check(nets.NLayerDiscriminator(6, 16, device="cpu"), O.patchgan_param_shapes(16, 6, False))
    _check(nets.MaskNLayerDiscriminator(6, 16, device="cpu"), O.patchgan_param_shapes(16, 6, True))


def test_state_dict_roundtrip_and_channels_last_storage():
    from gcc_b200 import nets
    net = nets.UnetGenertor(ngf=8, device="cpu")
    P = O._make_params(O.unet_param_shapes(8), "S.netG.")
    net.load_state_dict({k: v.detach() for k, v in P.items()})
    sd = net.state_dict()
    assert all(torch.equal(sd[k], P[k].detach()) for k in P)
    w = net.arena.params["model.model.0.weight"]
    assert w.shape == (8, 3, 4, 4) and w.stride() == (48, 1, 12, 3)  # [O][KH][KW][I] storage
    assert w.grad.stride() == w.stride()
    with pytest.raises(RuntimeError):
        net.load_state_dict({"bogus": torch.zeros(1)})


def test_prune_cfgs_bit_exact_vs_golden(golden_dir):
    from gcc_b200 import prune
    g = torch.load(os.path.join(golden_dir, "pix2pix_small_ops.pt"), weights_only=False)["prune"]
    G = {k: v.detach() for k, v in O._make_params(O.unet_param_shapes(16), "P.netG.").items()}
    for thr in (0.98, 1.0, 1.02):
        assert prune.unet_scale_prune_cfg(G, 16, thr) == tuple(g["scale_prune@%g" % thr])
    for thr in (2.0, 6.0, 10.0):
        assert prune.unet_norm_prune_cfg(G, 16, thr) == tuple(g["norm_prune@%g" % thr])
    mx, mn = prune.unet_max_min_bn_scale(G)
    assert (mx, mn) == pytest.approx(g["scale_prune.maxmin"], rel=1e-6)
    mx, mn = prune.max_min_conv_norm(G, "unet")
    assert (mx, mn) == pytest.approx(g["norm_prune.maxmin"], rel=1e-6)
    R = {k: v.detach() for k, v in O._make_params(O.resnet_param_shapes(16), "P.netG.").items()}
    for thr in (0.5, 2.3, 2.6):
        assert prune.resnet_prune_cfg(R, thr) == g["resnet_prune@%g" % thr]
    mx, mn = prune.max_min_conv_norm(R, "resnet")
    assert (mx, mn) == pytest.approx(g["resnet.maxmin"], rel=1e-6)


def test_options_overrides():
    from gcc_b200 import options
    o = options.parse(["--dataroot", "./database/cityscapes/", "--ngf", "32", "--darts_discriminator"])
    assert (o.lambda_L1, o.teacher_ndf, o.direction, o.load_size, o.gpu_ids) == (100.0, 128, "BtoA", 256, [0])
    assert o.threshold == 0.5 and o.arch_lr == 1e-4 and o.gan_mode == "hinge" and o.ndf == 128
    o = options.parse(["--dataroot", "x", "--lambda_scale", "0.01"])
    assert (o.n_epochs, o.n_epochs_decay) == (10, 15)
    with pytest.raises(NotImplementedError):
        options.parse(["--dataroot", "x", "--model", "nope"])


def test_cyclegan_prune_cfg_bit_exact_vs_golden(golden_dir):
    from gcc_b200 import prune
    g = torch.load(os.path.join(golden_dir, "cyclegan_prune.pt"), weights_only=False)
    R = {k: v.detach() for k, v in O._make_params(O.resnet_param_shapes(16), "P.netG.").items()}
    for thr in (0.8, 0.95, 1.05, 1.15):
        assert prune.cyclegan_prunenet_cfg(R, thr) == g["cfg@%g" % thr], thr
    assert prune.cyclegan_max_min_conv_norm(R) == pytest.approx(g["maxmin"], rel=1e-6)


def test_srgan_sagan_state_dict_layout_and_roundtrip():
    """SRGAN / SAGAN nets expose exactly the reference's state-dict names and shapes (the oracle's shape tables were
    loaded strictly into the reference modules by oracle/make_golden.py) and round-trip values, including the
    spectral-norm vector weight_v that is stored in channels-last column order internally."""
    from gcc_b200 import sagan, srgan
    from oracle import sagan_oracle as SA
    from oracle import srgan_oracle as SR

    def _check(net, shapes):        # same names and shapes; the ORDER of the keys is not part of the contract here
        sd = net.state_dict()
        assert sorted(sd.keys()) == sorted(shapes.keys())
        for k, s in shapes.items():
            assert tuple(sd[k].shape) == tuple(s), k
    cfg = [5, 8, 3, 7, 8, 6, 4, 8, 2, 8, 8, 1, 8, 5, 3, 8]
    _check(srgan.Generator(n_channels=8, device="cpu"), SR.sr_generator_shapes(8))
    _check(srgan.Generator(n_channels=8, filter_cfgs=cfg, device="cpu"), SR.sr_generator_shapes(8, cfg))
    _check(srgan.Discriminator(n_channels=8, device="cpu"), SR.sr_disc_shapes(8, False))
    _check(srgan.MaskDiscriminator(n_channels=8, device="cpu"), SR.sr_disc_shapes(8, True))
    _check(srgan.TruncatedVGG19(device="cpu"), SR.vgg_shapes())
    _check(sagan.Generator(ngf=8, device="cpu"), SA.generator_shapes(8))
    _check(sagan.Generator(ngf=8, filter_cfgs=[20, 12, 24, 8], device="cpu"), SA.generator_shapes(8, 128, [20, 12, 24, 8]))
    _check(sagan.Discriminator(ndf=8, device="cpu"), SA.disc_shapes(8, False))
    net = sagan.MaskDiscriminator(ndf=8, device="cpu")
    _check(net, SA.disc_shapes(8, True))
    P = SA.make_params(SA.disc_shapes(8, True), "t.netD.")
    net.load_state_dict({k: v.detach() for k, v in P.items()})
    sd = net.state_dict()
    assert all(torch.equal(sd[k], P[k].detach()) for k in P)
    # internal order of weight_v: (kh, kw, c) instead of the reference's (c, kh, kw)
    v_int = net.sn[1].v.detach()
    v_ref = P["l2.0.module.weight_v"].detach()
    assert torch.equal(v_int.reshape(16, 8).t().reshape(-1), v_ref)


def test_srgan_sagan_options():
    from gcc_b200 import options
    o = options.parse(["--model", "srgan", "--dataroot", "x"])
    assert (o.gan_mode, o.lr, o.batch_size, o.lr_policy, o.lr_decay_iters, o.lambda_SR_adversarial) == \
        ("vanilla", 1e-4, 16, "step", 15, 1e-3)
    o = options.parse(["--model", "sagan", "--dataroot", "x/celeb"])
    assert (o.gan_mode, o.lr, o.batch_size, o.crop_size, o.n_epochs, o.z_dim) == ("hinge", 1e-4, 64, 64, 100, 128)
    assert options.get_model_class(o).__name__ == "SAGANModel"


def test_random_pruned_cfgs_build_the_reference_shapes():
    """Arbitrary (non multiple-of-8) pruned widths, incl. dropped inner U-Net levels: the nets expose exactly the
    shapes the oracle's (= the reference's) shape tables give."""
    import random
    from gcc_b200 import nets, sagan, srgan
    from oracle import sagan_oracle as SA
    from oracle import srgan_oracle as SR
    rnd = random.Random(7)
    for trial in range(6):
        ngf = 8
        f = [rnd.randint(1, 40) for _ in range(15)]
        if trial % 3 == 2:                       # a zero entry drops the innermost level(s) (Pix2Pix.py:87,97)
            f[7] = 0
        c = [f[0], f[1], f[2], f[3], f[4], f[5], f[6]] + [0] * 8
        c[7] = f[7]
        skips = [f[6], f[5], f[4], f[3], f[2], f[1], f[0]]
        for i in range(7):
            c[8 + i] = f[8 + i] + skips[i]
        shapes = O.unet_param_shapes(ngf, f, c)
        sd = nets.UnetGenertor(ngf=ngf, filter_cfgs=f, channel_cfgs=c, device="cpu").state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(s) for k, s in shapes.items())
        rc = [rnd.randint(1, 33) for _ in range(23)]
        shapes = O.resnet_param_shapes(ngf, rc)
        sd = nets.MobileResnetGenerator(ngf=ngf, cfg=rc, device="cpu").state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(s) for k, s in shapes.items())
        sc = [rnd.randint(1, 19) for _ in range(16)]
        shapes = SR.sr_generator_shapes(ngf, sc)
        sd = srgan.Generator(n_channels=ngf, filter_cfgs=sc, device="cpu").state_dict()
        assert all(tuple(sd[k].shape) == tuple(s) for k, s in shapes.items()) and len(sd) == len(shapes)
        ac = [rnd.randint(1, 5) * 8 for _ in range(4)]      # attention needs channels divisible by 8 (in_dim // 8)
        shapes = SA.generator_shapes(ngf, 128, ac)
        sd = sagan.Generator(ngf=ngf, filter_cfgs=ac, device="cpu").state_dict()
        assert all(tuple(sd[k].shape) == tuple(s) for k, s in shapes.items()) and len(sd) == len(shapes)


def test_zero_pool_bump_allocation_and_reset():
    from gcc_b200 import ops
    pool = ops._ZeroPool()
    dev = torch.device("cpu")
    a = pool.take(10, dev)
    b = pool.take(3, dev)
    assert a.numel() == 10 and b.numel() == 3 and a.data_ptr() != b.data_ptr()
    assert (b.data_ptr() - a.data_ptr()) == 16 * 4            # 8-element granules
    a.add_(1.0)
    b.add_(2.0)
    pool.reset()
    assert float(a.sum()) == 0.0 and float(b.sum()) == 0.0      # the used prefix was re-zeroed
    assert pool.take(10, dev).data_ptr() == a.data_ptr()       # and the bump pointer rewound
    big = pool.take(1 << 21, dev)                              # larger than the pool: plain zeros, pool untouched
    assert big.numel() == 1 << 21 and float(big.abs().sum()) == 0.0
