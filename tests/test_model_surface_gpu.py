"""Model-surface checks on the B200: the reference's Pix2PixModel API (eval forward for test.py, checkpoints,
prune(), L1 sparsity, dropout, LR schedule) on the CUDA path, against the oracle where arithmetic is involved."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = {"ngf": 8, "teacher_ngf": 16, "ndf": 16, "teacher_ndf": 16}


def _opt(extra=(), **over):
    from gcc_b200 import options
    opt = options.parse(["--dataroot", "x/cityscapes", "--gpu_ids", "0", "--lambda_content", "50", "--lambda_gram", "1e4",
                         *extra])
    for k, v in {**TINY, **over}.items():
        setattr(opt, k, v)
    return opt


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-20))


def test_eval_forward_matches_oracle(cuda):
    """test.py path: netG in eval mode (BatchNorm running statistics), no_grad, batch 1."""
    from gcc_b200.pix2pix import Pix2PixModel
    from oracle import gcc_oracle as O
    model = Pix2PixModel(_opt(["--no_dropout"]))
    S = O.Pix2PixOracle(O.Opt(direction="BtoA", online_distillation=False, darts_discriminator=False, **TINY), "S")
    # non-trivial running statistics
    for k, v in S.G.items():
        if k.endswith("running_mean"):
            v.copy_(O.det_normal("rm." + k, v.shape, 0.0, 0.3))
        if k.endswith("running_var"):
            v.copy_(O.det_normal("rv." + k, v.shape, 1.0, 0.1).abs() + 0.2)
    model.netG.load_state_dict({k: v.detach() for k, v in S.G.items()})
    model.model_eval()
    S.train_mode = False
    A = O.det_image("eval.A", 1, 3, 256, 256)
    B = O.det_image("eval.B", 1, 3, 256, 256)
    with torch.no_grad():
        model.set_input({"A": A, "B": B, "A_paths": "a", "B_paths": "b"})
        model.forward()
        S.set_input(A, B)
        S.forward()
    assert _rel(model.fake_B.cpu(), S.fake_B) < 3e-2
    vis = model.get_current_visuals()
    assert list(vis) == ["real_A", "fake_B", "real_B"] and vis["fake_B"].shape == (1, 3, 256, 256)
    # running statistics untouched in eval mode
    sd = model.netG.state_dict()
    assert all(torch.equal(sd[k].cpu(), v) for k, v in S.G.items() if "running" in k)


def test_checkpoint_roundtrip(cuda, tmp_path):
    from gcc_b200.pix2pix import Pix2PixModel
    opt = _opt(["--darts_discriminator", "--no_dropout"])
    a = Pix2PixModel(opt)
    a.save_models(3, str(tmp_path), fid=12.5)
    ckpt = torch.load(os.path.join(tmp_path, "model_3.pth"), map_location="cpu")
    # the reference's keys (models/Pix2Pix.py:636-647) plus ONE extra entry with the resume state (optimizer moments, ...)
    assert set(ckpt) == {"G", "D", "epoch", "cfg", "fid", "gcc_b200"} and ckpt["cfg"] == (None, None)
    w = ckpt["G"]["model.model.0.weight"]
    assert w.shape == (8, 3, 4, 4) and w.is_contiguous() and w.dtype == torch.float32  # reference layout: NCHW fp32
    b = Pix2PixModel(opt)
    fid, _ = b.load_models(os.path.join(tmp_path, "model_3.pth"))
    assert fid == 12.5
    for (k, x), (_, y) in zip(a.netG.state_dict().items(), b.netG.state_dict().items()):
        assert torch.equal(x, y), k
    for (k, x), (_, y) in zip(a.netD.state_dict().items(), b.netD.state_dict().items()):
        assert torch.equal(x, y), k
    # the bf16 operand packs follow the loaded weights
    x = torch.rand(1, 3, 256, 256) * 2 - 1
    with torch.no_grad():
        for m in (a, b):
            m.set_input({"A": x, "B": x})
            m.forward()
    # train-mode BN statistics are accumulated with fp32 atomics (order varies run to run): equal up to bf16 rounding
    assert _rel(a.fake_B, b.fake_B) < 1e-2


def test_prune_builds_pruned_model(cuda):
    """Pix2PixModel.prune(threshold) -> new model with the selected channel counts (Pix2Pix.py:742-752)."""
    from gcc_b200.pix2pix import Pix2PixModel
    from gcc_b200 import prune as P
    opt = _opt(["--scale_prune", "--no_dropout"], ngf=16)
    m = Pix2PixModel(opt)
    hi, lo = m.max_min_bn_scale()
    assert lo < hi
    thr = (hi + lo) / 2
    pruned = m.prune(thr)
    fc, cc = pruned.get_cfg()
    assert (fc, cc) == P.unet_scale_prune_cfg(m.netG.state_dict(), 16, thr)
    assert len(fc) == 15 and any(c % 8 for c in fc)          # arbitrary (non multiple-of-8) widths
    x = torch.rand(1, 3, 256, 256) * 2 - 1
    with torch.no_grad():
        pruned.set_input({"A": x, "B": x})
        pruned.forward()
    assert torch.isfinite(pruned.fake_B).all()
    with pytest.raises(NotImplementedError):
        Pix2PixModel(_opt(["--no_dropout"])).prune(0.5)           # neither --scale_prune nor --norm_prune


@pytest.mark.parametrize("mode", ["weight", "scale"])
def test_l1_sparsity_matches_oracle(cuda, mode):
    """Pre-training phase (scripts/pix2pix/pretrain_for_pruning.sh): grad += lambda * sign(w) (Pix2Pix.py:554-563)."""
    from gcc_b200.pix2pix import Pix2PixModel
    from oracle import gcc_oracle as O
    lam = {"lambda_weight": 1e-3} if mode == "weight" else {"lambda_scale": 1e-2}
    opt = _opt(["--no_dropout"], **lam)
    model = Pix2PixModel(opt)
    S = O.Pix2PixOracle(O.Opt(direction="BtoA", online_distillation=False, darts_discriminator=False, **TINY, **lam), "S")
    model.netG.load_state_dict({k: v.detach() for k, v in S.G.items()})
    model.netD.load_state_dict({k: v.detach() for k, v in S.D.items()})
    model.sync_weights()
    A, B = O.det_image("l1.A", 1, 3, 256, 256), O.det_image("l1.B", 1, 3, 256, 256)
    model.set_input({"A": A, "B": B})
    model.forward()
    model.set_requires_grad(model.netD, False)
    model.optimizer_G.zero_grad()
    model.backward_G()
    S.set_input(A, B)
    S.forward()
    S._set_grad(S.d_weights, False)
    S.opt_G.zero_grad()
    S.backward_G()
    name = "model.model.1.model.1.weight" if mode == "weight" else "model.model.1.model.2.weight"
    g = model.arena_G.grads[name].float().cpu()
    ref = S.G[name].grad
    assert _rel(g, ref) < 5e-2
    # the parameters that do NOT get the sparsity term still match (it is applied to the right tensors only)
    other = "model.model.1.model.2.weight" if mode == "weight" else "model.model.1.model.1.weight"
    # (a 16-element BatchNorm-scale gradient at bf16 activations: 5 % is its noise floor, see DESIGN.md tolerances)
    assert _rel(model.arena_G.grads[other].float().cpu(), S.G[other].grad) < 8e-2


def test_dropout_step_and_lr_schedule(cuda):
    """Dropout on (the reference default) changes the output between calls in train mode and the step stays finite;
    update_learning_rate follows the linear policy (utils/util.py:290-294)."""
    from gcc_b200.pix2pix import Pix2PixModel, build_teacher
    opt = _opt(["--darts_discriminator", "--online_distillation"])
    assert not opt.no_dropout
    model = Pix2PixModel(opt)
    build_teacher(model, opt)
    model.model_train()
    x = {"A": torch.rand(1, 3, 256, 256) * 2 - 1, "B": torch.rand(1, 3, 256, 256) * 2 - 1}
    with torch.no_grad():
        model.set_input(x)
        model.forward()
        f1 = model.fake_B
        model.forward()
        f2 = model.fake_B
    assert not torch.equal(f1, f2)
    model.set_input(x)
    model.optimize_parameters()
    model.clipping_mask_alpha()
    model.optimizer_netD_arch()
    losses = model.get_current_losses()
    assert all(torch.isfinite(torch.tensor(v)) for v in losses.values())
    assert list(losses) == ["G_GAN", "G_L1", "D_real", "D_fake", "D_arch_diff", "D_arch", "teacher_D_arch_diff", "content",
                            "gram"]
    opt.n_epochs, opt.n_epochs_decay = 2, 2
    m2 = Pix2PixModel(opt)
    lrs = []
    for e in range(1, 5):
        m2.update_learning_rate(e)
        lrs.append(m2.optimizers[0].param_groups[0]["lr"])
    assert lrs == pytest.approx([2e-4, 2e-4 * (1 - 1 / 3), 2e-4 * (1 - 2 / 3), 0.0], rel=1e-6, abs=1e-12)
    assert opt.ema_beta == pytest.approx(1.0 - 4 / 4)
