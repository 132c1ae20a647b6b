"""GPU tests of the rows SURVEY.md 8(f) marks "next": the input pipeline in front of set_input (f4) and the pruning
front-end (f2: utils/prune_util.py:20-63 with gcc_b200.macs in place of thop)."""
import logging
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from gcc_b200 import _lib
    _lib.call("gcc_check_device")
    return True


def test_prefetcher_order_values_and_bytes(cuda):
    from gcc_b200.prefetch import Prefetcher
    g = torch.Generator().manual_seed(0)
    batches = [({"A": torch.rand(2, 3, 16, 16, generator=g) * 2 - 1, "B": torch.rand(2, 3, 16, 16, generator=g), "A_paths": "p%d" % i},
                {"A": torch.rand(2, 3, 16, 16, generator=g), "B": torch.rand(2, 3, 16, 16, generator=g), "A_paths": "v%d" % i})
               for i in range(7)]
    pf = Prefetcher(iter(batches), torch.device("cuda", 0), depth=2, dtype=torch.bfloat16)
    seen = 0
    for i, (tr, va) in enumerate(pf):
        assert tr["A_paths"] == "p%d" % i and va["A_paths"] == "v%d" % i
        for got, ref in ((tr, batches[i][0]), (va, batches[i][1])):
            for k in ("A", "B"):
                assert got[k].is_cuda and got[k].dtype == torch.bfloat16
                torch.cuda.current_stream().synchronize()
                assert torch.equal(got[k].float().cpu(), ref[k].to(torch.bfloat16).float()), (i, k)
        seen += 1
    assert seen == 7
    assert pf.h2d_bytes == 7 * 4 * 2 * 3 * 16 * 16 * 2          # bf16 staging: half the fp32 bytes
    pf.close()


def test_prefetched_batch_drives_the_model(cuda):
    """set_input accepts the prefetcher's bf16 device tensors: same losses as the fp32 host batch (the network rounds
    its input to bf16 anyway)."""
    from gcc_b200 import factory
    from gcc_b200.prefetch import Prefetcher
    tiny = dict(ngf=8, teacher_ngf=16, ndf=16, teacher_ndf=16)
    torch.manual_seed(3)
    a, _ = factory.build_pair(factory.make_opt("pix2pix", 0, ["--no_dropout"], **tiny))
    torch.manual_seed(3)
    b, _ = factory.build_pair(factory.make_opt("pix2pix", 0, ["--no_dropout"], **tiny))
    b.load_resume_state(a.resume_state())
    g = torch.Generator().manual_seed(5)
    tr, va = factory.synthetic_batch("pix2pix", 2, 256, g), factory.synthetic_batch("pix2pix", 2, 256, g)
    factory.run_iteration(a, tr, va)
    pf = Prefetcher(iter([(tr, va)]), torch.device("cuda", 0), dtype=torch.bfloat16)
    ptr, pva = next(pf)
    factory.run_iteration(b, ptr, pva)
    pf.close()
    la, lb = a.get_current_losses(), b.get_current_losses()
    for k in la:
        assert abs(la[k] - lb[k]) <= 2e-2 * abs(la[k]) + 5e-3, (k, la[k], lb[k])
    assert b.real_A.dtype == torch.float32


def test_binarysearch_threshold_and_prune(cuda, tmp_path):
    """prune_util.prune (utils/prune_util.py:49-63): load the pre-trained generator, bisect the BN-scale threshold
    until the pruned generator's MACs (thop units) are within 0.1 G of the budget, rebuild the model with that cfg."""
    from gcc_b200 import factory, macs, prune_util
    opt = factory.make_opt("pix2pix", 0, ["--scale_prune"], ngf=32, teacher_ngf=64, ndf=16, teacher_ndf=16)
    opt.online_distillation = False
    torch.manual_seed(0)
    model, _ = factory.build_pair(opt)
    # "pre-trained" generator: BatchNorm scales spread over (0, 1) so that thresholds matter
    with torch.no_grad():
        for name, p in model.netG.named_parameters():
            if p.dim() == 1 and name.endswith("weight"):
                p.copy_(torch.rand_like(p))
    model.sync_weights()
    model.save_models(7, str(tmp_path))
    full, _ = prune_util.get_flops_parms(model.netG, model.device, opt)
    assert full == pytest.approx(4.648, abs=0.01)          # U-Net ngf 32 in thop units
    opt.pretrain_path = os.path.join(str(tmp_path), "model_7.pth")
    opt.target_budget = 3.0                                # scripts/pix2pix/train.sh: --target_budget 3.0
    log = logging.getLogger("prune-test")
    fresh, _ = factory.build_pair(opt)
    pruned = prune_util.prune(fresh, opt, log)
    budget, params = prune_util.get_flops_parms(pruned.netG, pruned.device, opt)
    assert abs(budget - 3.0) <= 0.1
    f, c = pruned.get_cfg()
    assert len(f) == 15 and len(c) == 15 and all(isinstance(v, int) for v in f)
    assert budget == pytest.approx(macs.count_macs(pruned.netG, (256, 256), "thop") / 1e9)
    # the threshold search itself: the returned threshold reproduces a cfg within tolerance
    thr = prune_util.binarysearch_threshold(fresh, 3.0)
    b2, _ = prune_util.get_flops_parms(fresh.prune(thr).netG, fresh.device, opt)
    assert abs(b2 - 3.0) <= 0.1
    with pytest.raises(NotImplementedError):
        opt2 = factory.make_opt("pix2pix", 0, ["--scale_prune"], ngf=32, ndf=16)
        opt2.online_distillation = False
        prune_util.prune(fresh, opt2, log)                 # no target budget
