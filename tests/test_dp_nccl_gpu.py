"""Data parallel on real GPUs (NCCL, one process per GPU, world size 2; skipped on a single-GPU box):

  * replicas: after construction (rank 0's parameters are broadcast, gcc_b200/base.py) and two iterations on
    DIFFERENT per-rank batches, every parameter arena is bit-identical on both ranks -- eager and CUDA-graph replay.
  * global-batch parity (--sync_bn): 2 ranks x batch 2 == 1 device x batch 4 on the U-Net / BatchNorm model:
    synchronised BatchNorm statistics (forward and backward) and loss partial sums all-reduced before sqrt / abs
    (SURVEY.md 8e(2)).  Compared: all losses, and the averaged gradient arenas of student and teacher
    (rel-L2 <= 3e-2: same bf16 kernels, different summation order of the statistics).
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(ngf=8, teacher_ngf=16, ndf=16, teacher_ndf=16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def _batches(n, b, seed):
    from gcc_b200 import factory
    g = torch.Generator().manual_seed(seed)
    return [(factory.synthetic_batch("pix2pix", b, 256, g), factory.synthetic_batch("pix2pix", b, 256, g)) for _ in range(n)]


def _grads(model):
    out = {}
    for tag, m in (("S", model), ("T", model.teacher_model)):
        for name, a in m._gcc_arenas().items():
            out["%s.%s" % (tag, name)] = a.G.detach().float().cpu().clone()
    return out


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from gcc_b200 import factory
    try:
        if mode in ("replicas", "replicas_graph"):
            torch.manual_seed(100 + rank)            # different seeds: the constructor's broadcast must equalise
            opt = factory.make_opt("pix2pix", rank, [], **TINY)
            model, teacher = factory.build_pair(opt)
            data = _batches(3, 2, 50 + rank)
            if mode == "replicas_graph":
                from gcc_b200.graph import GraphedIteration
                gi = GraphedIteration(model).capture(data[0][0], data[0][1], warmup=1)
                for i in range(2):
                    gi.run(*data[1 + i])
            else:
                for i in range(2):
                    factory.run_iteration(model, *data[i])
            torch.cuda.synchronize()
            for m in (model, teacher):
                for name, a in m._gcc_arenas().items():
                    mine = a.P.detach().clone()
                    other = mine.clone()
                    dist.broadcast(other, 1)
                    if rank == 0:
                        assert torch.equal(mine, other), "arena %s differs between ranks (%s)" % (name, mode)
                        assert torch.isfinite(mine).all()
        else:                                        # global-batch parity
            torch.manual_seed(0)
            opt = factory.make_opt("pix2pix", rank, ["--no_dropout", "--sync_bn"], **TINY)
            model, teacher = factory.build_pair(opt)
            full = _batches(1, 4, 77)[0]
            shard = tuple({k: (v[2 * rank:2 * rank + 2] if torch.is_tensor(v) else v) for k, v in d.items()} for d in full)
            factory.run_iteration(model, *shard)
            torch.cuda.synchronize()
            if rank == 0:
                torch.save({"grads": _grads(model), "losses": model.get_current_losses(),
                            "bn": {k: v.cpu() for k, v in model.netD.state_dict().items() if "running" in k}},
                           os.path.join(out_dir, "dp2.pt"))
    finally:
        dist.destroy_process_group()


def _spawn(mode, out_dir="."):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), mode, out_dir), nprocs=2, join=True)


def test_replicas_stay_bit_identical_eager():
    _need2()
    _spawn("replicas")


def test_replicas_stay_bit_identical_graph_replay():
    _need2()
    _spawn("replicas_graph")


def test_two_ranks_b2_equal_one_device_b4_with_sync_bn(tmp_path):
    _need2()
    _spawn("parity", str(tmp_path))
    got = torch.load(os.path.join(str(tmp_path), "dp2.pt"), weights_only=False)
    from gcc_b200 import factory
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    opt = factory.make_opt("pix2pix", 0, ["--no_dropout"], **TINY)
    model, teacher = factory.build_pair(opt)
    full = _batches(1, 4, 77)[0]
    factory.run_iteration(model, *full)
    torch.cuda.synchronize()
    ref_l = model.get_current_losses()
    bad = []
    for k, v in ref_l.items():
        if abs(got["losses"][k] - v) > 2e-2 * abs(v) + 5e-3:
            bad.append(("loss", k, got["losses"][k], v))
    ref_g = _grads(model)
    for k, g in ref_g.items():
        a, b = got["grads"][k].double(), g.double()
        rel = float((a - b).norm() / (b.norm() + 1e-30))
        if rel > 3e-2:
            bad.append(("grad", k, rel))
    for k, v in model.netD.state_dict().items():
        if "running" in k:
            a, b = got["bn"][k].double(), v.cpu().double()
            rel = float((a - b).norm() / (b.norm() + 1e-30))
            if rel > 1e-2:
                bad.append(("bn", k, rel))
    assert not bad, bad
