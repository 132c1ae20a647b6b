"""Data-parallel host logic on CPU with the gloo backend (world size 2): the flat gradient arena is
averaged across ranks before each optimizer step and initial weights are replicated from rank 0."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gcc_b200.arena import ParamArena
    from gcc_b200.pix2pix import _allreduce_grads
    A = ParamArena("cpu")
    A.add("w", (4, 3, 2, 2), "conv")
    A.add("b", (5,), "vec")
    A.finalize()
    with torch.no_grad():
        A.params["w"].fill_(float(rank + 1))
        A.params["b"].fill_(float(10 * (rank + 1)))
    dist.broadcast(A.P, 0)  # what bench.py / a DP driver does once after construction
    A.grads["w"].fill_(float(rank))          # rank-dependent gradient
    A.grads["b"].copy_(torch.arange(5.0) * (rank + 1))
    _allreduce_grads(A)
    out[rank] = tuple(t.detach().clone().contiguous().tolist() for t in (A.params["w"], A.params["b"], A.grads["w"], A.grads["b"]))
    dist.destroy_process_group()


def test_arena_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        w, b, gw, gb = (torch.tensor(t) for t in out[r])
        assert torch.all(w == 1.0) and torch.all(b == 10.0)           # replicated from rank 0
        assert torch.allclose(gw, torch.full_like(gw, 0.5))             # mean of {0, 1}
        assert torch.allclose(gb, torch.arange(5.0) * 1.5)              # mean of {1x, 2x}


def _worker_multi(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gcc_b200.arena import ParamArena
    from gcc_b200.sagan import _MultiArenaOptimizer
    arenas = []
    for i in range(2):
        A = ParamArena("cpu", betas=(0.0, 0.9))
        A.add("p%d" % i, (6,), "vec")
        A.finalize()
        A.grads["p%d" % i].fill_(float(rank + 1 + i))
        A.steps = 0
        A.step = (lambda a: (lambda: setattr(a, "steps", a.steps + 1)))(A)   # the Adam kernel needs a GPU: count calls
        arenas.append(A)
    opt = _MultiArenaOptimizer([(arenas[0], 1), (arenas[1], 2)], 4e-4, (0.0, 0.9))
    opt.step()
    out[rank] = (arenas[0].steps, arenas[1].steps, arenas[0].G.tolist(), arenas[1].G.tolist(), arenas[0].lr, arenas[1].betas)
    dist.destroy_process_group()


def test_multi_arena_optimizer_world2():
    """SAGAN's optimizers (gcc_b200.sagan._MultiArenaOptimizer): every arena is averaged across ranks once, then
    stepped as many times as the reference's parameter list holds its tensors (SAGAN.py:318-326,339-344)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_multi, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        s0, s1, g0, g1, lr, betas = out[r]
        assert (s0, s1) == (1, 2)
        assert all(abs(v - 1.5) < 1e-6 for v in g0[:6])     # mean of {1, 2}
        assert all(abs(v - 2.5) < 1e-6 for v in g1[:6])     # mean of {2, 3}
        assert lr == 4e-4 and tuple(betas) == (0.0, 0.9)
