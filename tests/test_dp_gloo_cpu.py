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
