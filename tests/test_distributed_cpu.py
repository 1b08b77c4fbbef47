"""CPU, world size 2 over gloo: the host-side data-parallel contract of the trainer
(reference deepof/clustering/dataset.py:591-618 and DDP's gradient all-reduce):
every rank derives the same shuffled list of contiguous batch starts and takes starts[rank::world];
ONE all-reduce(sum) of the flat gradient buffer followed by the 1/world scale equals DDP's mean."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import loader_oracle as LO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, bs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepof_b200.loader import batch_starts
        ok = True
        for epoch in (1, 2):
            mine = batch_starts(n, bs, epoch, seed=5, rank=rank, world=world)
            full = batch_starts(n, bs, epoch, seed=5)
            gathered = [None] * world
            dist.all_gather_object(gathered, mine.tolist())
            inter = np.stack([np.asarray(g) for g in gathered], 1).reshape(-1)
            ok &= np.array_equal(inter, full[:len(inter)]) and len(inter) == (len(full) // world) * world
            ok &= np.array_equal(mine, LO.batch_starts(n, bs, epoch, seed=5, rank=rank, world=world))
        # flat-gradient all-reduce(sum) + 1/world == mean over ranks (what DDP hands the optimizer)
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        ok &= torch.allclose(g / world, torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
        # parameters start identical: broadcast from rank 0
        p = torch.full((10,), float(rank))
        dist.broadcast(p, src=0)
        ok &= bool((p == 0).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world2_batch_sharding_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1000, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
