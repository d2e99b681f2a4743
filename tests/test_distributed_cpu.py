"""N>1 host logic on CPU: world_size-2 gloo processes.  The data path has no
collective; what is exercised here is image sharding and the one all-reduce of the loss dict."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drn_wsod_pytorch_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = D.shard_indices(5, rank, world)
        losses = {"loss_cls": torch.tensor(1.0 + rank), "loss_cls_r0": torch.tensor(10.0 * (rank + 1)), "loss_cls_r1": torch.tensor(0.5)}
        red = D.reduce_dict(losses)
        summed = D.reduce_dict(losses, average=False)
        counts = D.gather_counts([len(mine), 4000 * len(mine)])
        # gradient blocks announced one by one (as the backward produces them), averaged across ranks
        sync = D.GradientSynchronizer()
        g1, g2 = torch.full((4, 3), float(rank + 1)), torch.arange(6, dtype=torch.float32) * (rank + 1)
        sync.ready(g1[:2])
        sync.ready(g1[2:])
        sync.ready(g2)
        n = sync.finish()
        assert n == 3 and sync.finish() == 0
        assert torch.equal(g1, torch.full((4, 3), 1.5)) and torch.equal(g2, torch.arange(6, dtype=torch.float32) * 1.5)
        out.put((rank, mine, {k: float(v) for k, v in red.items()}, {k: float(v) for k, v in summed.items()}, counts))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_and_reduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]  # every image exactly once
    for rank, _, red, summed, counts in res:
        assert red == {"loss_cls": 1.5, "loss_cls_r0": 15.0, "loss_cls_r1": 0.5}
        assert summed == {"loss_cls": 3.0, "loss_cls_r0": 30.0, "loss_cls_r1": 1.0}
        assert counts == [[3, 12000], [2, 8000]]


def test_single_process_is_identity():
    losses = {"a": torch.tensor(2.0)}
    assert D.reduce_dict(losses)["a"].item() == 2.0
    assert D.shard_indices(3, 0, 1) == [0, 1, 2]
    sync = D.GradientSynchronizer()
    g = torch.ones(3)
    sync.ready(g)
    assert sync.finish() == 0 and torch.equal(g, torch.ones(3))
