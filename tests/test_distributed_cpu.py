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


class _FakeHeads:
    """Stands in for _WSLROIHeads in the autograd bridge: `_backward_device` produces rank-dependent gradients the way
    the real one does -- blocks announced through sync.ready as they are finished (fc6's in row blocks of one buffer,
    the heads' as row slices of one concatenated buffer), then every parameter bound to its buffer."""

    def __init__(self, params, rank):
        self.params, self.rank, self.grad_sync = params, rank, None

    def _backward_device(self, dev_out, grad_vec):
        w6, b6, wh = self.params
        scale = float(self.rank + 1) * grad_vec.sum()
        g6 = torch.empty_like(w6)
        sync = self.grad_sync
        for r0 in range(0, g6.shape[0], 2):
            g6[r0:r0 + 2] = scale * (r0 + 1)
            if sync is not None:
                sync.ready(g6[r0:r0 + 2])
        gcat = torch.full((5, wh.shape[1]), 1.0) * scale  # concatenated heads; this parameter owns rows 1..3
        gb = torch.full_like(b6, 3.0) * scale
        grads = {w6: g6, b6: gb, wh: gcat[1:4]}
        if sync is not None:
            sync.ready(gcat)
            sync.ready(gb)
            for p, g in grads.items():
                sync.bind(p, g)
        return grads


def _check_bridge_gradients(rank, world):
    """after a multi-rank backward every trainable parameter's .grad must be the MEAN gradient."""
    from drn_wsod_pytorch_b200.modeling import _WSLLossBridge

    params = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3, 6))]
    heads = _FakeHeads(params, rank)
    sync = D.GradientSynchronizer().attach(heads)
    assert heads.grad_sync is sync
    mean = 2.0 * sum(r + 1.0 for r in range(world)) / world  # d(vec.sum())/d(vec) = [1, 1]: the fake's scale is (rank + 1) * 2
    for it in range(2):  # two backward passes without zero_grad: gradients accumulate (ITER_SIZE)
        vec = _WSLLossBridge.apply(heads, {}, None, 0, torch.tensor([0.25, 0.75]), *params)
        vec.sum().backward()
        assert all(p.grad is None for p in params) or it > 0  # nothing reaches p.grad before the collectives finished
        sync.finish()
        k = it + 1
        want6 = torch.tensor([1.0, 1.0, 3.0, 3.0]).view(4, 1).expand(4, 3) * mean * k
        assert torch.allclose(params[0].grad, want6), (rank, params[0].grad)
        assert torch.allclose(params[1].grad, torch.full((4,), 3.0 * mean * k))
        assert torch.allclose(params[2].grad, torch.full((3, 6), mean * k))
    # without a synchronizer the bridge hands the local gradients to autograd as before
    heads.grad_sync = None
    for p in params:
        p.grad = None
    _WSLLossBridge.apply(heads, {}, None, 0, torch.tensor([1.0]), *params).sum().backward()
    assert torch.allclose(params[1].grad, torch.full((4,), 3.0 * (rank + 1)))


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
        _check_bridge_gradients(rank, world)
        # loss dict averaged one step late, off the critical path
        red_late = D.LossReducer()
        first = red_late.submit({"a": torch.tensor(1.0 + rank), "b": torch.tensor(4.0)})
        assert first["a"].item() == 1.0 + rank  # nothing reduced yet: this rank's own values
        second = red_late.submit({"a": torch.tensor(10.0 * (rank + 1)), "b": torch.tensor(0.0)})
        assert second["a"].item() == 1.5 and second["b"].item() == 4.0
        last = red_late.flush()
        assert last["a"].item() == 15.0 and red_late.flush() is None
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
    assert D.LossReducer().submit(losses)["a"].item() == 2.0 and D.LossReducer().flush() is None
    sync = D.GradientSynchronizer()
    g = torch.ones(3)
    sync.ready(g)
    assert sync.finish() == 0 and torch.equal(g, torch.ones(3))
