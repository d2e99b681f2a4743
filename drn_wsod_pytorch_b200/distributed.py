"""Multi-GPU plumbing of the hot path: images are independent units, one per rank per
step, weights replicated, no data-path collective.  The only exchange of the forward+loss path is
the mean of the loss dict -- the equivalent of detectron2/utils/comm.py:234-263 `reduce_dict`, done
as ONE all-reduce of a packed vector (NCCL over NVLink on the GPU box, gloo in the CPU tests).
A training step adds the mean of the trainable parameters' gradients (what DistributedDataParallel does for the
reference, detectron2/engine/defaults.py:279-282): `GradientSynchronizer` all-reduces each gradient block the
moment the backward has produced it, so the transfer of fc6's blocks overlaps the GEMMs of the next ones.
"""
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(num_items: int, rank: int, world_size: int) -> List[int]:
    """Indices of the images rank `rank` processes: item i goes to rank i % world_size (the reference's
    per-rank batch = IMS_PER_BATCH // world_size, detectron2/data/build.py:268-275)."""
    assert 0 <= rank < world_size
    return list(range(rank, num_items, world_size))


def pack_losses(losses: Dict[str, torch.Tensor]):
    keys = sorted(losses)
    return keys, torch.stack([losses[k].reshape(()) for k in keys])


def reduce_dict(losses: Dict[str, torch.Tensor], average: bool = True) -> Dict[str, torch.Tensor]:
    """All-reduce a dict of scalar tensors across ranks with one collective; every rank gets the result
    (keys must match on all ranks, as in the reference)."""
    rank, ws = world()
    if ws == 1:
        return dict(losses)
    keys, vec = pack_losses(losses)
    if average and dist.get_backend() == "nccl":
        dist.all_reduce(vec, op=dist.ReduceOp.AVG)  # averaged inside the collective
    else:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        if average:
            vec = vec / ws
    return {k: vec[i] for i, k in enumerate(keys)}


class LossReducer:
    """`reduce_dict` off the critical path: the mean of step i's loss dict is launched on a side stream behind step i
    and handed out when step i+1 submits (one step late -- what the reference's metric writer tolerates: it gathers the
    metrics of a step after the optimizer has moved on, detectron2/engine/train_loop.py:237-289).  The compute stream
    never waits for the collective, so a step costs what it costs on one GPU; `flush()` joins the last one."""

    def __init__(self, average: bool = True):
        self.average = average
        self.stream = None
        self.prev = None  # (keys, reduced vector, work handle or None)

    def _take(self):
        if self.prev is None:
            return None
        keys, vec, work = self.prev
        self.prev = None
        if work is not None:
            work.wait()
            if self.average and not vec.is_cuda:
                vec = vec / world()[1]
        elif vec.is_cuda:
            torch.cuda.current_stream(vec.device).wait_stream(self.stream)
        return {k: vec[i] for i, k in enumerate(keys)}

    def submit(self, losses: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Start the reduction of `losses`; returns the reduced dict of the PREVIOUS submit (this rank's own values on
        the first call and at world size 1)."""
        rank, ws = world()
        if ws == 1:
            return dict(losses)
        out = self._take()
        keys, vec = pack_losses({k: v.detach() for k, v in losses.items()})
        if vec.is_cuda:
            if self.stream is None or self.stream.device != vec.device:
                self.stream = torch.cuda.Stream(vec.device)
            self.stream.wait_stream(torch.cuda.current_stream(vec.device))
            vec.record_stream(self.stream)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(vec, op=dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM)
            self.prev = (keys, vec, None)
        else:
            work = dist.all_reduce(vec, op=dist.ReduceOp.SUM, async_op=True)
            self.prev = (keys, vec, work)
        return out if out is not None else dict(losses)

    def flush(self):
        """Reduced dict of the last submit (None if there is none pending); the current stream waits for it."""
        return self._take()


def gather_counts(values: Sequence[int]) -> List[List[int]]:
    """All-gather small per-rank integer tuples (images processed, proposals seen) for throughput accounting."""
    rank, ws = world()
    t = torch.tensor(list(values), dtype=torch.int64)
    if ws == 1:
        return [t.tolist()]
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


class GradientSynchronizer:
    """Data-parallel gradient averaging without DistributedDataParallel: the backward calls `ready(tensor)` for every
    finished gradient block (a whole parameter gradient, or a row block of fc6's); each call launches an
    asynchronous all-reduce right away -- NCCL orders it after the kernels already queued on the current stream and
    runs it on its own stream, next to the backward's remaining GEMMs.  The backward then `bind`s every parameter to
    the buffer its blocks live in; `finish()` waits for the collectives, applies the 1/world scale (AVG inside NCCL;
    SUM then scale with gloo) and only THEN hands the averaged buffers to `p.grad` (accumulating, like
    DistributedDataParallel does across ITER_SIZE backward passes).  Handing them to autograd earlier would let
    AccumulateGrad clone the local, un-averaged values while the collective is still in flight.  Unused parameters
    (bbox_pred without REFINE_REG) never produce a block, so nothing like find_unused_parameters is needed.

    Install with `attach(model)`; call `finish()` after `backward()` and before `optimizer.step()`."""

    def __init__(self, group=None):
        self.group = group
        self.pending = []
        self.bound = []  # (parameter, averaged gradient buffer) pairs of the backward passes since the last finish()
        self.bytes = 0   # all-reduced so far
        self.steps = 0   # finish() calls

    def attach(self, model):
        """Route the gradients of `model`'s trainable tail through this synchronizer (identity at world size 1)."""
        rh = getattr(model, "roi_heads", model)
        rh.grad_sync = self if world()[1] > 1 else None
        return self

    def ready(self, tensor):
        rank, ws = world()
        if ws == 1:
            return
        self.bytes += tensor.numel() * tensor.element_size()
        if dist.get_backend(self.group) == "nccl":
            h = dist.all_reduce(tensor, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self.pending.append((h, None))
        else:
            h = dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((h, tensor))

    def bind(self, param, grad):
        """`grad` (all of whose blocks were announced through `ready`) becomes `param.grad` in `finish()`."""
        self.bound.append((param, grad))

    def finish(self):
        rank, ws = world()
        self.steps += 1
        for h, t in self.pending:
            h.wait()
            if t is not None:
                t.div_(ws)
        n, self.pending = len(self.pending), []
        bound, self.bound = self.bound, []
        with torch.no_grad():
            for p, g in bound:
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad.add_(g)
        return n


# ------------------------------------------------------------------------------------------------------------------
# fc6.weight without an all-reduce: gradient tiles scattered into the owners' windows by the weight-gradient GEMM itself,
# sharded optimizer step, refreshed bf16 rows stored into every rank's weight buffer (csrc/drn_dist.cu, drn_tc.cu)
# ------------------------------------------------------------------------------------------------------------------
def _map_peers(tensor, group=None):
    """Raw device pointers of `tensor` on every rank of the group, as seen from THIS process: the local pointer for this
    rank, CUDA-IPC mappings (drn_peer_open) of the peers' allocations for the others.  Every rank must call this with a
    tensor of the same shape.  Returns (pointers, bases to close)."""
    import ctypes

    from . import lib

    L = lib.load()
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    handle = ctypes.create_string_buffer(L.drn_peer_handle_bytes())
    off = ctypes.c_uint64(0)
    lib.call("drn_peer_get_handle", tensor, handle, ctypes.byref(off))
    mine = (rank, bytes(handle.raw), int(off.value), tuple(tensor.shape), str(tensor.dtype))
    everyone = [None] * ws
    dist.all_gather_object(everyone, mine, group=group)
    ptrs, opened = [], []
    for r, h, o, shape, dt in everyone:
        assert shape == tuple(tensor.shape) and dt == str(tensor.dtype), "peer windows must match in shape and dtype"
        if r == rank:
            ptrs.append(tensor.data_ptr())
            continue
        base = ctypes.c_void_p(0)
        lib.call("drn_peer_open", ctypes.create_string_buffer(h, len(h)), ctypes.byref(base))
        opened.append(base.value)
        ptrs.append(base.value + o)
    return ptrs, opened


class ShardedLinearTrainer:
    """Data-parallel training of ONE big linear layer (fc6) whose weight gradient never goes through a collective.

    Rank r owns rows [r * rows/n, (r+1) * rows/n) of the weight.  Per step:
      1. backward: `wgrad(dy_t, x_t)` runs the weight-gradient GEMM dW = dY^T X with the reduce-scatter epilogue -- every rank's
         tiles land in the owners' `slots` windows over NVLink while the GEMM is still computing the next tiles;
      2. `fence()`: a 4-byte all-reduce on the compute stream -- when it has completed, every rank's GEMM has (its stores are
         performed at kernel completion), so the slots are final;
      3. `step(...)`: the owner sums its slots, updates its shard of the fp32 master weight and of the momentum, and stores the
         refreshed bf16 kernel-layout rows into EVERY rank's weight buffer (1/n of the optimizer traffic per rank, the all-gather
         rides the update);
      4. `fence()` again before the next forward reads the weight buffers (and before anyone overwrites slots).
    The fp32 master is current only on its owner; `gather_master()` all-gathers it (checkpointing).  Everything else of the
    model (fc7, heads: < 5 % of the gradient bytes) stays on the GradientSynchronizer's all-reduce, which doubles as the
    reference implementation this path is checked against (tools/train_sharded_check.py)."""

    def __init__(self, linear, precision, c49, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.linear, self.c49 = linear, int(c49)
        rows, cols = linear.weight.shape
        assert rows % (128 * self.world) == 0, f"{rows} weight rows do not split into {self.world} blocks of whole 128-row tiles"
        assert 1 <= self.world <= 8
        self.rows_per = rows // self.world
        self.row0 = self.rank * self.rows_per
        dev = linear.weight.device
        self.packed = linear.packed(precision, permute_c49=self.c49)["w"]  # the bf16 [rows, cols] buffer the forward reads
        self.slots = torch.zeros((self.world, self.rows_per, cols), dtype=torch.float32, device=dev)
        self.momentum = None
        self.first = True
        self.slot_ptrs, self._opened = _map_peers(self.slots, group)
        self.packed_ptrs, opened = _map_peers(self.packed, group)
        self._opened += opened
        self._token = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.bytes_scattered = 0
        self.fence()

    def wgrad(self, dy_t, x_t):
        """dW = dY^T X scattered: dy_t [rows, Rp] bf16 (zero beyond R), x_t [cols, Rp] bf16, both K-major."""
        from . import ops

        ops.gemm_scatter(dy_t, x_t, self.slot_ptrs, self.rank)
        self.bytes_scattered += (self.world - 1) * self.rows_per * x_t.shape[0] * 4

    def fence(self):
        dist.all_reduce(self._token, group=self.group)

    def step(self, lr, momentum, weight_decay, nesterov):
        from . import ops

        w = self.linear.weight
        if momentum != 0 and self.momentum is None:
            self.momentum = torch.empty((self.rows_per, w.shape[1]), dtype=torch.float32, device=w.device)
        with torch.no_grad():
            ops.sgd_step_sharded(w.detach(), self.momentum, self.slots, self.packed_ptrs, self.row0, self.rows_per, self.c49, lr,
                                 momentum, weight_decay, nesterov, self.first)
        self.first = False

    def gather_master(self):
        """All-gather the fp32 master shards so that `linear.weight` is complete on every rank (state_dict / checkpoint)."""
        w = self.linear.weight.detach()
        dist.all_gather_into_tensor(w, w[self.row0:self.row0 + self.rows_per].clone(), group=self.group)

    def close(self):
        from . import lib

        for b in self._opened:
            lib.call("drn_peer_close", b)
        self._opened = []
