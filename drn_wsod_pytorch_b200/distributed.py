"""Multi-GPU plumbing of the hot path: images are independent units, one per rank per
step, weights replicated, no data-path collective.  The only exchange of the forward+loss path is
the mean of the loss dict -- the equivalent of detectron2/utils/comm.py:234-263 `reduce_dict`, done
as ONE all-reduce of a packed vector (NCCL over NVLink on the GPU box, gloo in the CPU tests).
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
