"""Multi-GPU plumbing: one process per GPU, clouds sharded over the batch axis, no data-path collective.

Every operator of the path is per-cloud (the reference launches grid = B everywhere) and BatchNorm is per-replica
under the reference's nn.DataParallel (utils.py:129-133), so inference needs no exchange at all and training
exchanges gradients only (28.3 MB all-reduce per step).  The reference's DataParallel additionally gathers the
393 MB all_feature tensor to GPU 0 each step; that gather does not exist here."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [lo, hi) slice of `total` units for `rank` (strong-scaling split of a fixed batch)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(rank, per_rank, stride=1000):
    """Weak-scaling shards: every rank generates its own `per_rank` clouds from disjoint seeds."""
    return range(stride * rank, stride * rank + per_rank)


def max_over_ranks(values, device):
    """Element-wise max over ranks of a list of floats (timings): the slowest rank defines the step time."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def wrap_ddp(model, device=None, find_unused_parameters=False):
    """Gradient all-reduce (average) overlapped with backward; BN stays per replica, like the reference.
    GripperRegionNetwork needs find_unused_parameters=True (linear_cls, and the refine head in pretrain_region:
    pointnet2.py:142, utils.py:106-107)."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    if device is not None and torch.device(device).type == "cuda":
        return DDP(model, device_ids=[torch.device(device).index], find_unused_parameters=find_unused_parameters)
    return DDP(model, find_unused_parameters=find_unused_parameters)
