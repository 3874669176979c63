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


class FlatGrads:
    """All gradients of a parameter list as views into ONE contiguous fp32 buffer, reduced with ONE all-reduce.

    This is the whole multi-GPU exchange of the training path ("NCCL gradient all-reduce only"): autograd accumulates
    into the pre-assigned `.grad` views in place, so after backward the buffer holds every gradient without a bucket
    copy; `all_reduce()` averages it over the ranks (NCCL: one ring/NVLS pass over ~22 MB / ~6 MB for the two REGNet
    networks); parameters that took no part in a step simply keep a zero gradient (Adam then leaves them unchanged),
    which is what DistributedDataParallel needs `find_unused_parameters=True` and a graph walk for.
    Replaces, per process, the gradient gather of the reference's nn.DataParallel (utils.py:129-133)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(total, dtype=first.dtype, device=first.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def zero(self):
        """Instead of optimizer.zero_grad(): keeps the views, one memset."""
        for p in self.params:             # somebody (zero_grad(set_to_none=True)) dropped a view: re-attach them all
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + max(self.nbytes, 1):
                self._reattach()
                break
        self.flat.zero_()

    def _reattach(self):
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def all_reduce(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:                              # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
