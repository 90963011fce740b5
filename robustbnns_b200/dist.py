"""Sample-parallel sharding helpers.

Posterior samples are sharded over ranks (one process per GPU); every rank holds
all inputs.  Position j of a call's sample list belongs to rank j % world, so the
samples of a smaller call are a prefix of each rank's share of a larger one.  The
only data-path collectives are sum-allreduces of [B, C] probability sums and
[B, D] gradient sums, issued through torch.distributed (NCCL on GPUs; gloo in
the CPU tests).
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def local_positions(n, rank, world_size):
    """Positions of an n-long sample list owned by `rank`: rank, rank+W, rank+2W, ..."""
    return list(range(rank, n, world_size))


def local_count(n, rank, world_size):
    return len(range(rank, n, world_size))


def allreduce_sum_(t):
    """In-place sum over ranks; no-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
