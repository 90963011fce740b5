"""Sample-parallel sharding helpers.

Posterior samples are sharded over ranks (one process per GPU); every rank holds
all inputs.  Position j of a call's sample list belongs to rank j % world, so the
samples of a smaller call are a prefix of each rank's share of a larger one.  The
only data-path collectives are sum-allreduces of [B, C] probability sums and
[B, D] gradient sums, issued through torch.distributed (NCCL on GPUs; gloo in
the CPU tests).
"""
import contextlib

import torch
import torch.distributed as dist

_replicated = 0        # > 0 inside `replicated()`: every rank works on its own, no collectives


def real_world():
    """(rank, world_size) of the process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def world():
    """(rank, world_size) the sample sharding sees: (0, 1) inside `replicated()`."""
    return (0, 1) if _replicated else real_world()


@contextlib.contextmanager
def replicated():
    """INPUT sharding (the alternative): inside this context every rank holds ALL posterior
    samples (the Philox draws are indexed by the global sample number, so all ranks draw identical banks without
    talking to each other) and works on its own slice of the inputs -- no data-path collective at all.  Used by
    `adversarialAttacks.attack`, whose PGD loop would otherwise all-reduce twice per iteration."""
    global _replicated
    _replicated += 1
    try:
        yield
    finally:
        _replicated -= 1


def all_gather_rows(local, per_rank, n_total):
    """Concatenate the ranks' row blocks (rank r holds rows [r * per_rank, min(n_total, (r + 1) * per_rank)))."""
    rank, w = real_world()
    if w == 1:
        return local
    pad = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(parts, pad)
    return torch.cat(parts)[:n_total]


def row_block(n_total, rank=None, world_size=None):
    """(lo, hi, per): the contiguous block of an n_total-row batch that `rank` moves over PCIe (host <-> device);
    per = rows per rank, the last blocks may be short or empty."""
    if rank is None:
        rank, world_size = world()
    per = (n_total + world_size - 1) // world_size if world_size > 1 else n_total
    return min(n_total, rank * per), min(n_total, (rank + 1) * per), per


def gather_rows_from_host(host, device, dtype):
    """Every rank needs ALL rows of `host` (a CPU tensor, ideally pinned) on its device.  One rank alone copies them
    all; with W ranks each one copies only its block of rows over PCIe and the blocks are all-gathered over NVLink
    (NCCL) -- W times less host traffic per rank, and no W-fold replication of it on the host's memory system."""
    rank, w = world()
    if w == 1 or not (dist.is_available() and dist.is_initialized()):
        return host.to(device=device, dtype=dtype, non_blocking=True)
    n = host.shape[0]
    lo, hi, per = row_block(n, rank, w)
    full = torch.empty((per * w,) + tuple(host.shape[1:]), dtype=dtype, device=device)
    mine = full[rank * per:(rank + 1) * per]
    if hi > lo:
        mine[:hi - lo].copy_(host[lo:hi], non_blocking=True)
    if hi - lo < per:
        mine[hi - lo:].zero_()
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(full, mine.clone())
    else:
        parts = [torch.empty_like(mine) for _ in range(w)]
        dist.all_gather(parts, mine.clone())
        full = torch.cat(parts)
    return full[:n]


def reduce_scatter_rows_(t_padded, per):
    """Sum over ranks of a [per * W, ...] tensor; returns this rank's [per, ...] block of the sum (NCCL reduce-scatter;
    all-reduce + slice on backends without it).  Identity for a single process / inside `replicated()`."""
    rank, w = world()
    if w == 1 or not (dist.is_available() and dist.is_initialized()):
        return t_padded[:per]
    if dist.get_backend() == "nccl":
        out = torch.empty((per,) + tuple(t_padded.shape[1:]), dtype=t_padded.dtype, device=t_padded.device)
        dist.reduce_scatter_tensor(out, t_padded, op=dist.ReduceOp.SUM)
        return out
    dist.all_reduce(t_padded, op=dist.ReduceOp.SUM)
    return t_padded[rank * per:(rank + 1) * per]


def local_positions(n, rank, world_size):
    """Positions of an n-long sample list owned by `rank`: rank, rank+W, rank+2W, ..."""
    return list(range(rank, n, world_size))


def local_count(n, rank, world_size):
    return len(range(rank, n, world_size))


def allreduce_sum_(t):
    """In-place sum over ranks; no-op for a single process and inside `replicated()`."""
    if _replicated:
        return t
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
