"""Multi-GPU host logic: the ensemble shards by contiguous system ranges, one process per GPU, no collective on the
hot path (systems never interact — reference src/main.rs:124-176 integrates one Universe per process)."""


def shard_range(n_total, rank, world):
    """Contiguous block of systems of `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_timing(elapsed_seconds, counts, dist=None):
    """max over ranks of the timed regions, sum over ranks of the counters (the only collective, after the timed region).

    elapsed_seconds / counts: torch tensors on the rank's device. With dist=None (single process) they are returned as is."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(elapsed_seconds, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return elapsed_seconds, counts


def gather_summaries(rows, dist=None):
    """Final gather of the per-system summaries (BASELINE north_star: the one place NCCL moves ensemble data).

    rows: [n_local_systems, k] tensor on the rank's device (status, time, dE/E, dL/L ...), the same shape on every rank.
    Returns the [world * n_local_systems, k] tensor in rank order on rank 0, None on the other ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return rows
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    out = torch.empty((world,) + tuple(rows.shape), dtype=rows.dtype, device=rows.device) if rank == 0 else None
    parts = list(out.unbind(0)) if rank == 0 else None
    dist.gather(rows.contiguous(), gather_list=parts, dst=0)
    return out.reshape((-1,) + tuple(rows.shape[1:])) if rank == 0 else None
