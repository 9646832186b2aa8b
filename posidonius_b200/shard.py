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
