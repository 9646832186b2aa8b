"""world_size-2 gloo test of the multi-GPU host logic (sharding + the one reduction), run on CPU."""
import os
import socket

import pytest

from posidonius_b200.shard import shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n
                covered.extend(range(a, b))
            assert covered == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posidonius_b200.shard import reduce_timing, shard_range
    a, b = shard_range(1001, rank, world)
    elapsed = torch.tensor([0.5 + rank, 2.0 - rank], dtype=torch.float64)
    counts = torch.tensor([b - a, 1], dtype=torch.int64)
    dist.barrier()
    elapsed, counts = reduce_timing(elapsed, counts, dist)
    from posidonius_b200.shard import gather_summaries
    rows = torch.arange(12, dtype=torch.float64).reshape(4, 3) + 100.0 * rank
    g = gather_summaries(rows, dist)
    out.put((rank, elapsed.tolist(), counts.tolist(), None if g is None else g.tolist()))
    dist.destroy_process_group()


def test_reduce_timing_gloo_world2():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, elapsed, counts, gathered in results:
        assert elapsed == [1.5, 2.0]      # max over ranks
        assert counts == [1001, 2]        # all systems accounted for exactly once
        if rank == 0:
            # per-system summaries of both ranks, in rank order, on rank 0 only
            assert gathered == [[float(3 * i + c + 100 * r) for c in range(3)] for r in range(2) for i in range(4)]
        else:
            assert gathered is None
