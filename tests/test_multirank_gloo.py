"""world_size-2 gloo test of the multi-GPU host logic (sharding, stats all-reduce, record gather)."""

from __future__ import annotations

import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from alpharat_b200.parallel import shard_range


def test_shard_range_partitions_every_index():
    for n in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank: int, world: int, port: int, q) -> None:
    import torch.distributed as dist

    from alpharat_b200.parallel import allreduce_stats, gather_bytes, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(101, rank, world)
        stats = {"total_games": hi - lo, "total_positions": 10 * (hi - lo), "min_turns": 5 + rank,
                 "max_turns": 20 + rank, "elapsed_secs": 1.0 + rank, "total_cheese_collected": 2.5 * (hi - lo)}
        total = allreduce_stats(stats)
        payload = np.arange(lo, hi, dtype=np.uint8)  # stands in for this rank's game summaries
        gathered = gather_bytes(payload, dst=0)
        q.put((rank, total, None if gathered is None else [g.tolist() for g in gathered]))
    finally:
        dist.destroy_process_group()


def test_two_rank_reduce_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, total, gathered = q.get(timeout=120)
        res[rank] = (total, gathered)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        total = res[rank][0]
        assert total["total_games"] == 101 and total["total_positions"] == 1010
        assert total["min_turns"] == 5 and total["max_turns"] == 21 and total["elapsed_secs"] == 2.0
        assert abs(total["total_cheese_collected"] - 252.5) < 1e-9
    assert res[1][1] is None
    assert res[0][1] == [list(range(0, 51)), list(range(51, 101))]
