"""Multi-GPU self-play: one process per GPU, games sharded by index, no collective while playing.

The reference shards games over CPU threads by an atomic index (`game_worker_loop`,
crates/alpharat-sampling/src/selfplay.rs:609-650) and aggregates `SelfPlayStats` at the end
(selfplay.rs:212-224).  Across GPUs the same decomposition needs no data-path exchange: rank r
plays the contiguous index range `shard_range(num_games, r, world)` on its own engine.  NCCL
(gloo in CPU tests) is used only afterwards, to all-reduce the stats vector and to gather the
recorded game summaries / position records on rank 0.
"""

from __future__ import annotations

import os
from typing import Sequence

import numpy as np

SUM_KEYS = ("total_games", "total_positions", "total_simulations", "p1_wins", "p2_wins", "draws",
            "total_cheese_collected", "total_cheese_available", "total_nn_evals", "total_terminals",
            "total_collisions", "cache_hits", "cache_misses", "path_nodes", "new_nodes",
            "kernel_launches", "h2d_bytes", "d2h_bytes")
MAX_KEYS = ("max_turns", "elapsed_secs", "device_ms")
MIN_KEYS = ("min_turns",)


def shard_range(num_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, end) of `num_items` for `rank` of `world`."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(num_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return None
    return dist


def _device():
    import torch

    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_stats(stats: dict) -> dict:
    """Sum / max / min the per-rank stats exactly like `SelfPlayStats::add_game` would."""
    import torch

    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return dict(stats)
    out = dict(stats)
    dev = _device()
    for keys, op in ((SUM_KEYS, dist.ReduceOp.SUM), (MAX_KEYS, dist.ReduceOp.MAX), (MIN_KEYS, dist.ReduceOp.MIN)):
        present = [k for k in keys if k in stats]
        if not present:
            continue
        t = torch.tensor([float(stats[k]) for k in present], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        for k, v in zip(present, t.tolist()):
            out[k] = type(stats[k])(v) if isinstance(stats[k], int) else v
    return out


def gather_bytes(payload: np.ndarray, dst: int = 0) -> list[np.ndarray] | None:
    """Gather variable-length uint8 payloads (recorded game batches) on `dst`."""
    import torch

    dist = _dist()
    payload = np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
    if dist is None or dist.get_world_size() == 1:
        return [payload]
    dev = _device()
    world, rank = dist.get_world_size(), dist.get_rank()
    size = torch.tensor([payload.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    buf[: payload.size] = torch.from_numpy(payload.copy()).to(dev)
    bufs = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != dst:
        return None
    return [b[:n].cpu().numpy() for b, n in zip(bufs, sizes)]


def cuda_self_play_distributed(*, num_games: int, games: Sequence | None = None, seed: int | None = None,
                               gather_records: bool = False, **kwargs):
    """`cuda_self_play` over every rank of the initialised process group.

    Rank r plays games `shard_range(num_games, r, world)` on GPU LOCAL_RANK and writes its own
    bundles into `output_dir` (bundle names are uuids, so ranks never collide).  Returns the
    all-reduced `SelfPlayStats`-like dict on every rank and, with `gather_records`, the list of
    per-rank summary payloads on rank 0.
    """
    from .games import make_games
    from .selfplay import cuda_self_play

    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    lo, hi = shard_range(num_games, rank, world)
    if games is None:
        game_kw = {k: kwargs[k] for k in ("width", "height", "cheese_count", "max_turns")}
        games = make_games(hi - lo, first_index=lo, cheese_symmetric=kwargs.get("cheese_symmetric", True), **game_kw)
    else:
        games = list(games)[lo:hi]
    base_seed = (seed if seed is not None else 0) + lo
    device = int(os.environ.get("LOCAL_RANK", "0"))
    stats, summaries, _, _ = cuda_self_play(num_games=hi - lo, games=games, seed=base_seed, device=device,
                                            return_records=True, **kwargs)
    local = {k: getattr(stats, k) for k in SUM_KEYS + MAX_KEYS + MIN_KEYS if hasattr(stats, k)}
    total = allreduce_stats(local)
    payloads = None
    if gather_records:
        import ctypes as C

        nbytes = (hi - lo) * C.sizeof(summaries._type_)
        payloads = gather_bytes(np.frombuffer(bytes(summaries), dtype=np.uint8)[:nbytes])
    return total, payloads
