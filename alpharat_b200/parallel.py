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


class _DeviceSpan:
    """A raw device allocation as `__cuda_array_interface__`, so that torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def device_bytes(ptr: int, nbytes: int):
    """uint8 torch tensor over `nbytes` of device memory at `ptr` (current CUDA device), zero-copy."""
    import torch

    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_DeviceSpan(ptr, nbytes), device="cuda")


def gather_device(t, dst: int = 0):
    """Gather variable-length uint8 DEVICE tensors on `dst` over the process group (NCCL: device to device).
    Returns the list of per-rank tensors on `dst`, None elsewhere."""
    import torch

    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return [t]
    world, rank = dist.get_world_size(), dist.get_rank()
    size = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    send = t if t.numel() == cap else torch.cat([t, torch.zeros(cap - t.numel(), dtype=torch.uint8, device=t.device)])
    recv = [torch.empty(cap, dtype=torch.uint8, device=t.device) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    return [b[:n] for b, n in zip(recv, sizes)]


def cuda_self_play_distributed(*, num_games: int, games: Sequence | None = None, seed: int | None = None,
                               gather_records: bool = False, **kwargs):
    """`cuda_self_play` over every rank of the initialised process group.

    Rank r plays games `shard_range(num_games, r, world)` on GPU LOCAL_RANK and writes its own
    bundles into `output_dir` (bundle names are uuids, so ranks never collide).  Returns the
    all-reduced `SelfPlayStats`-like dict on every rank and, with `gather_records`, on rank 0 the
    per-rank `(summaries_bytes, records_bytes)` payloads — packed position records gathered device to
    device over NCCL (game g of rank r owns `n_positions` consecutive records), numpy uint8 arrays.
    """
    from .engine import Engine
    from .selfplay import cuda_self_play

    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    lo, hi = shard_range(num_games, rank, world)
    if seed is None:
        # one entropy seed for the whole run, drawn on rank 0 (an unseeded single-GPU run uses entropy too)
        import secrets

        import torch

        t = torch.tensor([secrets.randbits(62) if rank == 0 else 0], dtype=torch.int64, device=_device())
        if dist is not None and world > 1:
            dist.broadcast(t, src=0)
        seed = int(t.item())
    device = int(os.environ.get("LOCAL_RANK", "0"))
    n_local = hi - lo
    eng = Engine(device=device, concurrent_games=max(1, min(kwargs.pop("concurrent_games", 4096), max(n_local, 1))),
                 pool_nodes=kwargs.pop("pool_nodes", 0), max_turns=kwargs["max_turns"],
                 max_batch_size=kwargs.get("batch_size", 8), max_simulations=kwargs["simulations"])
    try:
        # every generator keyword (maze_type, positions, densities, symmetry ...) reaches make_games through
        # cuda_self_play; the shard plays games [lo, hi) of the run: same layouts and seeds as a 1-GPU run
        stats = cuda_self_play(num_games=n_local, games=None if games is None else list(games)[lo:hi], seed=seed,
                               first_index=lo, device=device, engine=eng, **kwargs)
        local = {k: getattr(stats, k) for k in SUM_KEYS + MAX_KEYS + MIN_KEYS if hasattr(stats, k)}
        total = allreduce_stats(local)
        payloads = None
        if gather_records:
            import ctypes as C

            from . import _native as N

            summ, d_summ, d_rec, n_rec = eng.selfplay_pack_device(n_local)
            parts_s = gather_device(device_bytes(d_summ, n_local * C.sizeof(N.GameSummary)))
            parts_r = gather_device(device_bytes(d_rec, n_rec * C.sizeof(N.PositionRecord)))
            if parts_s is not None:
                payloads = [(a.cpu().numpy(), b.cpu().numpy()) for a, b in zip(parts_s, parts_r)]
        return total, payloads
    finally:
        eng.close()
