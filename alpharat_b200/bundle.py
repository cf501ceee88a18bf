"""Bundle `.npz` writer — the 26 arrays the reference's training pipeline loads.

Format follows `write_bundle` (crates/alpharat-sampling/src/recording.rs:23-167): names, dtypes
(`<i4`, `i1`, `<i2`, `<f4`, bool), shapes, `bundle_{uuid}.npz`, at most `max_games_per_bundle`
games each (`BundleWriter`, recording.rs:174-230), atomic tmp -> rename.  Consumer:
`alpharat/data/loader.py:114-130`.
"""

from __future__ import annotations

import os
import uuid
from pathlib import Path
from typing import Sequence

import numpy as np

from . import _native as N
from .games import GameSpec, maze_array

BUNDLE_KEYS = (
    "game_lengths", "maze", "initial_cheese", "cheese_outcomes", "max_turns", "result",
    "final_p1_score", "final_p2_score",
    "p1_pos", "p2_pos", "p1_score", "p2_score", "p1_mud", "p2_mud", "cheese_mask", "turn",
    "value_p1", "value_p2", "visit_counts_p1", "visit_counts_p2", "prior_p1", "prior_p2",
    "policy_p1", "policy_p2", "action_p1", "action_p2",
)


def positions_as_numpy(positions, n_games: int, stride: int) -> np.ndarray:
    """View the ctypes `ar_position_record` array as a structured numpy array [n_games, stride]."""
    dt = np.dtype(N.PositionRecord)
    return np.frombuffer(positions, dtype=dt, count=n_games * stride).reshape(n_games, stride)


def _cheese_mask(bits: np.ndarray, h: int, w: int) -> np.ndarray:
    """u8[..., 32] bitfield -> bool[..., h, w] (cell = y*w + x)."""
    unpacked = np.unpackbits(bits, axis=-1, bitorder="little")[..., : h * w]
    return unpacked.reshape(*bits.shape[:-1], h, w).astype(np.bool_)


def summaries_as_numpy(summaries, n_games: int) -> np.ndarray:
    """View the ctypes `ar_game_summary` array as a structured numpy array [n_games]."""
    return np.frombuffer(summaries, dtype=np.dtype(N.GameSummary), count=n_games)


_MAZE_CACHE: dict = {}


def _maze_cached(spec: GameSpec) -> np.ndarray:
    """`maze_array` memoised on the layout (self-play batches share one maze, or a handful)."""
    key = (spec.width, spec.height, tuple(map(tuple, spec.walls)), tuple(map(tuple, spec.mud)))
    m = _MAZE_CACHE.get(key)
    if m is None:
        if len(_MAZE_CACHE) > 4096:
            _MAZE_CACHE.clear()
        m = _MAZE_CACHE[key] = maze_array(spec).astype(np.int8)
    return m


def build_bundle_arrays(specs: Sequence[GameSpec], summaries, pos_np: np.ndarray, idx: Sequence[int],
                        summ_np: np.ndarray | None = None) -> dict:
    """The 26 arrays of one bundle for the games `idx` (all of one board size).  Vectorised over games and
    positions: the only per-game Python work is the maze lookup and the initial cheese list."""
    idx = np.asarray(idx, dtype=np.int64)
    if summ_np is None:
        summ_np = summaries_as_numpy(summaries, int(idx.max()) + 1)
    h, w = specs[idx[0]].height, specs[idx[0]].width
    for i in idx:
        if (specs[i].height, specs[i].width) != (h, w):
            raise ValueError(f"game {i} has dimensions {specs[i].width}x{specs[i].height}, expected {w}x{h}")
    sm = summ_np[idx]
    lengths = sm["n_positions"].astype(np.int32)
    if (lengths == 0).any():
        raise ValueError(f"game {int(idx[np.argmin(lengths)])} has no positions")
    block = pos_np[idx]
    rows = block[np.arange(block.shape[1])[None, :] < lengths[:, None]]  # game-major, turn order inside a game
    sr = rows["search"]
    init = np.zeros((len(idx), h, w), dtype=np.bool_)
    for k, i in enumerate(idx):
        c = np.asarray(specs[i].cheese, dtype=np.int64).reshape(-1, 2)
        init[k, c[:, 1], c[:, 0]] = True
    return {
        "game_lengths": lengths,
        "maze": np.stack([_maze_cached(specs[i]) for i in idx]),
        "initial_cheese": init,
        "cheese_outcomes": sm["cheese_outcomes"][:, : h * w].astype(np.int8).reshape(len(idx), h, w),
        "max_turns": np.array([specs[i].max_turns for i in idx], dtype=np.int16),
        "result": sm["result"].astype(np.int8),
        "final_p1_score": sm["final_p1_score"].astype(np.float32),
        "final_p2_score": sm["final_p2_score"].astype(np.float32),
        "p1_pos": np.stack([rows["p1_x"], rows["p1_y"]], axis=1).astype(np.int8),
        "p2_pos": np.stack([rows["p2_x"], rows["p2_y"]], axis=1).astype(np.int8),
        "p1_score": rows["p1_score"].astype(np.float32),
        "p2_score": rows["p2_score"].astype(np.float32),
        "p1_mud": rows["p1_mud"].astype(np.int8),
        "p2_mud": rows["p2_mud"].astype(np.int8),
        "cheese_mask": _cheese_mask(np.ascontiguousarray(rows["cheese"]), h, w),
        "turn": rows["turn"].astype(np.int16),
        "value_p1": sr["value_p1"].astype(np.float32),
        "value_p2": sr["value_p2"].astype(np.float32),
        "visit_counts_p1": sr["visit_counts_p1"].astype(np.float32),
        "visit_counts_p2": sr["visit_counts_p2"].astype(np.float32),
        "prior_p1": sr["prior_p1"].astype(np.float32),
        "prior_p2": sr["prior_p2"].astype(np.float32),
        "policy_p1": sr["policy_p1"].astype(np.float32),
        "policy_p2": sr["policy_p2"].astype(np.float32),
        "action_p1": rows["action_p1"].astype(np.int8),
        "action_p2": rows["action_p2"].astype(np.int8),
    }


def write_bundle(path: Path, arrays: dict) -> None:
    tmp = path.with_suffix(".npz.tmp")
    with open(tmp, "wb") as f:
        np.savez_compressed(f, **arrays)
    os.replace(tmp, path)


def write_bundles(output_dir: Path, specs: Sequence[GameSpec], summaries, positions, stride: int,
                  max_games_per_bundle: int = 32, workers: int | None = None) -> list[Path]:
    """One `bundle_{uuid}.npz` per `max_games_per_bundle` games.  Bundles are independent files, so they are
    assembled and deflated on a small thread pool (zlib and the numpy copies release the GIL) — the reference
    writes them on a dedicated writer thread while the games are played (recording.rs:174-230)."""
    output_dir.mkdir(parents=True, exist_ok=True)
    n = len(specs)
    pos_np = positions_as_numpy(positions, n, stride)
    step = max(1, int(max_games_per_bundle))
    chunks = [list(range(lo, min(lo + step, n))) for lo in range(0, n, step)]
    paths = [output_dir / f"bundle_{uuid.uuid4()}.npz" for _ in chunks]

    summ_np = summaries_as_numpy(summaries, n)

    def one(k: int) -> None:
        write_bundle(paths[k], build_bundle_arrays(specs, summaries, pos_np, chunks[k], summ_np))

    n_workers = min(len(chunks), workers if workers is not None else min(4, os.cpu_count() or 1))
    if n_workers <= 1:
        for k in range(len(chunks)):
            one(k)
    else:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=n_workers) as pool:
            list(pool.map(one, range(len(chunks))))  # re-raises the first failure
    return paths
