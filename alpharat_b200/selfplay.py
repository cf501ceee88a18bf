"""`cuda_self_play` — drop-in for `rust_self_play` (crates/alpharat-sampling/src/bindings.rs:268-483).

Same keyword arguments, same `SelfPlayStats` attributes (bindings.rs:28-158), same live
`SelfPlayProgress` (bindings.rs:167-201), same `bundle_*.npz` output.  Extra keywords:
`games=` (explicit `GameSpec`s), `seed=` (per-game RNG seeds = seed + game_index; the reference
seeds from entropy), `checkpoint=` (.pt, replaces `onnx_model_path`), `concurrent_games=`,
`engine=` (reuse a resident engine).
"""

from __future__ import annotations

import ctypes as C
import secrets
import time
from pathlib import Path
from typing import Sequence

from . import _native as N
from .engine import Engine, search_cfg
from .games import GameSpec, make_games, pods_array


def resolve_sampling_device(device: str | int) -> int:
    """GPU ordinal for a user-facing device name (rust_sampling.py:23-32 maps to ORT providers instead)."""
    if isinstance(device, int):
        return device
    d = device.lower()
    if d in ("auto", "cuda", "b200", "gpu", "tensorrt"):
        return 0
    for prefix in ("cuda:", "b200:", "gpu:"):
        if d.startswith(prefix) and d[len(prefix):].isdigit():
            return int(d[len(prefix):])
    raise ValueError(f"backend cuda cannot run on device {device!r} (expected 'cuda', 'cuda:<n>' or an int)")


class SelfPlayProgress:
    """Live counters, readable from another thread while `cuda_self_play` runs."""

    def __init__(self) -> None:
        self._p = N.Progress()

    games_completed = property(lambda s: int(s._p.games_completed))
    positions_completed = property(lambda s: int(s._p.positions_completed))
    simulations_completed = property(lambda s: int(s._p.simulations_completed))
    nn_evals_completed = property(lambda s: int(s._p.nn_evals_completed))


class SelfPlayStats:
    """Attribute-for-attribute twin of the reference's `SelfPlayStats` pyclass."""

    def __init__(self, s: N.Stats) -> None:
        self.total_games = int(s.total_games)
        self.total_positions = int(s.total_positions)
        self.total_simulations = int(s.total_simulations)
        self.elapsed_secs = float(s.elapsed_secs)
        self.p1_wins, self.p2_wins, self.draws = int(s.p1_wins), int(s.p2_wins), int(s.draws)
        self.total_cheese_collected = float(s.total_cheese_collected)
        self.total_cheese_available = int(s.total_cheese_available)
        self.min_turns, self.max_turns = int(s.min_turns), int(s.max_turns)
        self.total_nn_evals = int(s.total_nn_evals)
        self.total_terminals = int(s.total_terminals)
        self.total_collisions = int(s.total_collisions)
        self.cache_hits, self.cache_misses = int(s.cache_hits), int(s.cache_misses)
        # device-side extras (not in the reference)
        self.device_ms = float(s.device_ms)
        self.path_nodes, self.new_nodes = int(s.path_nodes), int(s.new_nodes)
        self.kernel_launches = int(s.kernel_launches)
        self.h2d_bytes, self.d2h_bytes = int(s.h2d_bytes), int(s.d2h_bytes)

    def _rate(self, x: float) -> float:
        return x / self.elapsed_secs if self.elapsed_secs > 0 else 0.0

    games_per_second = property(lambda s: s._rate(s.total_games))
    positions_per_second = property(lambda s: s._rate(s.total_positions))
    simulations_per_second = property(lambda s: s._rate(s.total_simulations))
    nn_evals_per_second = property(lambda s: s._rate(s.total_nn_evals))

    @property
    def cheese_utilization(self) -> float:
        return self.total_cheese_collected / self.total_cheese_available if self.total_cheese_available else 0.0

    @property
    def avg_turns(self) -> float:
        return self.total_positions / self.total_games if self.total_games else 0.0

    @property
    def draw_rate(self) -> float:
        return self.draws / self.total_games if self.total_games else 0.0

    @property
    def nn_eval_fraction(self) -> float:
        return self.total_nn_evals / self.total_simulations if self.total_simulations else 0.0

    @property
    def terminal_fraction(self) -> float:
        return self.total_terminals / self.total_simulations if self.total_simulations else 0.0

    @property
    def collision_fraction(self) -> float:
        t = self.total_nn_evals + self.total_terminals + self.total_collisions
        return self.total_collisions / t if t else 0.0

    @property
    def cache_hit_rate(self) -> float:
        t = self.cache_hits + self.cache_misses
        return self.cache_hits / t if t else 0.0


def cuda_self_play(
    *,
    width: int,
    height: int,
    cheese_count: int,
    max_turns: int,
    num_games: int,
    cheese_symmetric: bool = True,
    maze_type: str = "open",
    positions: str = "corners",
    wall_density: float = 0.7,
    mud_density: float = 0.1,
    maze_symmetric: bool = True,
    simulations: int,
    batch_size: int = 8,
    c_puct: float = 1.5,
    fpu_reduction: float = 0.2,
    force_k: float = 2.0,
    noise_epsilon: float = 0.0,
    noise_concentration: float = 10.83,
    collision_limit_min: int = 1,
    collision_limit_max: int = 256,
    collision_scaling_start: int = 800,
    collision_scaling_end: int = 50_000,
    collision_scaling_power: float = 1.0,
    num_threads: int = 4,  # accepted for signature parity; the GPU engine has no host worker pool
    output_dir: str | None,
    max_games_per_bundle: int = 32,
    onnx_model_path: str | None = None,
    device: str | int = "cuda",
    mux_max_batch_size: int = 256,
    cache_size: int = 0,
    progress: SelfPlayProgress | None = None,
    # --- extensions ---
    games: Sequence[GameSpec] | None = None,
    seed: int | None = None,
    first_index: int = 0,  # index of the first game in a larger run (multi-GPU shards)
    checkpoint: str | None = None,
    concurrent_games: int = 4096,
    pool_nodes: int = 0,
    engine: Engine | None = None,
    tree_engine: str = "warp",
    return_records: bool = False,
):
    """Play `num_games` games on the GPU and write bundles; returns `SelfPlayStats`."""
    if onnx_model_path is not None:
        raise ValueError("onnx_model_path is not used by backend cuda: pass checkpoint=<.pt>")
    if cache_size < 0:
        raise ValueError("cache_size must be >= 0")
    # Run seed: explicit `seed=` makes layouts and search reproducible; otherwise it comes from entropy, so
    # successive sampling runs (iterate: sample, train, sample) never replay the same boards — the reference
    # creates every game from entropy (config.create(None), bindings.rs:529-532).
    base_seed = seed if seed is not None else secrets.randbits(63)
    specs = list(games) if games is not None else make_games(
        num_games, width=width, height=height, cheese_count=cheese_count, max_turns=max_turns,
        cheese_symmetric=cheese_symmetric, maze_type=maze_type, positions=positions, wall_density=wall_density,
        mud_density=mud_density, maze_symmetric=maze_symmetric, first_index=first_index, layout_seed=base_seed)
    if games is not None and len(specs) != num_games:
        raise ValueError("len(games) != num_games")
    cfg = search_cfg(simulations=simulations, batch_size=batch_size, c_puct=c_puct,
                     fpu_reduction=fpu_reduction, force_k=force_k, noise_epsilon=noise_epsilon,
                     noise_concentration=noise_concentration, collision_limit_min=collision_limit_min,
                     collision_limit_max=collision_limit_max, collision_scaling_start=collision_scaling_start,
                     collision_scaling_end=collision_scaling_end, collision_scaling_power=collision_scaling_power)
    seeds = [(base_seed + first_index + i) & ((1 << 64) - 1) for i in range(len(specs))]
    dev = resolve_sampling_device(device)
    own = engine is None
    if own:
        mt = max([s.max_turns for s in specs] + [1])
        engine = Engine(device=dev, concurrent_games=max(1, min(concurrent_games, max(len(specs), 1))),
                        pool_nodes=pool_nodes, max_turns=mt, max_batch_size=batch_size,
                        max_simulations=simulations, tree_engine=tree_engine)
    try:
        if checkpoint is not None:
            from .weights import load_checkpoint_into

            load_checkpoint_into(engine, checkpoint)
        if engine.has_evaluator:
            engine.set_eval_cache(cache_size)  # per resident tree (the reference: per worker thread)
        t0 = time.perf_counter()
        pods = pods_array(specs)
        summaries, pos, stride, st = engine.selfplay(pods, cfg, seeds,
                                                     progress=progress._p if progress is not None else None)
        if output_dir is not None and len(specs) > 0:
            from .bundle import write_bundles

            try:
                write_bundles(Path(output_dir), specs, summaries, pos, stride, max_games_per_bundle)
            except OSError as e:  # IOError in the reference (bindings.rs:481)
                raise IOError(str(e)) from e
        st.elapsed_secs = time.perf_counter() - t0
        stats = SelfPlayStats(st)
        if return_records:
            return stats, summaries, pos, stride
        return stats
    finally:
        if own:
            engine.close()
