"""`run_cuda_sampling` — drop-in for `run_rust_sampling` (alpharat/data/rust_sampling.py:137-296).

Same keyword arguments, same `(batch_dir, metrics)` return value, same ExperimentManager protocol
(`prepare_batch` before play, `register_batch` only after success, bundles under `batch_dir/games`).
Differences that come with `backend: cuda`:

* the `.pt` checkpoint is handed to the engine as it is (no ONNX export, rust_sampling.py:118-134);
* `device` names a GPU (`"cuda"`, `"cuda:3"`, `"b200"`, an int) instead of an ONNX execution provider;
* `num_threads` / `mux_max_batch_size` are accepted and ignored (no host worker pool, no mux);
* `mcts` is a `CudaMCTSConfig` (its `concurrent_games`, `pool_nodes`, `seed` are passed through); a
  `RustMCTSConfig` works too — only the shared search fields are read.

The reference package owns `ExperimentManager`; it is imported from `alpharat.experiments` when the caller
does not pass one (any object with `prepare_batch` / `register_batch` of the same signature will do, which
is also how the tests drive this function without the reference installed).
"""

from __future__ import annotations

import logging
import time
from dataclasses import dataclass, fields
from pathlib import Path
from typing import Any

from .selfplay import SelfPlayProgress, cuda_self_play, resolve_sampling_device

logger = logging.getLogger(__name__)

_SEARCH_FIELDS = (
    "simulations", "batch_size", "c_puct", "fpu_reduction", "force_k", "noise_epsilon", "noise_concentration",
    "collision_limit_min", "collision_limit_max", "collision_scaling_start", "collision_scaling_end",
    "collision_scaling_power",
)


def resolve_training_device(device: str | int) -> str:
    """PyTorch device for the training half of an iterate loop (rust_sampling.py:35-45)."""
    return f"cuda:{resolve_sampling_device(device)}"


@dataclass
class CudaSamplingMetrics:
    """Field-for-field twin of `RustSamplingMetrics` (rust_sampling.py:48-115)."""

    total_games: int
    total_positions: int
    total_simulations: int
    elapsed_seconds: float
    p1_wins: int
    p2_wins: int
    draws: int
    total_cheese_collected: float
    total_cheese_available: int
    min_turns: int
    max_turns: int
    total_nn_evals: int
    total_terminals: int
    total_collisions: int
    cache_hits: int
    cache_misses: int

    def _rate(self, x: float) -> float:
        return x / self.elapsed_seconds if self.elapsed_seconds > 0 else 0.0

    games_per_second = property(lambda s: s._rate(s.total_games))
    positions_per_second = property(lambda s: s._rate(s.total_positions))
    simulations_per_second = property(lambda s: s._rate(s.total_simulations))
    nn_evals_per_second = property(lambda s: s._rate(s.total_nn_evals))

    @property
    def avg_turns(self) -> float:
        return self.total_positions / self.total_games if self.total_games > 0 else 0.0

    @property
    def cheese_utilization(self) -> float:
        return self.total_cheese_collected / self.total_cheese_available if self.total_cheese_available else 0.0

    @property
    def draw_rate(self) -> float:
        return self.draws / self.total_games if self.total_games > 0 else 0.0

    @property
    def nn_eval_fraction(self) -> float:
        return self.total_nn_evals / self.total_simulations if self.total_simulations > 0 else 0.0

    @property
    def terminal_fraction(self) -> float:
        return self.total_terminals / self.total_simulations if self.total_simulations > 0 else 0.0

    @property
    def collision_fraction(self) -> float:
        total = self.total_nn_evals + self.total_terminals + self.total_collisions
        return self.total_collisions / total if total > 0 else 0.0

    @property
    def cache_hit_rate(self) -> float:
        total = self.cache_hits + self.cache_misses
        return self.cache_hits / total if total > 0 else 0.0


def self_play_kwargs(game: Any, mcts: Any, num_games: int) -> dict[str, Any]:
    """Nested `GameConfig` + MCTS config -> flat `cuda_self_play` keywords (rust_sampling.py:198-238).

    `game` is read duck-typed: `width, height, max_turns, positions, cheese.{count,symmetric}, maze.type`
    and, for random mazes, `maze.{wall_density,mud_density,symmetric}` (alpharat/config/game.py:46-52).
    """
    kwargs: dict[str, Any] = {
        "width": game.width,
        "height": game.height,
        "cheese_count": game.cheese.count,
        "max_turns": game.max_turns,
        "num_games": num_games,
        "maze_type": game.maze.type,
        "cheese_symmetric": game.cheese.symmetric,
        "positions": game.positions,
    }
    if getattr(game.maze, "type", None) == "random":
        kwargs["wall_density"] = game.maze.wall_density
        kwargs["mud_density"] = game.maze.mud_density
        kwargs["maze_symmetric"] = game.maze.symmetric
    for f in _SEARCH_FIELDS:
        kwargs[f] = getattr(mcts, f)
    for f in ("concurrent_games", "pool_nodes", "seed", "tree_engine"):  # CudaMCTSConfig only
        if hasattr(mcts, f):
            kwargs[f] = getattr(mcts, f)
    return kwargs


def run_cuda_sampling(
    *,
    game: Any,
    mcts: Any,
    num_games: int,
    group: str,
    num_threads: int = 4,
    max_games_per_bundle: int = 32,
    mux_max_batch_size: int = 256,
    checkpoint: str | None = None,
    device: str | int = "cuda",
    cache_size: int = 0,
    experiments_dir: str | Path = "experiments",
    verbose: bool = True,
    experiment_manager: Any | None = None,
    self_play_fn: Any = cuda_self_play,
) -> tuple[Path, CudaSamplingMetrics]:
    """Create a batch, play `num_games` games on the GPU into it, register it; returns `(batch_dir, metrics)`.

    Raises whatever `cuda_self_play` raises (`RuntimeError` engine, `IOError` bundle writing); the batch is
    registered only when play succeeded, exactly like the reference (rust_sampling.py:246-254).
    """
    exp = experiment_manager
    if exp is None:
        from alpharat.experiments import ExperimentManager  # type: ignore[import-not-found]

        exp = ExperimentManager(Path(experiments_dir))
    if getattr(mcts, "device_ids", None) and device in ("cuda", "auto"):
        device = mcts.device_ids[0]

    batch_dir, batch_uuid = exp.prepare_batch(group=group, mcts_config=mcts, game=game, checkpoint_path=checkpoint)
    output_dir = Path(batch_dir) / "games"

    kwargs = self_play_kwargs(game, mcts, num_games)
    kwargs.update(
        num_threads=num_threads,
        output_dir=str(output_dir),
        max_games_per_bundle=max_games_per_bundle,
        mux_max_batch_size=mux_max_batch_size,
        checkpoint=checkpoint,
        device=resolve_sampling_device(device),
        cache_size=cache_size,
    )
    stats = _play_reporting(self_play_fn, kwargs, num_games) if verbose else self_play_fn(**kwargs)

    exp.register_batch(group=group, batch_uuid=batch_uuid, mcts_config=mcts, game=game, checkpoint_path=checkpoint)

    # every metric is a same-named attribute of the stats object except the elapsed time
    renamed = {"elapsed_seconds": "elapsed_secs"}
    metrics = CudaSamplingMetrics(**{f.name: getattr(stats, renamed.get(f.name, f.name))
                                     for f in fields(CudaSamplingMetrics)})
    logger.info("backend cuda: %s", _summary_line(metrics))
    return Path(batch_dir), metrics


def _summary_line(m: CudaSamplingMetrics) -> str:
    parts = [f"{m.total_games} games", f"{m.total_positions} positions", f"{m.elapsed_seconds:.1f} s",
             f"{m.simulations_per_second:,.0f} sims/s", f"{m.games_per_second * 3600:,.0f} games/h",
             f"nn {m.nn_eval_fraction:.0%} / terminal {m.terminal_fraction:.0%} / collision {m.collision_fraction:.1%}"]
    if m.cache_hits + m.cache_misses:
        parts.append(f"cache hit {m.cache_hit_rate:.0%}")
    return ", ".join(parts)


def _play_reporting(self_play_fn: Any, kwargs: dict[str, Any], num_games: int, period_s: float = 2.0) -> Any:
    """Run the blocking self-play call on a worker and report the engine's live counters while it runs.

    The engine keeps `SelfPlayProgress` up to date from the device (host-mapped counters); this thread
    only reads it.  One log line per `period_s` with games done, rate and an ETA — no progress-bar
    dependency, so it behaves the same in a terminal, a notebook and a job log."""
    from concurrent.futures import ThreadPoolExecutor

    progress = SelfPlayProgress()
    t0 = time.monotonic()
    with ThreadPoolExecutor(max_workers=1, thread_name_prefix="cuda-self-play") as pool:
        fut = pool.submit(self_play_fn, **dict(kwargs, progress=progress))
        next_report = t0 + period_s
        while not fut.done():
            time.sleep(0.05)
            now = time.monotonic()
            if now < next_report:
                continue
            next_report = now + period_s
            done = progress.games_completed
            rate = done / (now - t0)
            eta = (num_games - done) / rate if rate > 0 else float("inf")
            logger.info("self-play %d/%d games, %.0f games/s, %d positions, eta %.0f s", done, num_games, rate,
                        progress.positions_completed, eta)
        return fut.result()  # re-raises the worker's exception, if any
