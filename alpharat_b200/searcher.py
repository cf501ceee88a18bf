"""`CudaSearcher` — the `Searcher` protocol (alpharat/mcts/searcher.py:21-25) on the GPU engine."""

from __future__ import annotations

import secrets
from typing import Any, Protocol, runtime_checkable

import numpy as np

from . import _native as N
from .engine import Engine, search_cfg
from .games import pod_from_pyrat
from .result import SearchResult


@runtime_checkable
class Searcher(Protocol):
    def search(self, game: Any) -> SearchResult: ...


def result_from_pod(r: N.SearchResultPod) -> SearchResult:
    """f32 -> f64 with renormalised policies, exactly as RustSearcher.search (searcher.py:97-117)."""
    p1 = np.asarray(r.policy_p1[:], dtype=np.float64)
    p2 = np.asarray(r.policy_p2[:], dtype=np.float64)
    s1, s2 = p1.sum(), p2.sum()
    if s1 > 0:
        p1 /= s1
    if s2 > 0:
        p2 /= s2
    return SearchResult(
        policy_p1=p1, policy_p2=p2, value_p1=float(r.value_p1), value_p2=float(r.value_p2),
        visit_counts_p1=np.asarray(r.visit_counts_p1[:], dtype=np.float64),
        visit_counts_p2=np.asarray(r.visit_counts_p2[:], dtype=np.float64),
        prior_p1=np.asarray(r.prior_p1[:], dtype=np.float64),
        prior_p2=np.asarray(r.prior_p2[:], dtype=np.float64),
        total_visits=int(r.total_visits),
    )


class CudaSearcher:
    """Fresh-tree search per call, like `rust_mcts_search` (mcts/bindings.rs:228-304).

    `game` is a `PyRat`-shaped object (duck-typed) or an `ar_game_pod`.  `search_many` runs many
    positions in one launch, one warp each.
    """

    def __init__(self, simulations: int, c_puct: float, force_k: float, fpu_reduction: float,
                 batch_size: int = 8, noise_epsilon: float = 0.0, noise_concentration: float = 10.83,
                 collision_limit_min: int = 1, collision_limit_max: int = 256,
                 collision_scaling_start: int = 800, collision_scaling_end: int = 50_000,
                 collision_scaling_power: float = 1.0, checkpoint: str | None = None,
                 seed: int | None = None, device: int = 0, pool_nodes: int = 0,
                 concurrent: int = 256, max_turns: int = 120) -> None:
        self._cfg = search_cfg(simulations=simulations, batch_size=batch_size, c_puct=c_puct,
                               fpu_reduction=fpu_reduction, force_k=force_k, noise_epsilon=noise_epsilon,
                               noise_concentration=noise_concentration,
                               collision_limit_min=collision_limit_min, collision_limit_max=collision_limit_max,
                               collision_scaling_start=collision_scaling_start,
                               collision_scaling_end=collision_scaling_end,
                               collision_scaling_power=collision_scaling_power)
        self._seed = seed
        self._simulations = simulations
        self._batch_size = max(batch_size, 1)
        self._device, self._pool_nodes, self._concurrent = device, pool_nodes, concurrent
        self._checkpoint = checkpoint
        self._engine: Engine | None = None
        self._engine_turns = 0
        self._ensure_engine(max_turns)

    # The engine's depth stack is sized by max_turns (hard limit of this build: 250 turns); it is rebuilt
    # when a game with a larger max_turns arrives, so any PyRat game the reference Searcher accepts is
    # accepted here as long as max_turns <= 250.
    def _ensure_engine(self, max_turns: int) -> None:
        if self._engine is not None and max_turns <= self._engine_turns:
            return
        if self._engine is not None:
            self._engine.close()
        turns = min(250, max(max_turns, 1))
        pn = self._pool_nodes or (self._simulations + 64)  # a fresh tree per call: at most `simulations` + 1 nodes
        self._engine = Engine(device=self._device, concurrent_games=self._concurrent, pool_nodes=max(pn, 64),
                              max_turns=turns, max_batch_size=self._batch_size,
                              max_simulations=self._simulations)
        self._engine_turns = turns
        if self._checkpoint is not None:
            from .weights import load_checkpoint_into

            load_checkpoint_into(self._engine, self._checkpoint)

    def _pod(self, game: Any) -> N.GamePod:
        return game if isinstance(game, N.GamePod) else pod_from_pyrat(game)

    def search_many(self, games: list[Any], seeds: list[int] | None = None) -> list[N.SearchResultPod]:
        pods = (N.GamePod * len(games))(*[self._pod(g) for g in games])
        self._ensure_engine(max([p.max_turns for p in pods] + [1]))
        if seeds is None:
            base = self._seed
            seeds = [base if base is not None else secrets.randbits(64) for _ in games]
        out = self._engine.search_batch(pods, self._cfg, seeds)
        return [out[i] for i in range(len(games))]

    def search(self, game: Any) -> SearchResult:
        return result_from_pod(self.search_many([game])[0])
