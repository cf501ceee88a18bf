"""MCTS configuration with a `backend` discriminator.

`RustMCTSConfig` restates the reference's fields (alpharat/mcts/config.py:69-84) so that batch
metadata written by either backend round-trips; `CudaMCTSConfig` is the new `backend: cuda`.
`MCTSConfig` is the discriminated union the reference would adopt (today it is a bare alias,
config.py:138).
"""

from __future__ import annotations

from typing import Annotated, Literal, Union

from pydantic import BaseModel, ConfigDict, Field


class StrictBaseModel(BaseModel):
    """extra='forbid', like alpharat/config/base.py."""

    model_config = ConfigDict(extra="forbid")


class _SearchFields(StrictBaseModel):
    simulations: int = 100
    c_puct: float = 1.5
    force_k: float = 2.0
    fpu_reduction: float = 0.2
    batch_size: int = 8
    noise_epsilon: float = 0.0
    noise_concentration: float = 10.83
    collision_limit_min: int = 1
    collision_limit_max: int = 256
    collision_scaling_start: int = 800
    collision_scaling_end: int = 50_000
    collision_scaling_power: float = 1.0

    def for_evaluation(self):
        """Copy with Dirichlet noise disabled (config.py:86-90)."""
        if self.noise_epsilon == 0.0:
            return self
        return self.model_copy(update={"noise_epsilon": 0.0})

    def search_kwargs(self) -> dict:
        return {k: getattr(self, k) for k in _SearchFields.model_fields}


class RustMCTSConfig(_SearchFields):
    """Field-compatible restatement of the reference config (kept for metadata round trips)."""

    backend: Literal["rust"] = "rust"


class CudaMCTSConfig(_SearchFields):
    """`backend: cuda` — same search fields plus engine placement."""

    backend: Literal["cuda"] = "cuda"
    device_ids: list[int] = Field(default_factory=lambda: [0])
    concurrent_games: int = 4096
    pool_nodes: int = 0  # 0 = auto: max_turns * simulations + 2 nodes per tree when that fits in 70 % of free HBM
    seed: int | None = None
    # uniform-prior self-play kernel: "warp" (a warp per tree), "half" (two trees per warp: the fastest once
    # concurrent_games is about 148 SMs x 32 warps x 2 = 9472 and the run has several games per tree) or "thread"
    tree_engine: Literal["warp", "half", "thread"] = "warp"

    def build_searcher(self, checkpoint: str | None = None, device: str = "cuda"):
        from .searcher import CudaSearcher

        return CudaSearcher(**self.search_kwargs(), checkpoint=checkpoint, seed=self.seed,
                            device=self.device_ids[0], pool_nodes=self.pool_nodes)

    def build_agent(self, checkpoint: str | None = None, temperature: float = 1.0, device: str = "cuda"):
        """The reference wraps the searcher in `SearcherAgent` (ai/searcher_agent.py:16-68);
        that class lives in the reference package, which consumes any `Searcher`."""
        from alpharat.ai.searcher_agent import SearcherAgent  # type: ignore[import-not-found]

        return SearcherAgent(searcher=self.build_searcher(checkpoint=checkpoint, device=device),
                             temperature=temperature, simulations=self.simulations, checkpoint=checkpoint)


MCTSConfig = Annotated[Union[RustMCTSConfig, CudaMCTSConfig], Field(discriminator="backend")]
