"""alpharat_b200 — B200-native `backend: cuda` for alpharat's batched self-play MCTS.

Host-side mirror of the reference's interface for this path:

    CudaMCTSConfig   <- RustMCTSConfig          (alpharat/mcts/config.py:69-135)
    CudaSearcher     <- RustSearcher / Searcher (alpharat/mcts/searcher.py:21-117)
    cuda_self_play   <- rust_self_play          (crates/alpharat-sampling/src/bindings.rs:268-483)
    SelfPlayStats / SelfPlayProgress            (crates/alpharat-sampling/src/bindings.rs:28-201)
    run_cuda_sampling <- run_rust_sampling      (alpharat/data/rust_sampling.py:137-296)

Everything computes inside libalpharat_cuda.so (include/alpharat_cuda.h); there is no CPU path.
"""

from .config import CudaMCTSConfig, MCTSConfig, RustMCTSConfig
from .engine import Engine
from .games import GameSpec, make_games, pack_pod, pod_from_pyrat
from .result import SearchResult
from .searcher import CudaSearcher, Searcher
from .sampling import CudaSamplingMetrics, run_cuda_sampling
from .selfplay import SelfPlayProgress, SelfPlayStats, cuda_self_play

__all__ = [
    "CudaSamplingMetrics", "run_cuda_sampling", "CudaMCTSConfig", "CudaSearcher", "Engine", "GameSpec", "MCTSConfig", "RustMCTSConfig",
    "SearchResult", "Searcher", "SelfPlayProgress", "SelfPlayStats", "cuda_self_play",
    "make_games", "pack_pod", "pod_from_pyrat",
]
