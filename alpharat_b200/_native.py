"""ctypes mirror of include/alpharat_cuda.h and the loader of libalpharat_cuda.so.

The library is the product: there is no CPU fallback.  `load_library()` raises
`RuntimeError` when the shared object has not been built (run
`python -c "import __graft_entry__ as g; g.build()"` or `python -m alpharat_b200.build`).
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

AR_ABI_VERSION = 2
AR_TREE_WARP, AR_TREE_THREAD, AR_TREE_HALF = 0, 1, 2
AR_MAX_CELLS = 256

AR_OK = 0
AR_ERR_INVALID_ARG = 1
AR_ERR_CUDA = 2
AR_ERR_POOL_OVERFLOW = 3
AR_ERR_NONFINITE = 4
AR_ERR_UNSUPPORTED = 5
AR_ERR_NO_WEIGHTS = 6

AR_ARCH_UNIFORM = 0
AR_ARCH_MLP = 1
AR_ARCH_SYMMETRIC = 2
AR_ARCH_CNN = 3


class GamePod(C.Structure):
    _fields_ = [
        ("width", C.c_uint8), ("height", C.c_uint8),
        ("p1_x", C.c_uint8), ("p1_y", C.c_uint8), ("p2_x", C.c_uint8), ("p2_y", C.c_uint8),
        ("p1_mud", C.c_uint8), ("p2_mud", C.c_uint8),
        ("turn", C.c_uint16), ("max_turns", C.c_uint16),
        ("reserved0", C.c_uint16), ("reserved1", C.c_uint16),
        ("p1_score", C.c_float), ("p2_score", C.c_float),
        ("move_cost", C.c_uint8 * (AR_MAX_CELLS * 4)),
        ("cheese", C.c_uint8 * (AR_MAX_CELLS // 8)),
    ]


class SearchCfg(C.Structure):
    _fields_ = [
        ("simulations", C.c_uint32), ("batch_size", C.c_uint32),
        ("c_puct", C.c_float), ("fpu_reduction", C.c_float), ("force_k", C.c_float),
        ("noise_epsilon", C.c_float), ("noise_concentration", C.c_float),
        ("collision_limit_min", C.c_uint32), ("collision_limit_max", C.c_uint32),
        ("collision_scaling_start", C.c_uint32), ("collision_scaling_end", C.c_uint32),
        ("collision_scaling_power", C.c_float),
    ]


class SearchResultPod(C.Structure):
    _fields_ = [
        ("policy_p1", C.c_float * 5), ("policy_p2", C.c_float * 5),
        ("value_p1", C.c_float), ("value_p2", C.c_float),
        ("visit_counts_p1", C.c_float * 5), ("visit_counts_p2", C.c_float * 5),
        ("prior_p1", C.c_float * 5), ("prior_p2", C.c_float * 5),
        ("total_visits", C.c_uint32), ("nn_evals", C.c_uint32),
        ("terminals", C.c_uint32), ("collisions", C.c_uint32),
        ("raw_visits_p1", C.c_uint32 * 5), ("raw_visits_p2", C.c_uint32 * 5),
        ("node_count", C.c_uint32), ("reserved", C.c_uint32),
    ]


class PositionRecord(C.Structure):
    _fields_ = [
        ("p1_x", C.c_uint8), ("p1_y", C.c_uint8), ("p2_x", C.c_uint8), ("p2_y", C.c_uint8),
        ("p1_mud", C.c_uint8), ("p2_mud", C.c_uint8),
        ("action_p1", C.c_uint8), ("action_p2", C.c_uint8),
        ("turn", C.c_uint16), ("reserved", C.c_uint16),
        ("p1_score", C.c_float), ("p2_score", C.c_float),
        ("search", SearchResultPod),
        ("cheese", C.c_uint8 * (AR_MAX_CELLS // 8)),
    ]


class GameSummary(C.Structure):
    _fields_ = [
        ("game_index", C.c_uint32), ("n_positions", C.c_uint32),
        ("final_p1_score", C.c_float), ("final_p2_score", C.c_float),
        ("result", C.c_uint8), ("reserved", C.c_uint8 * 3),
        ("cheese_available", C.c_uint16), ("reserved1", C.c_uint16),
        ("total_simulations", C.c_uint64), ("total_nn_evals", C.c_uint64),
        ("total_terminals", C.c_uint64), ("total_collisions", C.c_uint64),
        ("cheese_outcomes", C.c_uint8 * AR_MAX_CELLS),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("total_games", C.c_uint32),
        ("p1_wins", C.c_uint32), ("p2_wins", C.c_uint32), ("draws", C.c_uint32),
        ("total_positions", C.c_uint64), ("total_simulations", C.c_uint64),
        ("total_nn_evals", C.c_uint64), ("total_terminals", C.c_uint64),
        ("total_collisions", C.c_uint64),
        ("cache_hits", C.c_uint64), ("cache_misses", C.c_uint64),
        ("elapsed_secs", C.c_double),
        ("total_cheese_collected", C.c_float), ("total_cheese_available", C.c_uint32),
        ("min_turns", C.c_uint32), ("max_turns", C.c_uint32),
        ("device_ms", C.c_double),
        ("path_nodes", C.c_uint64), ("new_nodes", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
    ]


class Progress(C.Structure):
    _fields_ = [
        ("games_completed", C.c_uint32), ("reserved", C.c_uint32),
        ("positions_completed", C.c_uint64),
        ("simulations_completed", C.c_uint64),
        ("nn_evals_completed", C.c_uint64),
    ]


class EngineCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32),
        ("concurrent_games", C.c_uint32), ("pool_nodes", C.c_uint32),
        ("max_cells", C.c_uint32), ("max_turns", C.c_uint32),
        ("max_batch_size", C.c_uint32), ("max_simulations", C.c_uint32),
        ("tree_engine", C.c_uint32),
    ]


class TensorDesc(C.Structure):
    _fields_ = [
        ("name", C.c_char_p), ("data", C.POINTER(C.c_float)),
        ("ndim", C.c_int32), ("shape", C.c_int64 * 4),
    ]


EXPORTED_SYMBOLS = (
    "ar_engine_create", "ar_engine_destroy", "ar_last_error", "ar_abi_version",
    "ar_engine_load_weights", "ar_search_batch", "ar_selfplay_run", "ar_selfplay_upload",
    "ar_selfplay_run_resident", "ar_selfplay_download", "ar_encode_observations",
    "ar_nn_forward", "ar_engine_set_eval_cache",
    "ar_stream_open", "ar_stream_close", "ar_stream_submit", "ar_stream_collect", "ar_stream_launch",
    "ar_stream_wait", "ar_stream_elapsed_ms", "ar_stream_times", "ar_selfplay_pack_device",
)

_LIB = None


def library_path() -> Path:
    env = os.environ.get("ALPHARAT_CUDA_LIB")
    if env:
        return Path(env)
    return Path(__file__).resolve().parent / "libalpharat_cuda.so"


def load_library() -> C.CDLL:
    """Load libalpharat_cuda.so and declare its prototypes.  Fails loudly when absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: the CUDA engine has not been built. "
            "Run `python -m alpharat_b200.build` (needs nvcc). There is no CPU fallback."
        )
    lib = C.CDLL(str(path))
    P = C.POINTER
    eng = C.c_void_p
    lib.ar_abi_version.restype = C.c_uint32
    lib.ar_engine_create.argtypes = [P(EngineCfg), P(eng)]
    lib.ar_engine_create.restype = C.c_int
    lib.ar_engine_destroy.argtypes = [eng]
    lib.ar_engine_destroy.restype = None
    lib.ar_last_error.argtypes = [eng]
    lib.ar_last_error.restype = C.c_char_p
    lib.ar_engine_load_weights.argtypes = [eng, C.c_int32, C.c_int32, C.c_int32, P(TensorDesc), C.c_int32]
    lib.ar_engine_load_weights.restype = C.c_int
    lib.ar_search_batch.argtypes = [eng, P(GamePod), C.c_int32, P(SearchCfg), P(C.c_uint64), P(SearchResultPod)]
    lib.ar_search_batch.restype = C.c_int
    lib.ar_selfplay_run.argtypes = [eng, P(GamePod), C.c_int32, P(SearchCfg), P(C.c_uint64),
                                    P(GameSummary), P(PositionRecord), C.c_int32, P(Progress), P(Stats)]
    lib.ar_selfplay_run.restype = C.c_int
    lib.ar_selfplay_upload.argtypes = [eng, P(GamePod), C.c_int32, P(C.c_uint64)]
    lib.ar_selfplay_upload.restype = C.c_int
    lib.ar_selfplay_run_resident.argtypes = [eng, P(SearchCfg), P(Stats)]
    lib.ar_selfplay_run_resident.restype = C.c_int
    lib.ar_selfplay_download.argtypes = [eng, P(GameSummary), P(PositionRecord), C.c_int32]
    lib.ar_selfplay_download.restype = C.c_int
    lib.ar_encode_observations.argtypes = [eng, P(GamePod), C.c_int32, P(C.c_float)]
    lib.ar_encode_observations.restype = C.c_int
    lib.ar_nn_forward.argtypes = [eng, P(GamePod), C.c_int32, P(C.c_float), P(C.c_float), P(C.c_float), P(C.c_float)]
    lib.ar_nn_forward.restype = C.c_int
    lib.ar_engine_set_eval_cache.argtypes = [eng, C.c_uint32]
    lib.ar_engine_set_eval_cache.restype = C.c_int
    lib.ar_selfplay_pack_device.argtypes = [eng, P(GameSummary), P(C.c_void_p), P(C.c_void_p), P(C.c_uint64)]
    lib.ar_selfplay_pack_device.restype = C.c_int
    lib.ar_stream_open.argtypes = [eng, C.c_int32, C.c_int32, C.c_int32]
    lib.ar_stream_open.restype = C.c_int
    lib.ar_stream_close.argtypes = [eng]
    lib.ar_stream_close.restype = None
    lib.ar_stream_submit.argtypes = [eng, C.c_int32, P(GamePod), C.c_int32, P(SearchCfg), P(C.c_uint64)]
    lib.ar_stream_submit.restype = C.c_int
    lib.ar_stream_collect.argtypes = [eng, C.c_int32, P(GameSummary), P(PositionRecord), C.c_int32, P(Stats)]
    lib.ar_stream_collect.restype = C.c_int
    lib.ar_stream_launch.argtypes = [eng, C.c_int32, P(SearchCfg)]
    lib.ar_stream_launch.restype = C.c_int
    lib.ar_stream_wait.argtypes = [eng, C.c_int32, P(Stats)]
    lib.ar_stream_wait.restype = C.c_int
    lib.ar_stream_times.argtypes = [eng, C.c_int32, P(C.c_double), P(C.c_double)]
    lib.ar_stream_times.restype = C.c_int
    lib.ar_stream_elapsed_ms.argtypes = [eng, C.c_int32, C.c_int32, P(C.c_double)]
    lib.ar_stream_elapsed_ms.restype = C.c_int
    if lib.ar_abi_version() != AR_ABI_VERSION:
        raise RuntimeError(f"ABI mismatch: library {lib.ar_abi_version()} != python {AR_ABI_VERSION}")
    _LIB = lib
    return lib
