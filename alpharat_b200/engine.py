"""Thin object wrapper over the C-ABI engine handle (include/alpharat_cuda.h)."""

from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _native as N

_STATUS_EXC = {
    N.AR_ERR_INVALID_ARG: ValueError,
    N.AR_ERR_CUDA: RuntimeError,
    N.AR_ERR_POOL_OVERFLOW: RuntimeError,
    N.AR_ERR_NONFINITE: RuntimeError,
    N.AR_ERR_UNSUPPORTED: NotImplementedError,
    N.AR_ERR_NO_WEIGHTS: RuntimeError,
}


def search_cfg(*, simulations: int, batch_size: int = 8, c_puct: float = 1.5, fpu_reduction: float = 0.2,
               force_k: float = 2.0, noise_epsilon: float = 0.0, noise_concentration: float = 10.83,
               collision_limit_min: int = 1, collision_limit_max: int = 256,
               collision_scaling_start: int = 800, collision_scaling_end: int = 50_000,
               collision_scaling_power: float = 1.0) -> N.SearchCfg:
    return N.SearchCfg(simulations, batch_size, c_puct, fpu_reduction, force_k, noise_epsilon,
                       noise_concentration, collision_limit_min, collision_limit_max,
                       collision_scaling_start, collision_scaling_end, collision_scaling_power)


class Engine:
    """One engine per GPU; single-threaded handle (see header)."""

    def __init__(self, *, device: int = 0, concurrent_games: int = 4096, pool_nodes: int = 0,
                 max_cells: int = 64, max_turns: int = 64, max_batch_size: int = 16,
                 max_simulations: int = 2048, tree_engine: str | int = "warp") -> None:
        self._lib = N.load_library()
        te = {"warp": N.AR_TREE_WARP, "thread": N.AR_TREE_THREAD, "half": N.AR_TREE_HALF}.get(tree_engine, tree_engine)
        cfg = N.EngineCfg(N.AR_ABI_VERSION, device, concurrent_games, pool_nodes, max_cells,
                          max_turns, max_batch_size, max_simulations, int(te))
        handle = C.c_void_p()
        st = self._lib.ar_engine_create(C.byref(cfg), C.byref(handle))
        if st != N.AR_OK:
            msg = self._lib.ar_last_error(None)
            raise _STATUS_EXC.get(st, RuntimeError)(f"ar_engine_create failed ({st}): {msg.decode() if msg else ''}")
        self._h = handle
        self.cfg = cfg
        self.has_evaluator = False

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.ar_engine_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "Engine":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def _check(self, st: int, what: str) -> None:
        if st != N.AR_OK:
            msg = self._lib.ar_last_error(self._h)
            raise _STATUS_EXC.get(st, RuntimeError)(f"{what} failed ({st}): {msg.decode() if msg else ''}")

    # --- evaluator -----------------------------------------------------------------------
    def load_weights(self, arch: int, width: int, height: int, tensors: dict[str, np.ndarray]) -> None:
        keep = []
        descs = (N.TensorDesc * max(len(tensors), 1))()
        for i, (name, arr) in enumerate(tensors.items()):
            a = np.ascontiguousarray(arr, dtype=np.float32)
            keep.append(a)
            descs[i].name = name.encode()
            descs[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            descs[i].ndim = a.ndim
            for d in range(4):
                descs[i].shape[d] = a.shape[d] if d < a.ndim else 1
        self._check(self._lib.ar_engine_load_weights(self._h, arch, width, height, descs, len(tensors)),
                    "ar_engine_load_weights")
        self.has_evaluator = arch != N.AR_ARCH_UNIFORM

    def set_eval_cache(self, entries_per_tree: int) -> None:
        """Evaluation cache of the NN-guided mode (`cache_size` of rust_self_play; CachedBackend,
        cached_backend.rs:54-120): positions per resident tree, 0 disables."""
        self._check(self._lib.ar_engine_set_eval_cache(self._h, int(entries_per_tree)), "ar_engine_set_eval_cache")

    # --- search ----------------------------------------------------------------------------
    def search_batch(self, pods, cfg: N.SearchCfg, seeds: Sequence[int]):
        n = len(pods)
        out = (N.SearchResultPod * max(n, 1))()
        sd = (C.c_uint64 * max(n, 1))(*[int(s) & ((1 << 64) - 1) for s in seeds])
        self._check(self._lib.ar_search_batch(self._h, pods, n, C.byref(cfg), sd, out), "ar_search_batch")
        return out

    @staticmethod
    def _record_array(count: int):
        """`PositionRecord[count]` over lazily mapped (untouched, unzeroed) memory: only the rows a game
        really played are ever written, so the pages of the unused tail of each stride are never faulted
        in (a zero-filled ctypes array of the bench's 65536 x 50 records costs 0.6 s on its own)."""
        if count * C.sizeof(N.PositionRecord) < (1 << 20):
            return (N.PositionRecord * count)()
        buf = np.empty(count * C.sizeof(N.PositionRecord), dtype=np.uint8)
        return (N.PositionRecord * count).from_buffer(buf)

    # --- self-play ---------------------------------------------------------------------------
    def selfplay(self, pods, cfg: N.SearchCfg, seeds: Sequence[int], *, stride: int | None = None,
                 progress: N.Progress | None = None):
        n = len(pods)
        if stride is None:
            stride = max([p.max_turns for p in pods] + [1])
        summaries = (N.GameSummary * max(n, 1))()
        positions = self._record_array(max(n * stride, 1))
        stats = N.Stats()
        sd = (C.c_uint64 * max(n, 1))(*[int(s) & ((1 << 64) - 1) for s in seeds])
        pr = C.byref(progress) if progress is not None else None
        self._check(self._lib.ar_selfplay_run(self._h, pods, n, C.byref(cfg), sd, summaries, positions,
                                              stride, pr, C.byref(stats)), "ar_selfplay_run")
        return summaries, positions, stride, stats

    def selfplay_upload(self, pods, seeds: Sequence[int]) -> None:
        n = len(pods)
        sd = (C.c_uint64 * max(n, 1))(*[int(s) & ((1 << 64) - 1) for s in seeds])
        self._check(self._lib.ar_selfplay_upload(self._h, pods, n, sd), "ar_selfplay_upload")

    def selfplay_run_resident(self, cfg: N.SearchCfg) -> N.Stats:
        stats = N.Stats()
        self._check(self._lib.ar_selfplay_run_resident(self._h, C.byref(cfg), C.byref(stats)),
                    "ar_selfplay_run_resident")
        return stats

    def selfplay_download(self, n: int, stride: int):
        summaries = (N.GameSummary * max(n, 1))()
        positions = self._record_array(max(n * stride, 1))
        self._check(self._lib.ar_selfplay_download(self._h, summaries, positions, stride), "ar_selfplay_download")
        return summaries, positions

    def selfplay_pack_device(self, n: int):
        """Packed records of the resident batch as device memory: `(summaries_host, d_summaries, d_records,
        n_records)`; the device pointers are ints (see `parallel.device_bytes`)."""
        summaries = (N.GameSummary * max(n, 1))()
        ds, dr, nr = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        self._check(self._lib.ar_selfplay_pack_device(self._h, summaries, C.byref(ds), C.byref(dr), C.byref(nr)),
                    "ar_selfplay_pack_device")
        return summaries, ds.value or 0, dr.value or 0, int(nr.value)

    # --- streaming self-play (continuous game feed, ar_stream_*) -------------------------------
    def stream_open(self, n_buffers: int, max_games: int, stride: int) -> None:
        self._check(self._lib.ar_stream_open(self._h, n_buffers, max_games, stride), "ar_stream_open")

    def stream_close(self) -> None:
        self._lib.ar_stream_close(self._h)

    def stream_submit(self, buffer: int, pods, cfg: N.SearchCfg | None, seeds) -> None:
        """Upload a batch into `buffer` and launch it; `cfg=None` uploads only (`stream_launch` plays it).
        `seeds` may be a ready `c_uint64` array."""
        n = len(pods)
        sd = seeds if isinstance(seeds, C.Array) else (C.c_uint64 * max(n, 1))(*[int(s) & ((1 << 64) - 1) for s in seeds])
        self._check(self._lib.ar_stream_submit(self._h, buffer, pods, n, C.byref(cfg) if cfg is not None else None, sd),
                    "ar_stream_submit")

    def stream_collect(self, buffer: int, n: int, stride: int):
        summaries = (N.GameSummary * max(n, 1))()
        positions = self._record_array(max(n * stride, 1))
        stats = N.Stats()
        self._check(self._lib.ar_stream_collect(self._h, buffer, summaries, positions, stride, C.byref(stats)),
                    "ar_stream_collect")
        return summaries, positions, stride, stats

    def stream_launch(self, buffer: int, cfg: N.SearchCfg) -> None:
        self._check(self._lib.ar_stream_launch(self._h, buffer, C.byref(cfg)), "ar_stream_launch")

    def stream_wait(self, buffer: int) -> N.Stats:
        stats = N.Stats()
        self._check(self._lib.ar_stream_wait(self._h, buffer, C.byref(stats)), "ar_stream_wait")
        return stats

    def stream_times(self, buffer: int) -> tuple[float, float]:
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._check(self._lib.ar_stream_times(self._h, buffer, C.byref(a), C.byref(b)), "ar_stream_times")
        return a.value, b.value

    def stream_elapsed_ms(self, first: int, last: int) -> float:
        ms = C.c_double(0.0)
        self._check(self._lib.ar_stream_elapsed_ms(self._h, first, last, C.byref(ms)), "ar_stream_elapsed_ms")
        return ms.value

    # --- evaluator entry points ------------------------------------------------------------
    def encode(self, pods) -> np.ndarray:
        n = len(pods)
        dim = 7 * pods[0].width * pods[0].height + 6 if n else 0
        out = np.zeros((n, dim), dtype=np.float32)
        self._check(self._lib.ar_encode_observations(self._h, pods, n, out.ctypes.data_as(C.POINTER(C.c_float))),
                    "ar_encode_observations")
        return out

    def nn_forward(self, pods):
        n = len(pods)
        p1 = np.zeros((n, 5), np.float32)
        p2 = np.zeros((n, 5), np.float32)
        v1 = np.zeros(n, np.float32)
        v2 = np.zeros(n, np.float32)
        f = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        self._check(self._lib.ar_nn_forward(self._h, pods, n, f(p1), f(p2), f(v1), f(v2)), "ar_nn_forward")
        return p1, p2, v1, v2
