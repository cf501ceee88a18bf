"""Shared test plumbing.

`oracle` (CPU restatement, tests only) is built on demand with oracle/Makefile and loaded via
ctypes with the same POD layouts as include/alpharat_cuda.h.  GPU tests are marked
`@pytest.mark.gpu`; everything else runs on a CPU-only box.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from alpharat_b200 import _native as N  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


EVAL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(N.GamePod), C.c_int, C.POINTER(C.c_float),
                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float))


def build_oracle() -> Path:
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return ROOT / "oracle" / "liboracle.so"


def load_oracle() -> C.CDLL:
    so = ROOT / "oracle" / "liboracle.so"
    srcs = list((ROOT / "oracle").glob("*.cpp")) + list((ROOT / "oracle").glob("*.hpp"))
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        try:
            build_oracle()
        except Exception:
            if not so.exists():
                raise
    lib = C.CDLL(str(so))
    P = C.POINTER
    u64x4 = P(C.c_uint64)
    lib.orc_rng_seed.argtypes = [C.c_uint64, u64x4]
    lib.orc_rng_next_u64.argtypes = [u64x4]
    lib.orc_rng_next_u64.restype = C.c_uint64
    lib.orc_rng_gen_range.argtypes = [u64x4, C.c_uint32]
    lib.orc_rng_gen_range.restype = C.c_uint32
    lib.orc_rng_uniform_f32.argtypes = [u64x4, C.c_float, C.c_float]
    lib.orc_rng_uniform_f32.restype = C.c_float
    lib.orc_sample_action.argtypes = [u64x4, P(C.c_float)]
    lib.orc_sample_action.restype = C.c_uint32
    lib.orc_collisions_left.argtypes = [C.c_uint32, P(N.SearchCfg)]
    lib.orc_collisions_left.restype = C.c_uint32
    lib.orc_game_make_move.argtypes = [P(N.GamePod), C.c_uint8, C.c_uint8]
    lib.orc_game_make_unmake_roundtrip.argtypes = [P(N.GamePod), C.c_uint8, C.c_uint8]
    lib.orc_game_make_unmake_roundtrip.restype = C.c_int
    lib.orc_game_effective_actions.argtypes = [P(N.GamePod), C.c_int, P(C.c_uint8)]
    lib.orc_game_over.argtypes = [P(N.GamePod)]
    lib.orc_game_over.restype = C.c_int
    lib.orc_obs_dim.argtypes = [C.c_int, C.c_int]
    lib.orc_obs_dim.restype = C.c_int
    lib.orc_encode.argtypes = [P(N.GamePod), C.c_int, P(C.c_float)]
    lib.orc_search.argtypes = [P(N.GamePod), P(N.SearchCfg), C.c_uint64, C.c_void_p, C.c_void_p,
                               P(C.c_float), P(N.SearchResultPod), P(C.c_int)]
    lib.orc_search.restype = C.c_int
    lib.orc_selfplay.argtypes = [P(N.GamePod), C.c_int, P(N.SearchCfg), P(C.c_uint64), C.c_int,
                                 C.c_void_p, C.c_void_p, P(N.GameSummary), P(N.PositionRecord),
                                 C.c_int, P(N.Stats)]
    lib.orc_selfplay.restype = C.c_int
    return lib


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()


def oracle_selfplay(lib, pods, cfg, seeds, n_threads=0, eval_cb=None):
    n = len(pods)
    stride = max([p.max_turns for p in pods] + [1])
    summaries = (N.GameSummary * max(n, 1))()
    positions = (N.PositionRecord * max(n * stride, 1))()
    stats = N.Stats()
    sd = (C.c_uint64 * max(n, 1))(*seeds)
    nt = n_threads or (os.cpu_count() or 1)
    cb = C.cast(eval_cb, C.c_void_p) if eval_cb is not None else None
    rc = lib.orc_selfplay(pods, n, C.byref(cfg), sd, nt, cb, None, summaries, positions, stride, C.byref(stats))
    assert rc == 0, f"oracle self-play failed: {rc}"
    return summaries, positions, stride, stats


def oracle_search(lib, pod, cfg, seed, eval_cb=None, const_values=None):
    out = N.SearchResultPod()
    clean = C.c_int(0)
    cv = (C.c_float * 2)(*const_values) if const_values is not None else None
    cb = C.cast(eval_cb, C.c_void_p) if eval_cb is not None else None
    rc = lib.orc_search(C.byref(pod), C.byref(cfg), seed, cb, None, cv, C.byref(out), C.byref(clean))
    return rc, out, bool(clean.value)


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
