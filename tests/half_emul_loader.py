"""TEST INFRASTRUCTURE: builds and loads tests/half_emul — the device code of the two warp-resident tree engines
(two trees per warp: half_emul.cpp / mcts_half.cuh; one tree per warp: warp_emul.cpp / mcts_device.cuh) compiled for the
host behind a shim of the warp intrinsics."""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

from alpharat_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "half_emul" / "half_emul.cpp"
SO = ROOT / "tests" / "half_emul" / "libhalf_emul.so"
DEPS = [SRC, ROOT / "tests" / "half_emul" / "simt_shim.h", ROOT / "alpharat_b200" / "csrc" / "mcts_half.cuh",
        ROOT / "alpharat_b200" / "csrc" / "mcts_device.cuh", ROOT / "alpharat_b200" / "csrc" / "host_tables.hpp",
        ROOT / "include" / "alpharat_cuda.h"]


class _Emul:
    """One emulated engine: `run(pods, n, cfg, seeds, pool_nodes, search_only, summaries, positions, stride, out, ctr)`."""

    def __init__(self, run):
        self.run = run


def _build(src: Path, so: Path, entry: str) -> _Emul:
    deps = DEPS + [src]
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                            "-Wno-unknown-pragmas", "-o", str(so), str(src)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {so.name} failed:\n{r.stderr[-4000:]}")
    lib = C.CDLL(str(so))
    P = C.POINTER
    fn = getattr(lib, entry)
    fn.argtypes = [P(N.GamePod), C.c_int, P(N.SearchCfg), P(C.c_uint64), C.c_int, C.c_int,
                   P(N.GameSummary), P(N.PositionRecord), C.c_int, P(N.SearchResultPod), P(C.c_ulonglong)]
    fn.restype = C.c_int
    e = _Emul(fn)
    e._lib = lib
    return e


def load_emul() -> _Emul:
    """Two trees per warp (mcts_half.cuh): 16 fibres per half."""
    return _build(SRC, SO, "half_emul_run")


def load_warp_emul() -> _Emul:
    """One tree per warp (mcts_device.cuh): 32 fibres."""
    return _build(ROOT / "tests" / "half_emul" / "warp_emul.cpp", ROOT / "tests" / "half_emul" / "libwarp_emul.so",
                  "warp_emul_run")


def emul_selfplay(lib, pods, cfg, seeds, pool_nodes=8192):
    n = len(pods)
    stride = max([p.max_turns for p in pods] + [1])
    summaries = (N.GameSummary * max(n, 1))()
    positions = (N.PositionRecord * max(n * stride, 1))()
    sd = (C.c_uint64 * max(n, 1))(*seeds)
    ctr = (C.c_ulonglong * 3)()
    rc = lib.run(pods, n, C.byref(cfg), sd, pool_nodes, 0, summaries, positions, stride, None, ctr)
    assert rc == 0, f"emulation failed: {rc}"
    return summaries, positions, stride, list(ctr)


def emul_search(lib, pods, cfg, seeds, pool_nodes=8192):
    n = len(pods)
    out = (N.SearchResultPod * max(n, 1))()
    sd = (C.c_uint64 * max(n, 1))(*seeds)
    ctr = (C.c_ulonglong * 3)()
    rc = lib.run(pods, n, C.byref(cfg), sd, pool_nodes, 1, None, None, 1, out, ctr)
    assert rc == 0, f"emulation failed: {rc}"
    return out
