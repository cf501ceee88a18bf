"""TEST INFRASTRUCTURE: builds and loads tests/half_emul (the two-trees-per-warp device code compiled for the host)."""

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


def load_emul() -> C.CDLL:
    if not SO.exists() or any(d.stat().st_mtime > SO.stat().st_mtime for d in DEPS):
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                            "-Wno-unknown-pragmas", "-o", str(SO), str(SRC)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {SO.name} failed:\n{r.stderr[-4000:]}")
    lib = C.CDLL(str(SO))
    P = C.POINTER
    lib.half_emul_run.argtypes = [P(N.GamePod), C.c_int, P(N.SearchCfg), P(C.c_uint64), C.c_int, C.c_int,
                                  P(N.GameSummary), P(N.PositionRecord), C.c_int, P(N.SearchResultPod),
                                  P(C.c_ulonglong)]
    lib.half_emul_run.restype = C.c_int
    return lib


def emul_selfplay(lib, pods, cfg, seeds, pool_nodes=8192):
    n = len(pods)
    stride = max([p.max_turns for p in pods] + [1])
    summaries = (N.GameSummary * max(n, 1))()
    positions = (N.PositionRecord * max(n * stride, 1))()
    sd = (C.c_uint64 * max(n, 1))(*seeds)
    ctr = (C.c_ulonglong * 3)()
    rc = lib.half_emul_run(pods, n, C.byref(cfg), sd, pool_nodes, 0, summaries, positions, stride, None, ctr)
    assert rc == 0, f"emulation failed: {rc}"
    return summaries, positions, stride, list(ctr)


def emul_search(lib, pods, cfg, seeds, pool_nodes=8192):
    n = len(pods)
    out = (N.SearchResultPod * max(n, 1))()
    sd = (C.c_uint64 * max(n, 1))(*seeds)
    ctr = (C.c_ulonglong * 3)()
    rc = lib.half_emul_run(pods, n, C.byref(cfg), sd, pool_nodes, 1, None, None, 1, out, ctr)
    assert rc == 0, f"emulation failed: {rc}"
    return out
