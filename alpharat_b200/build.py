"""Build libalpharat_cuda.so in-tree with nvcc for sm_100a.

    python -m alpharat_b200.build [--force]

The tree kernels (engine.cu) are compiled with -fmad=false so that f32 arithmetic follows the
reference operation-for-operation; the leaf-evaluator kernels (nn_kernels.cu) allow FMA.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libalpharat_cuda.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

UNITS = [
    ("engine.cu", ["-fmad=false"]),
    ("nn_kernels.cu", []),
    ("nn_symmetric.cu", []),
    ("nn_cnn.cu", []),
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: cannot build libalpharat_cuda.so")
    return nvcc


def _stale(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    headers = list(CSRC.glob("*.cuh")) + list((HERE.parent / "include").glob("*.h"))
    objs = []
    logs = []
    for name, extra in UNITS:
        src = CSRC / name
        if not src.exists():
            continue
        obj = CSRC / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", str(src), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    if force or _stale(OUT, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", str(OUT), *map(str, objs), "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(logs))
    (HERE / "build_ptxas.log").write_text("\n".join(logs)) if logs else None
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
