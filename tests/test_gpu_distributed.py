"""`cuda_self_play_distributed` on real GPUs: two ranks under torchrun play one seeded run of
classic mazes (walls, mud, random starts — the generator keywords must reach every rank), rank 0 receives every
rank's packed records over NCCL, and the result equals the 1-GPU run game for game."""

from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from alpharat_b200 import _native as N
from conftest import has_gpu

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _n_gpus() -> int:
    import torch

    return torch.cuda.device_count() if has_gpu() else 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs on one node")
def test_two_gpu_run_equals_one_gpu_run(tmp_path):
    from alpharat_b200.selfplay import cuda_self_play

    n, seed = 26, 1234
    out = tmp_path / "gathered.npz"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(ROOT / "tests" / "dist_worker.py"),
                        str(out), str(n), str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    z = np.load(out)
    kw = dict(width=7, height=5, cheese_count=6, max_turns=24, maze_type="classic", positions="random",
              simulations=200, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103, output_dir=None)
    stats, summ, pos, stride = cuda_self_play(num_games=n, seed=seed, concurrent_games=16, return_records=True, **kw)
    assert int(z["total_games"]) == n == stats.total_games
    assert int(z["total_positions"]) == stats.total_positions
    assert int(z["total_simulations"]) == stats.total_simulations
    ssz, rsz = C.sizeof(N.GameSummary), C.sizeof(N.PositionRecord)
    gs = (N.GameSummary * n).from_buffer_copy(z["summaries"].tobytes())
    assert len(z["summaries"]) == n * ssz and len(z["records"]) == stats.total_positions * rsz
    rec = z["records"].tobytes()
    off = 0
    for g in range(n):
        a, b = gs[g], summ[g]
        # the shard's game indices are local to the rank; everything else is the 1-GPU game
        assert (a.n_positions, a.final_p1_score, a.final_p2_score, a.result, a.total_simulations) == (
            b.n_positions, b.final_p1_score, b.final_p2_score, b.result, b.total_simulations), g
        assert bytes(a.cheese_outcomes) == bytes(b.cheese_outcomes), g
        for t in range(a.n_positions):
            want = bytes(memoryview(pos[g * stride + t]).cast("B")) if False else C.string_at(
                C.addressof(pos[g * stride + t]), rsz)
            assert rec[off:off + rsz] == want, (g, t)
            off += rsz
    assert off == len(rec)
