"""The driver-facing contract of bench.py, on the arm that runs without a GPU: `--impl reference` prints exactly one
JSON line on stdout with the keys the CUDA arm prints (same metric / unit / config), `impl: reference`, a
`cpu_baseline` describing the run and an `e2e` that repeats the line's value with zero copies."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-seconds", "1.5"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "self-play MCTS simulations/sec" and d["unit"] == "simulations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"] and d["config"]["games_per_step_per_gpu"] == 131072
    # the CUDA arm builds its config with the same function
    src = (ROOT / "bench.py").read_text()
    assert src.count('"config": bench_config(args,') == 2


def test_cuda_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm fails loudly instead of timing something else."""
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
