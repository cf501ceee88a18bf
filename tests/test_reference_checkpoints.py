"""Checkpoints written in the reference's format load through `alpharat_b200.weights` with the tensor names the
CUDA evaluators ask for.  Runs the reference's own model classes (PyRatMLP, SymmetricMLP, PyRatCNN via
`CNNModelConfig.build_model`) in a subprocess with Hydra stubbed, so nothing of the reference leaks into the test
process; skipped where the reference tree is not mounted."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")

SCRIPT = r'''
import json, sys, types
ROOT, REF, OUT = sys.argv[1:4]
sys.path[:0] = [ROOT, ROOT + "/tests", REF]
for name in ("hydra", "hydra.core", "hydra.core.global_hydra", "omegaconf"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["hydra"].compose = lambda *a, **k: None
sys.modules["hydra"].initialize_config_dir = lambda *a, **k: None
sys.modules["hydra.core.global_hydra"].GlobalHydra = type("GlobalHydra", (), {})
sys.modules["omegaconf"].OmegaConf = type("OmegaConf", (), {})
sys.modules["omegaconf"].DictConfig = dict
import torch
from alpharat.nn.architectures.cnn.config import CNNModelConfig
from alpharat.nn.models.mlp import PyRatMLP
from alpharat.nn.models.symmetric import SymmetricMLP
from alpharat_b200 import _native as N
from alpharat_b200.weights import load_checkpoint_into
from nn_ref import make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict

w = h = 7
cnn_cfg = CNNModelConfig(trunk={"channels": 64, "blocks": [{"type": "res"}, {"type": "res"}, {"type": "gpool", "gpool_channels": 32}]})
cnn_cfg.set_data_dimensions(w, h)
cases = {
    "mlp": (PyRatMLP(obs_dim=7 * w * h + 6, hidden_dim=256), {"architecture": "mlp", "hidden_dim": 256},
            make_mlp_state_dict(0, 7 * w * h + 6), N.AR_ARCH_MLP),
    "symmetric": (SymmetricMLP(w, h, hidden_dim=256), {"architecture": "symmetric", "hidden_dim": 256},
                  make_symmetric_state_dict(2, w, h), N.AR_ARCH_SYMMETRIC),
    "cnn": (cnn_cfg.build_model(), cnn_cfg.model_dump(), make_cnn_state_dict(3, ("res", "res", "gpool")), N.AR_ARCH_CNN),
}

class Engine:
    def load_weights(self, arch, width, height, tensors):
        self.got = (arch, width, height, {k: tuple(v.shape) for k, v in tensors.items()},
                    {str(v.dtype) for v in tensors.values()})

report = {}
for name, (model, model_cfg, ours, arch_id) in cases.items():
    path = f"{OUT}/{name}.pt"
    # the keys nn/training/loop.py:395-421 writes and load_model_from_checkpoint reads (config/checkpoint.py:24-104)
    torch.save({"model_state_dict": model.state_dict(), "config": {"model": model_cfg}, "width": w, "height": h}, path)
    eng = Engine()
    load_checkpoint_into(eng, path)
    arch, gw, gh, shapes, dtypes = eng.got
    ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    report[name] = {
        "arch_ok": arch == arch_id and (gw, gh) == (w, h),
        "all_tensors_passed": {k: v for k, v in shapes.items() if not k.endswith("num_batches_tracked")}
                              == {k: v for k, v in ref_shapes.items() if not k.endswith("num_batches_tracked")},
        "f32_only": dtypes == {"float32"},
        # every tensor the CUDA loaders read (the generators of tests/nn_ref.py mirror them) exists with that shape
        "missing_or_misshaped": sorted(k for k, v in ours.items() if ref_shapes.get(k) != tuple(v.shape)),
        "extra_in_reference": sorted(k for k in ref_shapes if k not in ours and not k.endswith("num_batches_tracked")),
    }
print("REPORT " + json.dumps(report))
'''


@pytest.mark.skipif(not (REFERENCE / "alpharat" / "nn" / "models" / "mlp.py").exists(),
                    reason="the reference tree is only mounted in the build container")
def test_reference_format_checkpoints_load(tmp_path):
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(ROOT), str(REFERENCE), str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = next(ln for ln in r.stdout.splitlines() if ln.startswith("REPORT "))
    report = json.loads(line[len("REPORT "):])
    for name, rep in report.items():
        assert rep["arch_ok"] and rep["all_tensors_passed"] and rep["f32_only"], (name, rep)
        assert rep["missing_or_misshaped"] == [], (name, rep["missing_or_misshaped"])
        assert rep["extra_in_reference"] == [], (name, rep["extra_in_reference"])


@pytest.mark.skipif(not (REFERENCE / "alpharat" / "nn" / "models" / "mlp.py").exists(),
                    reason="the reference tree is only mounted in the build container")
def test_committed_goldens_are_what_the_reference_produces(tmp_path):
    """`tests/golden/make_golden.py` is run again (the reference's FlatObservationBuilder, PyRatMLP, SymmetricMLP,
    PyRatCNN and the Rust encoder fixtures) into a scratch directory; every committed golden file must hold the
    same arrays.  A golden file edited by hand, or generated from anything but the reference, fails here."""
    import numpy as np

    gold = ROOT / "tests" / "golden"
    r = subprocess.run([sys.executable, str(gold / "make_golden.py")], capture_output=True, text=True, timeout=900,
                       env={**__import__("os").environ, "GOLDEN_OUT": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-3000:]
    fresh = sorted(p.name for p in tmp_path.iterdir())
    assert fresh == sorted(p.name for p in gold.iterdir() if p.suffix in (".npz", ".json"))
    for name in fresh:
        if name.endswith(".json"):
            assert json.loads((tmp_path / name).read_text()) == json.loads((gold / name).read_text()), name
            continue
        a, b = np.load(tmp_path / name), np.load(gold / name)
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            assert a[k].shape == b[k].shape and np.allclose(a[k], b[k], rtol=0, atol=2e-6), (name, k)
