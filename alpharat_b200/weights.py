"""Checkpoint -> engine weights (the only place PyTorch is used, per the north star).

Reads what `load_model_from_checkpoint` reads (alpharat/config/checkpoint.py:24-104): keys
`model_state_dict`, `config.model`, `width`, `height` (written at nn/training/loop.py:395-421).
The state_dict goes to the engine as named f32 tensors; BatchNorm folding, bf16 conversion and
packing happen inside the library.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _native as N

_ARCH = {"mlp": N.AR_ARCH_MLP, "symmetric": N.AR_ARCH_SYMMETRIC, "cnn": N.AR_ARCH_CNN}


def state_dict_to_numpy(sd: dict[str, Any]) -> dict[str, np.ndarray]:
    out = {}
    for k, v in sd.items():
        a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        if a.dtype.kind in "iub":
            a = a.astype(np.float32)
        out[k] = np.ascontiguousarray(a, dtype=np.float32)
    return out


def load_state_dict_into(engine, arch: str, width: int, height: int, sd: dict[str, Any]) -> None:
    if arch not in _ARCH:
        raise NotImplementedError(f"architecture {arch!r} has no CUDA evaluator")
    engine.load_weights(_ARCH[arch], width, height, state_dict_to_numpy(sd))


def load_checkpoint_into(engine, checkpoint_path: str) -> None:
    import torch

    ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    cfg = ck.get("config", {})
    model_cfg = cfg.get("model", {}) if isinstance(cfg, dict) else {}
    arch = model_cfg.get("architecture") if isinstance(model_cfg, dict) else getattr(model_cfg, "architecture", None)
    if arch is None:  # the reference loader refuses too (alpharat/config/checkpoint.py:78-82)
        raise ValueError(f"checkpoint {checkpoint_path!r} has no config.model.architecture discriminator")
    load_state_dict_into(engine, arch, int(ck["width"]), int(ck["height"]), ck["model_state_dict"])
