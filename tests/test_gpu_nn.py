"""GPU leaf evaluator (on-device encode + tcgen05 MLP) against golden vectors of the real reference.

Tolerances (stated, bf16 operands with fp32 accumulation):
  vs reference torch fp32 (tests/golden/mlp_7x7.npz):  |d policy| <= 2e-2,  |d value| <= 3e-2 * max(1, |v|)
  vs the bf16-emulating restatement (tests/nn_ref.py):  |d policy| <= 2e-3,  |d value| <= 4e-3 * max(1, |v|)
Observation encoding is exact f32 work: <= 1e-6 like the reference's own parity test (parity.rs:16).
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine
from alpharat_b200.games import GameSpec, pods_array
from nn_ref import (cnn_forward, make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict, mlp_forward,
                    random_positions, symmetric_forward)

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_encode_matches_reference_builder():
    specs = random_positions(96, 7, 7, seed=123)
    gold = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    with Engine(concurrent_games=4, max_turns=120) as eng:
        obs = eng.encode(pods_array(specs))
    assert obs.shape == gold.shape == (96, 349)
    assert np.abs(obs - gold).max() <= 1e-6


def test_encode_matches_rust_fixtures(oracle):
    """The 7 JSON fixtures of crates/alpharat-sampling/tests/parity.rs (moves replayed by the oracle)."""
    fx = json.loads((GOLD / "encoder_fixtures.json").read_text())
    with Engine(concurrent_games=4, max_turns=120) as eng:
        for f in fx:
            spec = GameSpec(f["width"], f["height"], f["max_turns"], tuple(f["p1"]), tuple(f["p2"]),
                            [tuple(c) for c in f["cheese"]],
                            walls=[(tuple(a), tuple(b)) for a, b in f["walls"]],
                            mud=[(tuple(a), tuple(b), v) for a, b, v in f["mud"]])
            pods = pods_array([spec])
            for d1, d2 in f["moves"]:
                oracle.orc_game_make_move(pods, d1, d2)
            obs = eng.encode(pods)[0]
            exp = np.asarray(f["expected"], dtype=np.float32)
            assert obs.shape == exp.shape, f["name"]
            assert np.abs(obs - exp).max() <= 1e-6, f["name"]


@pytest.mark.parametrize("n", [1, 96, 128, 129, 1000])
def test_mlp_forward_matches_reference(n):
    base = random_positions(96, 7, 7, seed=123)
    specs = [base[i % 96] for i in range(n)]
    sd = make_mlp_state_dict(0, 349)
    gold = np.load(GOLD / "mlp_7x7.npz")
    obs = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_MLP, 7, 7, sd)
        p1, p2, v1, v2 = eng.nn_forward(pods_array(specs))
    idx = np.arange(n) % 96
    # reference torch fp32
    assert np.abs(p1 - gold["policy_p1"][idx]).max() <= 2e-2
    assert np.abs(p2 - gold["policy_p2"][idx]).max() <= 2e-2
    for v, g in ((v1, gold["value_p1"][idx]), (v2, gold["value_p2"][idx])):
        assert (np.abs(v - g) <= 3e-2 * np.maximum(1.0, np.abs(g))).all()
    # bf16-emulating restatement: isolates layout bugs from precision
    e1, e2, ev1, ev2 = mlp_forward(sd, obs[idx], emulate_bf16=True)
    assert np.abs(p1 - e1).max() <= 2e-3
    assert np.abs(p2 - e2).max() <= 2e-3
    for v, g in ((v1, ev1), (v2, ev2)):
        assert (np.abs(v - g) <= 4e-3 * np.maximum(1.0, np.abs(g))).all()
    assert np.allclose(p1.sum(1), 1, atol=1e-5) and np.allclose(p2.sum(1), 1, atol=1e-5)
    assert (v1 >= 0).all() and (v2 >= 0).all()


@pytest.mark.parametrize("hidden", [128, 64, 200])
def test_mlp_narrower_trunks_run_zero_padded(hidden):
    """hidden_dim < 256: the loader pads the trunk to the kernel's 256 columns with zero
    weights and biases, which leaves the outputs unchanged.  128: against the real reference's PyRatMLP(hidden_dim=128)
    golden; the other widths against the fp32 / bf16 restatements (validated against the reference at 128 and 256)."""
    n = 300
    base = random_positions(96, 7, 7, seed=123)
    specs = [base[i % 96] for i in range(n)]
    idx = np.arange(n) % 96
    sd = make_mlp_state_dict(1, 349, hidden=hidden)
    obs = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_MLP, 7, 7, sd)
        out = eng.nn_forward(pods_array(specs))
    if hidden == 128:
        gold = np.load(GOLD / "mlp_7x7_h128.npz")
        ref = (gold["policy_p1"][idx], gold["policy_p2"][idx], gold["value_p1"][idx], gold["value_p2"][idx])
        _check_vs_fp32(out, ref, 2e-2, 3e-2, "mlp hidden 128 vs reference golden")
    _check_vs_fp32(out, mlp_forward(sd, obs[idx]), 2e-2, 3e-2, f"mlp hidden {hidden} fp32")
    _check_vs_fp32(out, mlp_forward(sd, obs[idx], emulate_bf16=True), 2e-3, 4e-3, f"mlp hidden {hidden} bf16 restatement")


def test_mlp_wider_than_256_fails_loudly():
    sd = make_mlp_state_dict(1, 349, hidden=512)
    with Engine(concurrent_games=4, max_turns=120) as eng:
        with pytest.raises(RuntimeError, match="hidden_dim"):
            eng.load_weights(N.AR_ARCH_MLP, 7, 7, sd)


@pytest.mark.parametrize("w,h,n", [(5, 5, 300), (4, 3, 70), (6, 5, 129)])
def test_mlp_forward_small_boards(w, h, n):
    """Boards whose observation needs fewer K-blocks than the 256-wide hidden layers (5x5: 3, 4x3: 2): the operand
    buffer is sized by the larger of the two.  Checked against the fp32 restatement of PyRatMLP (validated against
    the real reference at 7x7 in test_oracle_golden.py) on the device encoder's observations."""
    base = random_positions(40, w, h, seed=11)
    specs = [base[i % 40] for i in range(n)]
    sd = make_mlp_state_dict(5, 7 * w * h + 6)
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_MLP, w, h, sd)
        pods = pods_array(specs)
        obs = eng.encode(pods)
        out = eng.nn_forward(pods)
    _check_vs_fp32(out, mlp_forward(sd, obs), 2e-2, 3e-2, f"mlp {w}x{h} fp32")
    _check_vs_fp32(out, mlp_forward(sd, obs, emulate_bf16=True), 2e-3, 4e-3, f"mlp {w}x{h} bf16 restatement")


def _check_vs_fp32(out, ref, tol_p, tol_v, tag):
    p1, p2, v1, v2 = out
    assert np.abs(p1 - ref[0]).max() <= tol_p, (tag, float(np.abs(p1 - ref[0]).max()))
    assert np.abs(p2 - ref[1]).max() <= tol_p, (tag, float(np.abs(p2 - ref[1]).max()))
    for v, g in ((v1, ref[2]), (v2, ref[3])):
        err = np.abs(v - g) / np.maximum(1.0, np.abs(g))
        assert err.max() <= tol_v, (tag, float(err.max()))
    assert np.allclose(p1.sum(1), 1, atol=1e-5) and np.allclose(p2.sum(1), 1, atol=1e-5)
    assert (v1 >= 0).all() and (v2 >= 0).all()


def _positions(w, h, n):
    base = random_positions(96, 7, 7, seed=123) if (w, h) == (7, 7) else random_positions(40, w, h, seed=321)
    return [base[i % len(base)] for i in range(n)], len(base)


@pytest.mark.parametrize("w,h,n", [(7, 7, 1), (7, 7, 96), (7, 7, 64), (7, 7, 65), (7, 7, 1000), (5, 5, 40), (5, 5, 333)])
def test_symmetric_forward_matches_reference(w, h, n):
    """SymmetricMLP on tcgen05 vs the real reference's fp32 outputs (goldens) — stated tolerance for bf16
    operands: |d policy| <= 2.5e-2, |d value| <= 4e-2 * max(1, |v|)."""
    specs, nb = _positions(w, h, n)
    sd = make_symmetric_state_dict(2, w, h)
    g = np.load(GOLD / f"symmetric_{w}x{h}.npz")
    idx = np.arange(n) % nb
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_SYMMETRIC, w, h, sd)
        out = eng.nn_forward(pods_array(specs))
    _check_vs_fp32(out, [g[k][idx] for k in ("policy_p1", "policy_p2", "value_p1", "value_p2")], 2.5e-2, 4e-2,
                   f"symmetric {w}x{h} n={n}")


@pytest.mark.parametrize("hidden", [128, 96])
def test_symmetric_narrower_models_run_zero_padded(hidden):
    """hidden_dim < 256: the loader pads every stage to the kernel's 256 columns (cat(shared, p_i) keeps the kernel's
    stride: columns [0, H) and [256, 256 + H)).  128: against the real reference's SymmetricMLP(hidden_dim=128) golden;
    96: against the fp32 restatement (validated against the reference at 128 and 256)."""
    n = 200
    specs, nb = _positions(7, 7, n)
    idx = np.arange(n) % nb
    sd = make_symmetric_state_dict(4, 7, 7, hidden=hidden)
    obs = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_SYMMETRIC, 7, 7, sd)
        out = eng.nn_forward(pods_array(specs))
    if hidden == 128:
        g = np.load(GOLD / "symmetric_7x7_h128.npz")
        _check_vs_fp32(out, [g[k][idx] for k in ("policy_p1", "policy_p2", "value_p1", "value_p2")], 2.5e-2, 4e-2,
                       "symmetric hidden 128 vs reference golden")
    _check_vs_fp32(out, symmetric_forward(sd, obs[idx], 7, 7), 2.5e-2, 4e-2, f"symmetric hidden {hidden} fp32")


def test_symmetric_swaps_outputs_when_players_swap():
    """Structural P1/P2 symmetry (symmetric.py:20-24): the two-rows-per-position mapping must keep it exactly."""
    specs, _ = _positions(7, 7, 96)
    swapped = [GameSpec(s.width, s.height, s.max_turns, s.p2, s.p1, s.cheese, walls=s.walls, mud=s.mud, turn=s.turn,
                        p1_score=s.p2_score, p2_score=s.p1_score, p1_mud=s.p2_mud, p2_mud=s.p1_mud) for s in specs]
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_SYMMETRIC, 7, 7, make_symmetric_state_dict(2, 7, 7))
        a = eng.nn_forward(pods_array(specs))
        b = eng.nn_forward(pods_array(swapped))
    assert np.array_equal(a[0], b[1]) and np.array_equal(a[1], b[0])
    assert np.array_equal(a[2], b[3]) and np.array_equal(a[3], b[2])


@pytest.mark.parametrize("tag,blocks", [("gpool", ("res", "res", "gpool")), ("res", ("res",))])
@pytest.mark.parametrize("w,h,n", [(7, 7, 1), (7, 7, 96), (7, 7, 3), (7, 7, 601), (5, 5, 40), (5, 5, 333)])
def test_cnn_forward_matches_reference(tag, blocks, w, h, n):
    """PyRatCNN (configs/model/cnn_gpool.yaml and cnn.yaml trunks) as an implicit GEMM on tcgen05 vs the real
    reference's fp32 outputs — stated tolerance for bf16 operands: |d policy| <= 3e-2,
    |d value| <= 5e-2 * max(1, |v|)."""
    specs, nb = _positions(w, h, n)
    sd = make_cnn_state_dict(3, blocks)
    g = np.load(GOLD / f"cnn_{tag}_{w}x{h}.npz")
    idx = np.arange(n) % nb
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_CNN, w, h, sd)
        out = eng.nn_forward(pods_array(specs))
    _check_vs_fp32(out, [g[k][idx] for k in ("policy_p1", "policy_p2", "value_p1", "value_p2")], 3e-2, 5e-2,
                   f"cnn {tag} {w}x{h} n={n}")


@pytest.mark.parametrize("channels,gpool", [(32, 16), (48, 32), (16, 16)])
def test_cnn_narrower_trunks_run_zero_padded(channels, gpool):
    """Trunks under 64 channels: the loader widens every channel axis with zeros (BatchNorm scale and shift 0), the
    extra channels stay exactly 0 through the residual stream.  32 channels / gpool 16: against the real reference's
    golden; the other widths against the fp32 restatement (validated against the reference at 32 and 64 channels)."""
    n = 150
    blocks = ("res", "gpool", "res")
    specs, nb = _positions(7, 7, n)
    idx = np.arange(n) % nb
    sd = make_cnn_state_dict(6, blocks, channels=channels, gpool_channels=gpool)
    obs = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    with Engine(concurrent_games=4, max_turns=120) as eng:
        eng.load_weights(N.AR_ARCH_CNN, 7, 7, sd)
        out = eng.nn_forward(pods_array(specs))
    if channels == 32:
        g = np.load(GOLD / "cnn_gpool_7x7_c32.npz")
        _check_vs_fp32(out, [g[k][idx] for k in ("policy_p1", "policy_p2", "value_p1", "value_p2")], 3e-2, 5e-2,
                       "cnn 32 channels vs reference golden")
    _check_vs_fp32(out, cnn_forward(sd, obs[idx], 7, 7), 3e-2, 5e-2, f"cnn {channels} channels fp32")


def test_cnn_wider_than_64_channels_fails_loudly():
    sd = make_cnn_state_dict(6, ("res",), channels=128)
    with Engine(concurrent_games=4, max_turns=120) as eng:
        with pytest.raises(RuntimeError, match="64"):
            eng.load_weights(N.AR_ARCH_CNN, 7, 7, sd)


@pytest.mark.parametrize("w,h,n", [(8, 8, 37), (6, 4, 100), (4, 4, 90), (4, 7, 61)])
def test_cnn_and_symmetric_forward_other_boards(w, h, n):
    """Board geometries the golden files do not cover: one position per tile (8x8), non-square boards, five
    positions per tile (4x4).  Reference = the numpy restatements of PyRatCNN / SymmetricMLP (validated against
    the real reference at 5x5 and 7x7 in test_oracle_golden.py) on the device encoder's observations."""
    base = random_positions(40, w, h, seed=21)
    specs = [base[i % 40] for i in range(n)]
    pods = pods_array(specs)
    with Engine(concurrent_games=4, max_turns=120) as eng:
        obs = eng.encode(pods)
        for blocks in (("res", "res", "gpool"), ("gpool", "res")):
            sd = make_cnn_state_dict(7, blocks)
            eng.load_weights(N.AR_ARCH_CNN, w, h, sd)
            _check_vs_fp32(eng.nn_forward(pods), cnn_forward(sd, obs, w, h), 3e-2, 5e-2, f"cnn {blocks} {w}x{h}")
        if 5 * w * h + 1 <= 256:
            sd = make_symmetric_state_dict(8, w, h)
            eng.load_weights(N.AR_ARCH_SYMMETRIC, w, h, sd)
            _check_vs_fp32(eng.nn_forward(pods), symmetric_forward(sd, obs, w, h), 3e-2, 5e-2, f"symmetric {w}x{h}")


def test_unsupported_evaluator_shapes_fail_loudly():
    sd = make_symmetric_state_dict(2, 7, 7, hidden=512)
    with Engine(concurrent_games=4, max_turns=120) as eng:
        with pytest.raises(RuntimeError, match="hidden_dim"):
            eng.load_weights(N.AR_ARCH_SYMMETRIC, 7, 7, sd)
        with pytest.raises((RuntimeError, ValueError), match="stem.weight"):
            eng.load_weights(N.AR_ARCH_CNN, 7, 7, sd)


def test_nn_forward_without_weights_fails_loudly():
    specs = random_positions(2, 7, 7, seed=1)
    with Engine(concurrent_games=4, max_turns=120) as eng:
        with pytest.raises(RuntimeError, match="no evaluator"):
            eng.nn_forward(pods_array(specs))
