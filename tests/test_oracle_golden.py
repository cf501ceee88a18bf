"""Oracle / restatements vs golden vectors produced by the REAL reference (tests/golden/make_golden.py)."""

from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from alpharat_b200 import _native as N
from alpharat_b200.engine import search_cfg
from alpharat_b200.games import GameSpec, make_games, pods_array
from conftest import oracle_search, oracle_selfplay
from nn_ref import (cnn_forward, make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict, mlp_forward,
                    random_positions, symmetric_forward)

GOLD = Path(__file__).resolve().parent / "golden"


def _encode(oracle, pods):
    n = len(pods)
    dim = oracle.orc_obs_dim(pods[0].width, pods[0].height)
    out = np.zeros((n, dim), np.float32)
    oracle.orc_encode(pods, n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def test_encoder_matches_rust_fixtures(oracle):
    """crates/alpharat-sampling/tests/parity.rs: 7 fixtures, tolerance 1e-6, moves replayed through the engine."""
    fx = json.loads((GOLD / "encoder_fixtures.json").read_text())
    assert len(fx) == 7
    for f in fx:
        spec = GameSpec(f["width"], f["height"], f["max_turns"], tuple(f["p1"]), tuple(f["p2"]),
                        [tuple(c) for c in f["cheese"]],
                        walls=[(tuple(a), tuple(b)) for a, b in f["walls"]],
                        mud=[(tuple(a), tuple(b), v) for a, b, v in f["mud"]])
        pods = pods_array([spec])
        for d1, d2 in f["moves"]:
            oracle.orc_game_make_move(pods, d1, d2)
        obs = _encode(oracle, pods)[0]
        exp = np.asarray(f["expected"], np.float32)
        assert obs.shape == exp.shape, f["name"]
        assert np.abs(obs - exp).max() <= 1e-6, f["name"]


def test_encoder_matches_python_builder(oracle):
    specs = random_positions(96, 7, 7, seed=123)
    gold = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    assert np.abs(_encode(oracle, pods_array(specs)) - gold).max() <= 1e-6


def test_mlp_restatement_matches_reference_model():
    """tests/nn_ref.mlp_forward == PyRatMLP.predict of the real reference (export_onnx --verify uses 1e-5)."""
    sd = make_mlp_state_dict(0, 349)
    obs = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    g = np.load(GOLD / "mlp_7x7.npz")
    p1, p2, v1, v2 = mlp_forward(sd, obs)
    assert np.abs(p1 - g["policy_p1"]).max() <= 1e-5 and np.abs(p2 - g["policy_p2"]).max() <= 1e-5
    assert np.abs(v1 - g["value_p1"]).max() <= 1e-5 and np.abs(v2 - g["value_p2"]).max() <= 1e-5


def test_mlp_restatement_matches_reference_model_hidden_128():
    """The same for PyRatMLP(hidden_dim=128): the golden the zero-padded CUDA evaluator is held to."""
    sd = make_mlp_state_dict(1, 349, hidden=128)
    obs = np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    g = np.load(GOLD / "mlp_7x7_h128.npz")
    p1, p2, v1, v2 = mlp_forward(sd, obs)
    assert np.abs(p1 - g["policy_p1"]).max() <= 1e-5 and np.abs(p2 - g["policy_p2"]).max() <= 1e-5
    assert np.abs(v1 - g["value_p1"]).max() <= 1e-5 and np.abs(v2 - g["value_p2"]).max() <= 1e-5


def _golden_obs(oracle, w, h):
    if (w, h) == (7, 7):
        return np.load(GOLD / "flat_builder_7x7.npz")["obs"]
    pods = pods_array(random_positions(40, w, h, seed=321))
    obs = np.zeros((40, 7 * w * h + 6), np.float32)
    oracle.orc_encode(pods, 40, obs.ctypes.data_as(C.POINTER(C.c_float)))
    return obs


@pytest.mark.parametrize("w,h", [(7, 7), (5, 5)])
def test_symmetric_and_cnn_restatements_match_reference_models(oracle, w, h):
    """nn_ref.symmetric_forward / cnn_forward == SymmetricMLP.predict / PyRatCNN.predict of the real
    reference (goldens written by tests/golden/make_golden.py), same 1e-5 bar as export_onnx --verify."""
    obs = _golden_obs(oracle, w, h)
    cases = [(f"symmetric_{w}x{h}.npz", symmetric_forward(make_symmetric_state_dict(2, w, h), obs, w, h))]
    for tag, blocks in (("gpool", ("res", "res", "gpool")), ("res", ("res",))):
        cases.append((f"cnn_{tag}_{w}x{h}.npz", cnn_forward(make_cnn_state_dict(3, blocks), obs, w, h)))
    for name, out in cases:
        g = np.load(GOLD / name)
        for a, k in zip(out, ("policy_p1", "policy_p2", "value_p1", "value_p2")):
            assert np.abs(a - g[k]).max() <= 1e-5, (name, k)


def test_symmetric_restatement_matches_reference_model_hidden_128(oracle):
    """SymmetricMLP(hidden_dim=128) of the real reference: the golden the zero-padded CUDA evaluator is held to."""
    obs = _golden_obs(oracle, 7, 7)
    out = symmetric_forward(make_symmetric_state_dict(4, 7, 7, hidden=128), obs, 7, 7)
    g = np.load(GOLD / "symmetric_7x7_h128.npz")
    for a, k in zip(out, ("policy_p1", "policy_p2", "value_p1", "value_p2")):
        assert np.abs(a - g[k]).max() <= 1e-5, k


def test_cnn_restatement_matches_reference_model_32_channels(oracle):
    """PyRatCNN with a 32-channel res / gpool16 / res trunk of the real reference: the golden the zero-padded CUDA
    evaluator is held to."""
    obs = _golden_obs(oracle, 7, 7)
    out = cnn_forward(make_cnn_state_dict(6, ("res", "gpool", "res"), channels=32, gpool_channels=16), obs, 7, 7)
    g = np.load(GOLD / "cnn_gpool_7x7_c32.npz")
    for a, k in zip(out, ("policy_p1", "policy_p2", "value_p1", "value_p2")):
        assert np.abs(a - g[k]).max() <= 1e-5, k


def test_make_unmake_roundtrip(oracle):
    for spec in random_positions(40, 7, 7, seed=9):
        pod = pods_array([spec])
        for d1 in range(5):
            for d2 in range(5):
                assert oracle.orc_game_make_unmake_roundtrip(pod, d1, d2) == 1


def test_search_is_seed_deterministic_and_clean(oracle):
    """python/tests/test_search.py:111-135 on the open 5x5 centre position."""
    spec = GameSpec(5, 5, 100, (2, 2), (2, 2), [(0, 0), (4, 4), (0, 4), (4, 0), (1, 3)])
    pod = pods_array([spec])[0]
    for sims in (10, 50, 100, 200):
        cfg = search_cfg(simulations=sims, batch_size=8)
        rc, a, clean = oracle_search(oracle, pod, cfg, 42)
        rc2, b, _ = oracle_search(oracle, pod, cfg, 42)
        assert rc == rc2 == 0 and clean
        assert bytes(a) == bytes(b)
        assert a.total_visits == sims
        assert abs(sum(a.policy_p1) - 1) < 1e-5 and abs(sum(a.policy_p2) - 1) < 1e-5
        assert sum(a.raw_visits_p1) == sims - 1 == sum(a.raw_visits_p2)


def test_constant_value_backend_values_propagate(oracle):
    """ConstantValueBackend (backend.rs:114-129): root value = leaf value when no cheese is reachable."""
    spec = GameSpec(5, 5, 100, (0, 0), (4, 4), [(2, 2)], walls=[((2, 2), (2, 3)), ((2, 2), (3, 2)), ((2, 2), (1, 2)), ((2, 2), (2, 1))])
    pod = pods_array([spec])[0]
    cfg = search_cfg(simulations=64, batch_size=8)
    rc, r, clean = oracle_search(oracle, pod, cfg, 3, const_values=(2.5, 1.25))
    assert rc == 0 and clean
    assert abs(r.value_p1 - 2.5) < 1e-5 and abs(r.value_p2 - 1.25) < 1e-5


def test_backend_error_is_propagated_and_tree_cleaned(oracle):
    """FailingBackend (search.rs:3161-3170, 3721-3737): error code surfaces, virtual losses reverted."""
    from conftest import EVAL_CB

    cb = EVAL_CB(lambda *a: 7)
    spec = GameSpec(5, 5, 100, (0, 0), (4, 4), [(2, 2), (1, 1)])
    rc, _, clean = oracle_search(oracle, pods_array([spec])[0], search_cfg(simulations=32), 1, eval_cb=cb)
    assert rc == 7 and clean


def test_selfplay_stats_identities(oracle):
    n = 48
    specs = make_games(n, width=5, height=5, cheese_count=5, max_turns=30)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=100, batch_size=8)
    summ, pos, stride, st = oracle_selfplay(oracle, pods, cfg, list(range(n)), n_threads=4)
    again = oracle_selfplay(oracle, pods, cfg, list(range(n)), n_threads=1)
    assert bytes(summ) == bytes(again[0]), "results must not depend on the worker count"
    assert st.total_games == n == st.p1_wins + st.p2_wins + st.draws
    assert st.total_positions == sum(summ[i].n_positions for i in range(n))
    assert st.total_cheese_available == 5 * n
    for i in range(n):
        s = summ[i]
        assert s.game_index == i and 1 <= s.n_positions <= 30
        collected = sum(1 for c in bytes(s.cheese_outcomes)[:25] if c != 2)
        assert collected == s.final_p1_score + s.final_p2_score
        total = 0
        for t in range(s.n_positions):
            r = pos[i * stride + t].search
            assert r.nn_evals + r.terminals == 100
            total += r.total_visits
        assert total == s.total_simulations


@pytest.mark.skipif(not Path("/root/reference/alpharat/eval/game.py").exists(),
                    reason="the reference tree is only mounted in the build container")
def test_termination_rule_matches_the_reference_python(oracle):
    """`check_game_over` belongs to the third-party engine; the only statement of the rule inside the reference
    is `is_terminal` (alpharat/eval/game.py:31-44: turn limit, no cheese left, strict majority of the total).  The
    oracle's restatement is compared with that function itself (imported without the package __init__, which needs
    the engine) on random states, half-point scores included."""
    import importlib
    import sys
    import types

    import numpy as np

    from alpharat_b200.games import GameSpec, pods_array

    saved = {k: sys.modules.get(k) for k in ("alpharat", "alpharat.eval", "alpharat.eval.game")}
    try:
        for name, path in (("alpharat", "/root/reference/alpharat"), ("alpharat.eval", "/root/reference/alpharat/eval")):
            pkg = types.ModuleType(name)
            pkg.__path__ = [path]
            sys.modules[name] = pkg
        is_terminal = importlib.import_module("alpharat.eval.game").is_terminal

        class Fake:
            def __init__(self, spec):
                self.turn, self.max_turns = spec.turn, spec.max_turns
                self.player1_score, self.player2_score = spec.p1_score, spec.p2_score
                self._cheese = list(spec.cheese)

            def cheese_positions(self):
                return self._cheese

        r = np.random.default_rng(7)
        cells = [(x, y) for y in range(5) for x in range(5)]
        specs = []
        for _ in range(4000):
            k = int(r.integers(0, 8))
            cheese = [cells[i] for i in r.choice(25, size=k, replace=False)]
            mt = int(r.integers(1, 40))
            specs.append(GameSpec(5, 5, mt, (0, 0), (4, 4), cheese, turn=int(r.integers(0, mt + 3)),
                                  p1_score=float(r.integers(0, 13)) / 2, p2_score=float(r.integers(0, 13)) / 2))
        pods = pods_array(specs)
        got = [bool(oracle.orc_game_over(C.byref(pods[i]))) for i in range(len(specs))]
        want = [bool(is_terminal(Fake(s))) for s in specs]
        assert got == want
        assert 0.2 < sum(want) / len(want) < 0.95  # both outcomes are exercised
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
