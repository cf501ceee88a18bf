"""GPU parity (through the C-ABI) against the oracle: uniform-prior search and self-play.

Bar: bit-exact — raw integer root edge visits, total_visits / nn_evals / terminals / collisions,
node_count, sampled actions, scores, turns, cheese, and also the f32 policy / value / pruned
visit arrays (same arithmetic, same order, no FMA).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import GameSpec, make_games, pods_array
from conftest import oracle_search, oracle_selfplay

pytestmark = pytest.mark.gpu

# Both device engines of the uniform-prior path (include/alpharat_cuda.h: AR_TREE_WARP, AR_TREE_THREAD)
# are held to the same bit-exact bar.
ENGINES = pytest.mark.parametrize("tree_engine", ["warp", "thread", "half"])


def _bytes(x) -> bytes:
    return bytes(memoryview(x).cast("B"))


def assert_result_equal(a: N.SearchResultPod, b: N.SearchResultPod, ctx: str) -> None:
    for f, _ in N.SearchResultPod._fields_:
        va, vb = getattr(a, f), getattr(b, f)
        if hasattr(va, "__len__"):
            assert _bytes(va) == _bytes(vb), f"{ctx}: {f}: {list(va)} != {list(vb)}"
        else:
            assert va == vb or (isinstance(va, float) and np.float32(va).tobytes() == np.float32(vb).tobytes()), \
                f"{ctx}: {f}: {va} != {vb}"


def compare_selfplay(gpu, cpu, n):
    gs, gp, gstride, _ = gpu
    cs, cp, cstride, _ = cpu
    for i in range(n):
        a, b = gs[i], cs[i]
        assert a.n_positions == b.n_positions, f"game {i}: length {a.n_positions} != {b.n_positions}"
        for f in ("game_index", "final_p1_score", "final_p2_score", "result", "cheese_available",
                  "total_simulations", "total_nn_evals", "total_terminals", "total_collisions"):
            assert getattr(a, f) == getattr(b, f), f"game {i}: {f}: {getattr(a, f)} != {getattr(b, f)}"
        assert bytes(a.cheese_outcomes) == bytes(b.cheese_outcomes), f"game {i}: cheese_outcomes"
        for t in range(a.n_positions):
            pa, pb = gp[i * gstride + t], cp[i * cstride + t]
            for f in ("p1_x", "p1_y", "p2_x", "p2_y", "p1_mud", "p2_mud", "action_p1", "action_p2",
                      "turn", "p1_score", "p2_score"):
                assert getattr(pa, f) == getattr(pb, f), f"game {i} move {t}: {f}"
            assert bytes(pa.cheese) == bytes(pb.cheese), f"game {i} move {t}: cheese"
            assert_result_equal(pa.search, pb.search, f"game {i} move {t}")


@ENGINES
def test_selfplay_5x5_config_a(oracle, tree_engine):
    """BASELINE config 1, all 1000 games: 5x5 open, 5 cheese, 30 turns, 100 sims, batch 8."""
    n = 1000
    specs = make_games(n, width=5, height=5, cheese_count=5, max_turns=30)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=100, batch_size=8)
    seeds = list(range(n))
    with Engine(concurrent_games=128, max_turns=30, max_batch_size=8, max_simulations=100, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, seeds)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay(gpu, cpu, n)
    assert gpu[3].total_games == n
    assert gpu[3].total_positions == cpu[3].total_positions
    assert gpu[3].path_nodes == cpu[3].path_nodes
    assert gpu[3].new_nodes == cpu[3].new_nodes


@ENGINES
def test_selfplay_7x7_tuned(oracle, tree_engine):
    """BASELINE config 2 parameters (7x7_rust_tuned, noise 0), 512 games through 256 resident trees."""
    n = 512
    specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=50)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
    seeds = [1000 + i for i in range(n)]
    with Engine(concurrent_games=256, max_turns=50, max_batch_size=16, max_simulations=1897, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, seeds)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay(gpu, cpu, n)


@ENGINES
def test_selfplay_classic_mazes_with_walls_and_mud(oracle, tree_engine):
    """walls, mud timers (the [4,4,4,4,4] stuck outcome), random starts, non-square board."""
    n = 48
    specs = make_games(n, width=7, height=5, cheese_count=6, max_turns=40, maze_type="classic", positions="random",
                       first_index=4000)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
    seeds = [31 * i + 5 for i in range(n)]
    with Engine(concurrent_games=32, max_turns=40, max_batch_size=16, max_simulations=300, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, seeds)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay(gpu, cpu, n)


@ENGINES
def test_search_batch_matches_oracle(oracle, tree_engine):
    specs = make_games(32, width=5, height=5, cheese_count=5, max_turns=30)
    specs += [
        GameSpec(5, 5, 100, (2, 2), (2, 2), [(0, 0), (4, 4), (0, 4), (4, 0), (1, 3)]),
        GameSpec(5, 5, 100, (0, 0), (4, 0), [(2, 0)], walls=[((x, 0), (x, 1)) for x in range(5)]),
        GameSpec(5, 5, 100, (2, 3), (4, 4), [(0, 0)], mud=[((2, 2), (2, 3), 3)], p1_mud=3, turn=1),
        GameSpec(5, 5, 1, (0, 0), (0, 1), [(4, 4)], turn=1),  # terminal root
        GameSpec(7, 5, 80, (0, 0), (6, 4), [(3, 2), (6, 0)], walls=[((1, 1), (1, 2))], mud=[((4, 3), (4, 4), 2)]),
    ]
    pods = pods_array(specs)
    for sims, bs in ((10, 8), (50, 8), (100, 8), (200, 8), (100, 1), (300, 16)):
        cfg = search_cfg(simulations=sims, batch_size=bs)
        seeds = [7 * i + sims for i in range(len(specs))]
        with Engine(concurrent_games=64, max_turns=100, max_batch_size=16, max_simulations=sims, pool_nodes=1024,
                    tree_engine=tree_engine) as eng:
            out = eng.search_batch(pods, cfg, seeds)
        for i in range(len(specs)):
            rc, ref, clean = oracle_search(oracle, pods[i], cfg, seeds[i])
            assert rc == 0 and clean
            assert_result_equal(out[i], ref, f"sims={sims} bs={bs} pos {i}")


def test_large_scale_properties():
    """Full-size config 2 slice: properties that need no oracle (accounting identities)."""
    n = 512
    specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=50, first_index=10_000)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
    with Engine(concurrent_games=n, max_turns=50, max_batch_size=16, max_simulations=1897) as eng:
        summ, pos, stride, st = eng.selfplay(pods, cfg, list(range(n)))
    for i in range(n):
        s = summ[i]
        assert 1 <= s.n_positions <= 50
        assert s.final_p1_score + s.final_p2_score <= 10
        prev_total = 0
        for t in range(s.n_positions):
            r = pos[i * stride + t].search
            assert sum(r.raw_visits_p1) == r.total_visits - 1 == sum(r.raw_visits_p2)
            assert r.nn_evals + r.terminals == 1897
            assert abs(sum(r.policy_p1) - 1.0) < 1e-5 and abs(sum(r.policy_p2) - 1.0) < 1e-5
            assert r.total_visits >= 1897
        collected = sum(1 for c in bytes(s.cheese_outcomes)[:49] if c != 2)
        assert collected == s.final_p1_score + s.final_p2_score
    assert st.total_games == n


@ENGINES
def test_selfplay_with_dirichlet_noise(oracle, tree_engine):
    """Production sampling config (`7x7_rust_tuned.yaml`: noise_epsilon 0.25, concentration 10.83).

    The Gamma sampler is the oracle's restatement (Marsaglia-Tsang over a polar normal), drawn
    from the same per-game stream, so GPU and oracle agree draw for draw; against the real
    reference (rand_distr's ziggurat) the noise agrees in distribution only.
    """
    n = 24
    specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=50, first_index=300)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=600, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
                     noise_epsilon=0.25, noise_concentration=10.83)
    seeds = [5000 + i for i in range(n)]
    with Engine(concurrent_games=n, max_turns=50, max_batch_size=16, max_simulations=600, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, seeds)
        quiet = eng.selfplay(pods, search_cfg(simulations=600, batch_size=16, c_puct=0.512,
                                              fpu_reduction=0.459, force_k=0.103), seeds)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay(gpu, cpu, n)
    # the noise is really applied: root priors of the first move are no longer uniform
    pri = np.array(gpu[1][0].search.prior_p1[:])
    assert abs(pri.sum() - 1.0) < 1e-5 and pri.std() > 1e-3
    assert list(quiet[1][0].search.prior_p1[:]) != list(gpu[1][0].search.prior_p1[:])
