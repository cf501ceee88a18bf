"""CPU parity of the thread-per-tree device core (alpharat_b200/csrc/tree_thread.cuh).

The header is plain C++ apart from its vector loads and the warp-cooperative compaction, so the
same per-tree state machine the CUDA kernel runs is compiled here for the host (tests/tt_emul) and
stepped round-robin over T logical threads.  Bar: bit-exact against the oracle — records, raw visit
tables, counters, f32 policy / value arrays — and every borrowed arena page returned at the end.
The CUDA build of the same header is checked against the same oracle by the `-m gpu` tests.
"""

from __future__ import annotations

import pytest

from alpharat_b200.engine import search_cfg
from alpharat_b200.games import GameSpec, make_games, pods_array
from conftest import oracle_search, oracle_selfplay
from test_gpu_parity_uniform import assert_result_equal, compare_selfplay
from tt_emul_loader import emul_search, emul_selfplay, load_emul


@pytest.fixture(scope="module")
def emul():
    return load_emul()


def _check(emul, oracle, specs, cfg, seeds, n_threads, n_pages=None):
    pods = pods_array(specs)
    n = len(specs)
    g = emul_selfplay(emul, pods, cfg, seeds, n_threads=n_threads, n_pages=n_pages)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay((g[0], g[1], g[2], None), cpu, n)
    assert g[3][0] == cpu[3].path_nodes and g[3][1] == cpu[3].new_nodes
    return g[3]


def test_config_a_5x5(emul, oracle):
    n = 96
    _check(emul, oracle, make_games(n, width=5, height=5, cheese_count=5, max_turns=30),
           search_cfg(simulations=100, batch_size=8), list(range(n)), n_threads=16)


def test_config_b_7x7_tuned_multi_page_trees(emul, oracle):
    """Trees outgrow their first 1024-record page: page allocation, multi-visit levels (collision budget > 1),
    compaction across pages and page release are all exercised; a tight arena makes the allocator scan."""
    n = 40
    ctr = _check(emul, oracle, make_games(n, width=7, height=7, cheese_count=10, max_turns=50),
                 search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103),
                 [1000 + i for i in range(n)], n_threads=16, n_pages=16 * 8 + 80)
    assert ctr[2] > 16 + 8  # pages beyond the trees' own were in use at some point


def test_walls_mud_nonsquare(emul, oracle):
    n = 48
    _check(emul, oracle, make_games(n, width=7, height=5, cheese_count=6, max_turns=40, maze_type="classic",
                                    positions="random", first_index=4000),
           search_cfg(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103),
           [31 * i + 5 for i in range(n)], n_threads=32)


def test_dirichlet_noise(emul, oracle):
    n = 16
    _check(emul, oracle, make_games(n, width=7, height=7, cheese_count=10, max_turns=50, first_index=300),
           search_cfg(simulations=600, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
                      noise_epsilon=0.25, noise_concentration=10.83), [5000 + i for i in range(n)], n_threads=8)


def test_search_batch_edge_positions(emul, oracle):
    specs = make_games(16, width=5, height=5, cheese_count=5, max_turns=30)
    specs += [
        GameSpec(5, 5, 100, (2, 2), (2, 2), [(0, 0), (4, 4), (0, 4), (4, 0), (1, 3)]),
        GameSpec(5, 5, 100, (0, 0), (4, 0), [(2, 0)], walls=[((x, 0), (x, 1)) for x in range(5)]),
        GameSpec(5, 5, 100, (2, 3), (4, 4), [(0, 0)], mud=[((2, 2), (2, 3), 3)], p1_mud=3, turn=1),
        GameSpec(5, 5, 1, (0, 0), (0, 1), [(4, 4)], turn=1),  # terminal root
        GameSpec(7, 5, 80, (0, 0), (6, 4), [(3, 2), (6, 0)], walls=[((1, 1), (1, 2))], mud=[((4, 3), (4, 4), 2)]),
    ]
    pods = pods_array(specs)
    for sims, bs in ((10, 8), (100, 8), (100, 1), (300, 16), (3000, 16)):
        cfg = search_cfg(simulations=sims, batch_size=bs)
        seeds = [7 * i + sims for i in range(len(specs))]
        out = emul_search(emul, pods, cfg, seeds)
        for i in range(len(specs)):
            rc, ref, clean = oracle_search(oracle, pods[i], cfg, seeds[i])
            assert rc == 0 and clean
            assert_result_equal(out[i], ref, f"sims={sims} bs={bs} pos {i}")


def test_boards_over_64_cells(emul, oracle):
    """11x11 and 15x15 (the sizes of the reference's search benches, crates/alpharat-mcts/benches/search.rs:36-78) and the
    largest board an ar_game_pod holds (16x16): the thread engine's four-word cheese bitboard, self-play and search."""
    n = 10
    _check(emul, oracle, make_games(n, width=11, height=11, cheese_count=21, max_turns=40, first_index=11),
           search_cfg(simulations=200, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103),
           [70 + i for i in range(n)], n_threads=8)
    n = 6
    _check(emul, oracle, make_games(n, width=16, height=16, cheese_count=40, max_turns=30, maze_type="classic",
                                    positions="random", first_index=16),
           search_cfg(simulations=150, batch_size=8), [5 + i for i in range(n)], n_threads=4)
    specs = make_games(6, width=15, height=15, cheese_count=41, max_turns=200, first_index=15)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=2000, batch_size=64)
    out = emul_search(emul, pods, cfg, list(range(6)), n_threads=3)
    for i in range(6):
        rc, ref, clean = oracle_search(oracle, pods[i], cfg, i)
        assert rc == 0 and clean
        assert_result_equal(out[i], ref, f"15x15 pos {i}")
