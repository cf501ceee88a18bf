"""CPU parity of the device code of the two warp-resident tree engines: two trees per warp
(alpharat_b200/csrc/mcts_half.cuh, `emul` = "half") and one tree per warp (mcts_device.cuh, `emul` = "warp", 32 fibres
running the per-warp loop of `selfplay_uniform_kernel`, tests/half_emul/warp_emul.cpp).

The header is compiled for the host behind a shim of the CUDA intrinsics (tests/half_emul/simt_shim.h): the 16 lanes of
a half run as 16 cooperative fibers, every shuffle / ballot / barrier is a rendezvous of the whole half, and the shim
aborts when the lanes do not execute the same sequence of collectives — the invariant the kernel's `__activemask()`
member masks rest on.  The per-half loop of `selfplay_half_kernel` is mirrored by tests/half_emul/half_emul.cpp; even
games run as lanes 0-15, odd games as lanes 16-31 (the ballot shifts and the +128-byte pool of the second half).
Bar: bit-exact against the oracle — records, raw visit tables, counters, f32 policy / value arrays.  What the emulation
does not cover: the two halves sharing one instruction stream (there is nothing to emulate: they exchange no data), and
the FFMA sequences of `div_guard<true>` / `sqrt_count<true>` (replaced by `/` and `sqrtf`, which is what they compute);
the CUDA build of the same header is held to the same oracle by the `-m gpu` tests.
"""

from __future__ import annotations

import pytest

from alpharat_b200.engine import search_cfg
from alpharat_b200.games import GameSpec, make_games, pods_array
from conftest import oracle_search, oracle_selfplay
from half_emul_loader import emul_search, emul_selfplay, load_emul, load_warp_emul
from test_gpu_parity_uniform import assert_result_equal, compare_selfplay


@pytest.fixture(scope="module", params=["half", "warp"])
def emul(request):
    e = load_emul() if request.param == "half" else load_warp_emul()
    e.scale = 1.0 if request.param == "half" else 0.34  # 32 fibres per tree are slower: fewer games for the warp code
    return e


def _n(emul, n):
    return max(2, int(n * emul.scale))


def _check(emul, oracle, specs, cfg, seeds, pool_nodes=8192):
    pods = pods_array(specs)
    n = len(specs)
    g = emul_selfplay(emul, pods, cfg, seeds, pool_nodes=pool_nodes)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay((g[0], g[1], g[2], None), cpu, n)
    assert g[3][0] == cpu[3].path_nodes and g[3][1] == cpu[3].new_nodes
    assert g[3][2] > 0  # collectives were executed (and checked) by the shim
    return g[3]


def test_config_a_5x5(emul, oracle):
    n = _n(emul, 96)
    _check(emul, oracle, make_games(n, width=5, height=5, cheese_count=5, max_turns=30),
           search_cfg(simulations=100, batch_size=8), list(range(n)))


def test_config_b_7x7_tuned(emul, oracle):
    """BASELINE config 2 parameters: tree reuse with in-place compaction, multi-visit levels once a tree holds 800
    nodes, parked split levels, forced playouts at the root."""
    n = _n(emul, 12)
    _check(emul, oracle, make_games(n, width=7, height=7, cheese_count=10, max_turns=50),
           search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103),
           [1000 + i for i in range(n)], pool_nodes=32768)


def test_walls_mud_nonsquare(emul, oracle):
    n = _n(emul, 32)
    _check(emul, oracle, make_games(n, width=7, height=5, cheese_count=6, max_turns=40, maze_type="classic",
                                    positions="random", first_index=4000),
           search_cfg(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103),
           [31 * i + 5 for i in range(n)])


def test_full_bitboard_8x8_large_batches_and_collision_budgets(emul, oracle):
    n = _n(emul, 10)
    _check(emul, oracle, make_games(n, width=8, height=8, cheese_count=20, max_turns=20, first_index=11),
           search_cfg(simulations=400, batch_size=64, c_puct=1.1, fpu_reduction=0.3, force_k=1.0,
                      collision_limit_min=4, collision_limit_max=64, collision_scaling_start=20,
                      collision_scaling_end=600, collision_scaling_power=0.7), [100 + i for i in range(n)])


def test_dirichlet_noise(emul, oracle):
    """Root noise on: the root's priors are no longer uniform, so its selections take the general path (prior shuffles,
    lane-chain sum of the visited mass) while every other node uses the tables; the Gamma sampler's f64 libm calls are
    the oracle's own on the host."""
    n = _n(emul, 8)
    _check(emul, oracle, make_games(n, width=7, height=7, cheese_count=10, max_turns=50, first_index=300),
           search_cfg(simulations=600, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
                      noise_epsilon=0.25, noise_concentration=10.83), [5000 + i for i in range(n)])


def test_results_do_not_depend_on_the_order_in_which_the_lanes_run(emul, oracle, monkeypatch):
    """The fibres of the shim run one after the other between two collectives; SIMT_REVERSE runs them in the opposite
    order.  Code that needs a barrier it does not have (one lane stores, another reads) gives order-dependent results
    here.  mcts_half.cuh passes as it is; the one-tree-per-warp code relies on the warp's lockstep in `save_path`, which
    is marked there (AR_LOCKSTEP: a barrier in this host build only)."""
    monkeypatch.setenv("SIMT_REVERSE", "1")
    n = _n(emul, 12)
    _check(emul, oracle, make_games(n, width=7, height=5, cheese_count=6, max_turns=40, maze_type="classic",
                                    positions="random", first_index=4000),
           search_cfg(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103),
           [31 * i + 5 for i in range(n)])


def test_multi_visit_levels_and_tiny_searches(emul, oracle):
    """Collision budgets far above 1 from the first batch (build_level with visits-to-change estimates and parked
    levels), and searches smaller than one batch."""
    specs = make_games(8, width=7, height=7, cheese_count=10, max_turns=50, first_index=300)
    pods = pods_array(specs)
    for sims, bs in ((3, 16), (16, 16), (500, 16), (257, 5)):
        cfg = search_cfg(simulations=sims, batch_size=bs, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
                         collision_limit_min=8, collision_limit_max=200, collision_scaling_start=0,
                         collision_scaling_end=300, collision_scaling_power=1.0)
        seeds = [sims + 13 * i for i in range(8)]
        out = emul_search(emul, pods, cfg, seeds, pool_nodes=2048)
        for i in range(8):
            rc, ref, clean = oracle_search(oracle, pods[i], cfg, seeds[i])
            assert rc == 0 and clean
            assert_result_equal(out[i], ref, f"sims={sims} bs={bs} pos {i}")


def test_games_that_are_over_before_they_start(emul, oracle):
    specs = [GameSpec(5, 5, 10, (0, 0), (4, 4), [(2, 2)], turn=10), GameSpec(5, 5, 10, (0, 0), (4, 4), [], turn=0),
             GameSpec(5, 5, 10, (0, 0), (4, 4), [(2, 2)], p1_score=3.0), GameSpec(5, 5, 10, (1, 1), (3, 3), [(2, 2), (0, 4)])]
    _check(emul, oracle, specs, search_cfg(simulations=60, batch_size=8), [5, 6, 7, 8])


def test_search_batch_edge_positions(emul, oracle):
    specs = make_games(8, width=5, height=5, cheese_count=5, max_turns=30)
    specs += [
        GameSpec(5, 5, 100, (2, 2), (2, 2), [(0, 0), (4, 4), (0, 4), (4, 0), (1, 3)]),
        GameSpec(5, 5, 100, (0, 0), (4, 0), [(2, 0)], walls=[((x, 0), (x, 1)) for x in range(5)]),
        GameSpec(5, 5, 100, (2, 3), (4, 4), [(0, 0)], mud=[((2, 2), (2, 3), 3)], p1_mud=3, turn=1),
        GameSpec(5, 5, 1, (0, 0), (0, 1), [(4, 4)], turn=1),  # terminal root
        GameSpec(7, 5, 80, (0, 0), (6, 4), [(3, 2), (6, 0)], walls=[((1, 1), (1, 2))], mud=[((4, 3), (4, 4), 2)]),
    ]
    pods = pods_array(specs)
    for sims, bs in ((10, 8), (100, 8), (100, 1), (300, 16)):
        cfg = search_cfg(simulations=sims, batch_size=bs)
        seeds = [7 * i + sims for i in range(len(specs))]
        out = emul_search(emul, pods, cfg, seeds)
        for i in range(len(specs)):
            rc, ref, clean = oracle_search(oracle, pods[i], cfg, seeds[i])
            assert rc == 0 and clean
            assert_result_equal(out[i], ref, f"sims={sims} bs={bs} pos {i}")
