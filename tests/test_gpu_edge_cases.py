"""Edge cases of the uniform-prior path on the GPU, bit-exact against the oracle: empty and ragged inputs,
the largest board (64 cells), the largest batch, aggressive collision budgets (the multi-visit level
builder), games that are over before they start, and the loud failures."""

from __future__ import annotations

import pytest

from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import GameSpec, make_games, pods_array
from conftest import oracle_search, oracle_selfplay
from test_gpu_parity_uniform import assert_result_equal, compare_selfplay

pytestmark = pytest.mark.gpu

# the two warp-resident engines of the uniform-prior path (one tree per warp, two trees per warp)
WARP_ENGINES = pytest.mark.parametrize("tree_engine", ["warp", "half"])


def test_empty_inputs():
    cfg = search_cfg(simulations=50, batch_size=8)
    with Engine(concurrent_games=8, max_turns=30, max_batch_size=8, max_simulations=50) as eng:
        summ, pos, stride, st = eng.selfplay(pods_array([]), cfg, [])
        assert st.total_games == 0 and st.total_positions == 0 and st.total_simulations == 0
        assert len(eng.search_batch(pods_array([]), cfg, [])) <= 1


@WARP_ENGINES
def test_ragged_boards_and_lengths_in_one_call(oracle, tree_engine):
    """Boards of different sizes and max_turns share one call and fewer slots than games."""
    specs = (make_games(6, width=5, height=5, cheese_count=5, max_turns=12)
             + make_games(5, width=7, height=7, cheese_count=10, max_turns=25, first_index=40)
             + make_games(4, width=7, height=5, cheese_count=6, max_turns=9, maze_type="classic", positions="random")
             + make_games(3, width=8, height=8, cheese_count=12, max_turns=17, first_index=70)
             + [GameSpec(4, 3, 5, (0, 0), (3, 2), [(1, 1), (2, 1)])])
    pods = pods_array(specs)
    n = len(specs)
    cfg = search_cfg(simulations=120, batch_size=8)
    seeds = [3 * i + 1 for i in range(n)]
    with Engine(concurrent_games=5, max_turns=25, max_batch_size=8, max_simulations=120, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, seeds)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay(gpu, cpu, n)


@WARP_ENGINES
def test_full_bitboard_8x8_and_max_batch(oracle, tree_engine):
    specs = make_games(10, width=8, height=8, cheese_count=20, max_turns=20, first_index=11)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=400, batch_size=64, c_puct=1.1, fpu_reduction=0.3, force_k=1.0,
                     collision_limit_min=4, collision_limit_max=64, collision_scaling_start=20,
                     collision_scaling_end=600, collision_scaling_power=0.7)
    seeds = [100 + i for i in range(10)]
    with Engine(concurrent_games=10, max_turns=20, max_batch_size=64, max_simulations=400, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, seeds)
    cpu = oracle_selfplay(oracle, pods, cfg, seeds)
    compare_selfplay(gpu, cpu, 10)
    assert gpu[3].total_collisions > 0


@WARP_ENGINES
def test_multi_visit_levels_and_tiny_searches(oracle, tree_engine):
    """Collision budgets far above 1 from the first batch (general build_gather_level with visits-to-change
    estimates and parked levels), and searches smaller than one batch."""
    specs = make_games(12, width=7, height=7, cheese_count=10, max_turns=50, first_index=300)
    pods = pods_array(specs)
    for sims, bs in ((3, 16), (16, 16), (500, 16), (257, 5)):
        cfg = search_cfg(simulations=sims, batch_size=bs, c_puct=0.512, fpu_reduction=0.459, force_k=0.103,
                         collision_limit_min=8, collision_limit_max=200, collision_scaling_start=0,
                         collision_scaling_end=300, collision_scaling_power=1.0)
        seeds = [sims + 13 * i for i in range(12)]
        with Engine(concurrent_games=12, max_turns=50, max_batch_size=16, max_simulations=sims, pool_nodes=2048, tree_engine=tree_engine) as eng:
            out = eng.search_batch(pods, cfg, seeds)
        for i in range(12):
            rc, ref, clean = oracle_search(oracle, pods[i], cfg, seeds[i])
            assert rc == 0 and clean
            assert_result_equal(out[i], ref, f"sims={sims} bs={bs} pos {i}")


@WARP_ENGINES
def test_games_that_are_over_before_they_start(oracle, tree_engine):
    specs = [GameSpec(5, 5, 10, (0, 0), (4, 4), [(2, 2)], turn=10),          # turn == max_turns
             GameSpec(5, 5, 10, (0, 0), (4, 4), [], turn=0),                  # no cheese
             GameSpec(5, 5, 10, (0, 0), (4, 4), [(2, 2)], p1_score=3.0),      # P1 already has the majority
             GameSpec(5, 5, 10, (1, 1), (3, 3), [(2, 2), (0, 4)])]            # a normal game next to them
    pods = pods_array(specs)
    cfg = search_cfg(simulations=60, batch_size=8)
    with Engine(concurrent_games=2, max_turns=10, max_batch_size=8, max_simulations=60, tree_engine=tree_engine) as eng:
        gpu = eng.selfplay(pods, cfg, [5, 6, 7, 8])
    cpu = oracle_selfplay(oracle, pods, cfg, [5, 6, 7, 8])
    compare_selfplay(gpu, cpu, 4)
    assert [gpu[0][i].n_positions for i in range(3)] == [0, 0, 0] and gpu[0][3].n_positions > 0


@WARP_ENGINES
def test_results_do_not_depend_on_the_number_of_resident_trees(tree_engine):
    specs = make_games(24, width=5, height=5, cheese_count=5, max_turns=15)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=80, batch_size=8)
    seeds = list(range(24))
    runs = []
    for conc in (1, 7, 24):
        with Engine(concurrent_games=conc, max_turns=15, max_batch_size=8, max_simulations=80, tree_engine=tree_engine) as eng:
            runs.append(eng.selfplay(pods, cfg, seeds))
    compare_selfplay(runs[0], runs[1], 24)
    compare_selfplay(runs[0], runs[2], 24)


def test_loud_failures():
    specs = make_games(4, width=7, height=7, cheese_count=10, max_turns=50)
    pods = pods_array(specs)
    with Engine(concurrent_games=4, max_turns=50, max_batch_size=8, max_simulations=2000, pool_nodes=256) as eng:
        with pytest.raises(ValueError, match="batch_size"):
            eng.selfplay(pods, search_cfg(simulations=100, batch_size=16), [0, 1, 2, 3])
        with pytest.raises(RuntimeError, match="pool"):  # 2000 simulations cannot fit 256 nodes
            eng.selfplay(pods, search_cfg(simulations=2000, batch_size=8), [0, 1, 2, 3])
        with pytest.raises(ValueError, match="max_turns"):
            eng.selfplay(pods_array(make_games(1, width=5, height=5, cheese_count=5, max_turns=80)),
                         search_cfg(simulations=10, batch_size=8), [0])
        with pytest.raises((NotImplementedError, RuntimeError), match="max_cells"):  # 81 cells > this engine's 64
            eng.selfplay(pods_array([GameSpec(9, 9, 10, (0, 0), (8, 8), [(4, 4)])]),
                         search_cfg(simulations=10, batch_size=8), [0])
    with pytest.raises(NotImplementedError, match="AR_TREE_THREAD"):  # big boards: thread engine only
        Engine(concurrent_games=4, max_turns=50, max_cells=256)
    with Engine(concurrent_games=4, max_turns=50, max_cells=256, tree_engine="thread") as eng:
        from nn_ref import make_mlp_state_dict
        from alpharat_b200 import _native as N

        with pytest.raises(NotImplementedError, match="64 cells"):  # evaluators keep the one-word bitboard
            eng.load_weights(N.AR_ARCH_MLP, 7, 7, make_mlp_state_dict(0, 349))


def test_boards_over_64_cells_on_the_thread_engine(oracle):
    """11x11 / 15x15 (the reference's search benches, crates/alpharat-mcts/benches/search.rs:36-78) and 16x16, the largest
    board an ar_game_pod holds: self-play and fresh-tree search on `tree_engine="thread"` (four-word cheese bitboard),
    bit-exact against the oracle."""
    cfg = search_cfg(simulations=200, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
    with Engine(concurrent_games=32, max_turns=40, max_batch_size=16, max_simulations=200, max_cells=256,
                tree_engine="thread") as eng:
        for specs, seeds in (
            (make_games(24, width=11, height=11, cheese_count=21, max_turns=40, first_index=11), [70 + i for i in range(24)]),
            (make_games(12, width=16, height=16, cheese_count=40, max_turns=30, maze_type="classic", positions="random",
                        first_index=16), [5 + i for i in range(12)]),
        ):
            pods = pods_array(specs)
            gpu = eng.selfplay(pods, cfg, seeds)
            cpu = oracle_selfplay(oracle, pods, cfg, seeds)
            compare_selfplay(gpu, cpu, len(specs))
    specs = make_games(16, width=15, height=15, cheese_count=41, max_turns=200, first_index=15)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=5000, batch_size=64)
    with Engine(concurrent_games=16, max_turns=200, max_batch_size=64, max_simulations=5000, max_cells=256,
                tree_engine="thread") as eng:
        out = eng.search_batch(pods, cfg, list(range(16)))
    for i in range(16):
        rc, ref, clean = oracle_search(oracle, pods[i], cfg, i)
        assert rc == 0 and clean
        assert_result_equal(out[i], ref, f"15x15 pos {i}")


def test_run_cuda_sampling_writes_a_registered_batch(oracle, tmp_path):
    """The `run_rust_sampling` drop-in end to end on the GPU: the bundles it leaves in the batch directory hold
    the oracle's games for the same seed (visit tables and actions bit for bit)."""
    import numpy as np

    from alpharat_b200 import CudaMCTSConfig
    from alpharat_b200.sampling import run_cuda_sampling
    from test_host_api import _FakeManager, _Game

    mgr = _FakeManager(tmp_path)
    mcts = CudaMCTSConfig(simulations=60, batch_size=8, concurrent_games=16, seed=5)
    batch_dir, m = run_cuda_sampling(game=_Game(), mcts=mcts, num_games=9, group="uniform_5x5",
                                     max_games_per_bundle=4, verbose=True, experiment_manager=mgr)
    assert len(mgr.registered) == 1 and m.total_games == 9
    files = sorted((batch_dir / "games").glob("bundle_*.npz"))
    assert len(files) == 3
    specs = make_games(9, width=5, height=5, cheese_count=5, max_turns=30, layout_seed=5)  # the run seed keys the boards
    cfg = search_cfg(simulations=60, batch_size=8)
    summ, pos, stride, st = oracle_selfplay(oracle, pods_array(specs), cfg, [5 + i for i in range(9)])
    assert (m.total_positions, m.total_simulations, m.p1_wins, m.p2_wins, m.draws) == (
        st.total_positions, st.total_simulations, st.p1_wins, st.p2_wins, st.draws)
    want = {}
    for g in range(9):
        k = summ[g].n_positions
        acts = tuple((pos[g * stride + t].action_p1, pos[g * stride + t].action_p2) for t in range(k))
        vis = tuple(tuple(float(v) for v in pos[g * stride + t].search.visit_counts_p1) for t in range(k))
        want[(acts, vis)] = want.get((acts, vis), 0) + 1
    got = {}
    for f in files:
        z = np.load(f)
        ends = np.cumsum(z["game_lengths"])
        for b, e in zip(np.concatenate([[0], ends[:-1]]), ends):
            acts = tuple(zip(z["action_p1"][b:e].tolist(), z["action_p2"][b:e].tolist()))
            vis = tuple(tuple(float(v) for v in row) for row in z["visit_counts_p1"][b:e])
            got[(acts, vis)] = got.get((acts, vis), 0) + 1
    assert got == want


@WARP_ENGINES
def test_streaming_batches_match_blocking_runs(oracle, tree_engine):
    """ar_stream_*: three batches in flight through two buffers give, game for game, the records of the
    blocking call and of the oracle (launches overlap on the device and share the tree slots)."""
    from conftest import oracle_selfplay
    from test_gpu_parity_uniform import compare_selfplay

    n = 96
    cfg = search_cfg(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
    batches = []
    for b in range(3):
        specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=50, first_index=7000 + b * n)
        batches.append((pods_array(specs), [900 + b * n + i for i in range(n)]))
    with Engine(concurrent_games=64, max_turns=50, max_batch_size=16, max_simulations=300, tree_engine=tree_engine) as eng:
        eng.stream_open(2, n, 50)
        eng.stream_submit(0, batches[0][0], cfg, batches[0][1])
        eng.stream_submit(1, batches[1][0], cfg, batches[1][1])
        got = [eng.stream_collect(0, n, 50)]
        eng.stream_submit(0, batches[2][0], cfg, batches[2][1])
        got.append(eng.stream_collect(1, n, 50))
        got.append(eng.stream_collect(0, n, 50))
        assert eng.stream_elapsed_ms(1, 0) > 0.0
        with pytest.raises(ValueError):
            eng.stream_collect(0, n, 50)  # nothing in flight
        eng.stream_close()
        blocking = eng.selfplay(batches[1][0], cfg, batches[1][1])
    for b in range(3):
        cpu = oracle_selfplay(oracle, batches[b][0], cfg, batches[b][1])
        compare_selfplay(got[b], cpu, n)
        assert got[b][3].total_games == n and got[b][3].path_nodes == cpu[3].path_nodes
    compare_selfplay(got[1], blocking, n)
