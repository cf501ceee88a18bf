"""Host-side mirror of the reference interface: config, C-ABI surface, game generator, bundles."""

from __future__ import annotations

import ctypes as C
import inspect
import subprocess
from pathlib import Path

import numpy as np
import pytest
from pydantic import TypeAdapter, ValidationError

import alpharat_b200 as ab
from alpharat_b200 import _native as N
from alpharat_b200.bundle import BUNDLE_KEYS, write_bundles
from alpharat_b200.engine import search_cfg
from alpharat_b200.games import move_cost_table, GameSpec, make_games, maze_array, pods_array
from conftest import oracle_selfplay

ROOT = Path(__file__).resolve().parent.parent


def test_config_is_a_discriminated_union_and_round_trips():
    ta = TypeAdapter(ab.MCTSConfig)
    cuda = ta.validate_python({"backend": "cuda", "simulations": 1897, "c_puct": 0.512, "batch_size": 16})
    assert isinstance(cuda, ab.CudaMCTSConfig) and cuda.simulations == 1897
    rust = ta.validate_python({"backend": "rust", "simulations": 554})
    assert isinstance(rust, ab.RustMCTSConfig)
    assert ta.validate_python(cuda.model_dump()) == cuda  # batch metadata / manifest round trip
    with pytest.raises(ValidationError):
        ta.validate_python({"backend": "cuda", "not_a_field": 1})  # extra='forbid' like the reference
    assert set(ab.RustMCTSConfig.model_fields) - {"backend"} <= set(ab.CudaMCTSConfig.model_fields)
    noisy = ab.CudaMCTSConfig(noise_epsilon=0.25)
    assert noisy.for_evaluation().noise_epsilon == 0.0 and ab.CudaMCTSConfig().for_evaluation().noise_epsilon == 0.0


def test_self_play_signature_covers_rust_self_play():
    """Every keyword of rust_self_play (crates/alpharat-sampling/src/bindings.rs:268-304) is accepted."""
    ref = ("width height cheese_count max_turns num_games cheese_symmetric maze_type positions wall_density "
           "mud_density maze_symmetric simulations batch_size c_puct fpu_reduction force_k noise_epsilon "
           "noise_concentration collision_limit_min collision_limit_max collision_scaling_start "
           "collision_scaling_end collision_scaling_power num_threads output_dir max_games_per_bundle "
           "onnx_model_path device mux_max_batch_size cache_size progress").split()
    params = inspect.signature(ab.cuda_self_play).parameters
    assert [p for p in ref if p not in params] == []
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in params.values())


def test_library_loads_and_exports_every_declared_symbol():
    so = N.library_path()
    if not so.exists():
        subprocess.run(["python", "-m", "alpharat_b200.build"], check=True, cwd=ROOT)
    lib = N.load_library()
    header = (ROOT / "include" / "alpharat_cuda.h").read_text()
    for sym in N.EXPORTED_SYMBOLS:
        assert sym + "(" in header.replace(" (", "("), f"{sym} not declared in the header"
        assert getattr(lib, sym) is not None
    assert lib.ar_abi_version() == N.AR_ABI_VERSION


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("ALPHARAT_CUDA_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(N, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        N.load_library()


def test_pod_layout_matches_the_c_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu", '
                   "sizeof(ar_game_pod),sizeof(ar_search_cfg),sizeof(ar_search_result),sizeof(ar_position_record),"
                   "sizeof(ar_game_summary),sizeof(ar_stats),sizeof(ar_progress),sizeof(ar_engine_cfg),sizeof(ar_tensor_desc));}"
                   % (ROOT / "include" / "alpharat_cuda.h"))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    sizes = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    py = [C.sizeof(t) for t in (N.GamePod, N.SearchCfg, N.SearchResultPod, N.PositionRecord, N.GameSummary,
                                N.Stats, N.Progress, N.EngineCfg, N.TensorDesc)]
    assert sizes == py


def test_game_generator_properties():
    specs = make_games(200, width=7, height=7, cheese_count=10, max_turns=50)
    again = make_games(200, width=7, height=7, cheese_count=10, max_turns=50)
    for a, b in zip(specs, again):
        assert a.cheese == b.cheese  # SplitMix64(game_index): reproducible
    for s in specs:
        cells = set(s.cheese)
        assert len(cells) == 10 and (0, 0) not in cells and (6, 6) not in cells
        assert all((6 - x, 6 - y) in cells for x, y in cells)  # 180-degree symmetry
    odd = make_games(20, width=5, height=5, cheese_count=5, max_turns=30)
    assert all((2, 2) in s.cheese for s in odd)  # odd symmetric count uses the centre
    assert len({tuple(sorted(s.cheese)) for s in specs}) > 150
    shifted = make_games(3, width=7, height=7, cheese_count=10, max_turns=50, first_index=5)
    assert shifted[0].cheese == specs[5].cheese
    with pytest.raises(ValueError):
        make_games(1, width=5, height=5, cheese_count=5, max_turns=30, maze_type="hexagonal")


@pytest.mark.parametrize("maze_type,sym", [("classic", True), ("random", True), ("random", False)])
def test_random_maze_generator_is_connected_symmetric_and_reproducible(maze_type, sym):
    """Same knobs as make_games (bindings.rs:489-533): maze_type classic/random, random positions."""
    kw = dict(width=7, height=5, cheese_count=6, max_turns=40, maze_type=maze_type, positions="random",
              wall_density=0.6, mud_density=0.25, maze_symmetric=sym)
    specs = make_games(40, **kw)
    assert [(s.walls, s.mud, s.p1, s.p2) for s in specs] == [(s.walls, s.mud, s.p1, s.p2) for s in make_games(40, **kw)]
    delta = {0: (0, 1), 1: (1, 0), 2: (0, -1), 3: (-1, 0)}
    n_walls = n_mud = 0
    for s in specs:
        mc = move_cost_table(7, 5, s.walls, s.mud).reshape(5, 7, 4)
        seen, todo = {s.p1}, [s.p1]
        while todo:  # every cell reachable: MazeParams.connected
            x, y = todo.pop()
            for d, (dx, dy) in delta.items():
                if mc[y, x, d] and (x + dx, y + dy) not in seen:
                    seen.add((x + dx, y + dy))
                    todo.append((x + dx, y + dy))
        assert len(seen) == 35
        if sym or maze_type == "classic":
            assert (mc[::-1, ::-1, :][:, :, [2, 3, 0, 1]] == mc).all()
        assert set(np.unique(mc)) <= {0, 1, 2, 3}
        assert s.p1 != s.p2 and s.p2 == (6 - s.p1[0], 4 - s.p1[1])
        assert s.p1 not in s.cheese and s.p2 not in s.cheese
        n_walls += len(s.walls)
        n_mud += len(s.mud)
    assert n_walls > 0 and n_mud > 0


def test_maze_array_and_pod_packing():
    spec = GameSpec(5, 5, 100, (1, 0), (4, 4), [(4, 0)], walls=[((2, 2), (2, 3))], mud=[((1, 0), (2, 0), 3)])
    m = maze_array(spec)
    assert m.shape == (5, 5, 4) and m.dtype == np.int8
    assert m[0, 0, 2] == -1 and m[0, 0, 3] == -1 and m[4, 4, 0] == -1 and m[4, 4, 1] == -1
    assert m[2, 2, 0] == -1 and m[3, 2, 2] == -1  # wall both ways
    assert m[0, 1, 1] == 3 and m[0, 2, 3] == 3  # mud both ways
    assert m[1, 1, 0] == 1
    pod = pods_array([spec])[0]
    assert (pod.width, pod.height, pod.p1_x, pod.p2_y, pod.max_turns) == (5, 5, 1, 4, 100)
    assert pod.cheese[0] == 1 << 4 and sum(pod.cheese) == 16
    assert pod.move_cost[(0 * 5 + 1) * 4 + 1] == 3 and pod.move_cost[(2 * 5 + 2) * 4 + 0] == 0


def test_bundle_format(oracle, tmp_path):
    """The 26 keys, dtypes and shapes of write_bundle (recording.rs:139-166; test_rust_sampling.py:115-213)."""
    n = 7
    specs = make_games(n, width=5, height=5, cheese_count=5, max_turns=30)
    pods = pods_array(specs)
    summ, pos, stride, st = oracle_selfplay(oracle, pods, search_cfg(simulations=50), list(range(n)), n_threads=2)
    paths = write_bundles(tmp_path / "games", specs, summ, pos, stride, max_games_per_bundle=3)
    assert len(paths) == 3 and all(p.name.startswith("bundle_") and p.suffix == ".npz" for p in paths)
    assert not list((tmp_path / "games").glob("*.tmp"))
    total_games = total_pos = 0
    for p in paths:
        z = np.load(p)
        assert set(z.files) == set(BUNDLE_KEYS) and len(BUNDLE_KEYS) == 26
        k, npos = len(z["game_lengths"]), int(z["game_lengths"].sum())
        total_games += k
        total_pos += npos
        assert z["game_lengths"].dtype == np.int32 and z["maze"].dtype == np.int8 and z["maze"].shape == (k, 5, 5, 4)
        assert z["initial_cheese"].dtype == np.bool_ and z["initial_cheese"].shape == (k, 5, 5)
        assert z["cheese_outcomes"].dtype == np.int8 and z["max_turns"].dtype == np.int16 and z["result"].dtype == np.int8
        assert z["final_p1_score"].dtype == np.float32
        assert z["p1_pos"].shape == (npos, 2) and z["p1_pos"].dtype == np.int8
        assert z["cheese_mask"].shape == (npos, 5, 5) and z["cheese_mask"].dtype == np.bool_
        assert z["turn"].dtype == np.int16 and z["action_p1"].dtype == np.int8
        for key in ("visit_counts_p1", "prior_p2", "policy_p1"):
            assert z[key].shape == (npos, 5) and z[key].dtype == np.float32
        assert z["value_p1"].shape == (npos,)
        assert np.allclose(z["policy_p1"].sum(1), 1, atol=1e-5)
        assert (z["initial_cheese"].sum((1, 2)) == 5).all()
        first = np.cumsum(np.concatenate([[0], z["game_lengths"][:-1]]))
        assert (z["cheese_mask"][first] == z["initial_cheese"]).all() and (z["turn"][first] == 0).all()
        assert set(np.unique(z["cheese_outcomes"])) <= {0, 1, 2, 3}
    assert total_games == n and total_pos == st.total_positions


# ---- run_cuda_sampling (drop-in for run_rust_sampling, alpharat/data/rust_sampling.py:137-296) -----------
class _Cheese:
    count, symmetric = 5, True


class _OpenMaze:
    type = "open"


class _RandomMaze:
    type, wall_density, mud_density, symmetric = "random", 0.5, 0.2, False


class _Game:
    width, height, max_turns, positions = 5, 5, 30, "corners"
    cheese = _Cheese()

    def __init__(self, maze=None):
        self.maze = maze or _OpenMaze()

    def model_dump(self):
        return {"width": 5, "height": 5}


class _FakeManager:
    """prepare_batch / register_batch with the signatures of experiments/manager.py:161-245."""

    def __init__(self, root):
        self.root, self.registered, self.prepared = Path(root), [], []

    def prepare_batch(self, group, mcts_config, game, checkpoint_path=None):
        d = self.root / "batches" / group / f"uuid{len(self.prepared)}"
        (d / "games").mkdir(parents=True)
        self.prepared.append((group, d.name, mcts_config, checkpoint_path))
        return d, d.name

    def register_batch(self, group, batch_uuid, mcts_config, game, checkpoint_path=None, created_at=None):
        self.registered.append((group, batch_uuid, mcts_config.model_dump(), checkpoint_path))


def test_run_cuda_sampling_signature_covers_run_rust_sampling():
    from alpharat_b200.sampling import CudaSamplingMetrics, run_cuda_sampling

    ref = ("game mcts num_games group num_threads max_games_per_bundle mux_max_batch_size checkpoint device "
           "cache_size experiments_dir verbose").split()
    params = inspect.signature(run_cuda_sampling).parameters
    assert [p for p in ref if p not in params] == []
    ref_metrics = ("total_games total_positions total_simulations elapsed_seconds p1_wins p2_wins draws "
                   "total_cheese_collected total_cheese_available min_turns max_turns total_nn_evals "
                   "total_terminals total_collisions cache_hits cache_misses").split()
    assert list(CudaSamplingMetrics.__dataclass_fields__) == ref_metrics
    for derived in ("games_per_second positions_per_second simulations_per_second avg_turns cheese_utilization "
                    "draw_rate nn_evals_per_second nn_eval_fraction terminal_fraction collision_fraction "
                    "cache_hit_rate").split():
        assert isinstance(getattr(CudaSamplingMetrics, derived), property)


def test_run_cuda_sampling_batch_protocol(oracle, tmp_path):
    """prepare -> play into batch_dir/games -> register only on success; kwargs flattened like the reference."""
    from alpharat_b200.sampling import resolve_sampling_device, run_cuda_sampling, self_play_kwargs
    from alpharat_b200.selfplay import SelfPlayStats

    seen = {}

    def oracle_backed_self_play(**kw):  # stands in for the GPU engine on a CPU-only box
        seen.update(kw)
        specs = make_games(kw["num_games"], width=kw["width"], height=kw["height"], cheese_count=kw["cheese_count"],
                           max_turns=kw["max_turns"])
        cfg = search_cfg(simulations=kw["simulations"], batch_size=kw["batch_size"], c_puct=kw["c_puct"])
        summ, pos, stride, st = oracle_selfplay(oracle, pods_array(specs), cfg, list(range(len(specs))), n_threads=2)
        write_bundles(Path(kw["output_dir"]), specs, summ, pos, stride, kw["max_games_per_bundle"])
        st.elapsed_secs = 0.5
        if kw.get("progress") is not None:
            assert kw["progress"].games_completed == 0
        return SelfPlayStats(st)

    mgr = _FakeManager(tmp_path)
    mcts = ab.CudaMCTSConfig(simulations=40, batch_size=8, c_puct=1.5, concurrent_games=64, seed=11)
    batch_dir, m = run_cuda_sampling(game=_Game(), mcts=mcts, num_games=6, group="uniform_5x5",
                                     max_games_per_bundle=4, device="cuda:0", verbose=False,
                                     experiment_manager=mgr, self_play_fn=oracle_backed_self_play)
    assert batch_dir == tmp_path / "batches" / "uniform_5x5" / "uuid0"
    assert mgr.registered == [("uniform_5x5", "uuid0", mcts.model_dump(), None)]
    assert mgr.registered[0][2]["backend"] == "cuda"
    assert len(list((batch_dir / "games").glob("bundle_*.npz"))) == 2
    assert seen["output_dir"] == str(batch_dir / "games") and seen["device"] == 0
    assert (seen["concurrent_games"], seen["seed"], seen["simulations"], seen["maze_type"]) == (64, 11, 40, "open")
    assert seen["tree_engine"] == "warp"  # CudaMCTSConfig default; "half" = two trees per warp
    assert "wall_density" not in seen and seen["cheese_symmetric"] is True and seen["positions"] == "corners"
    assert m.total_games == 6 and m.total_positions > 0 and m.elapsed_seconds == 0.5
    assert m.games_per_second == 12.0 and 0 < m.cheese_utilization <= 1 and m.avg_turns == m.total_positions / 6
    assert m.p1_wins + m.p2_wins + m.draws == 6

    kw = self_play_kwargs(_Game(_RandomMaze()), ab.RustMCTSConfig(simulations=7), 3)
    assert (kw["wall_density"], kw["mud_density"], kw["maze_symmetric"], kw["simulations"]) == (0.5, 0.2, False, 7)
    assert "concurrent_games" not in kw  # a RustMCTSConfig carries only the shared search fields

    def failing(**kw):
        raise RuntimeError("engine failure")

    with pytest.raises(RuntimeError, match="engine failure"):
        run_cuda_sampling(game=_Game(), mcts=mcts, num_games=2, group="g", verbose=True,
                          experiment_manager=mgr, self_play_fn=failing)
    assert len(mgr.prepared) == 2 and len(mgr.registered) == 1  # failed batch prepared, never registered

    assert [resolve_sampling_device(d) for d in ("auto", "cuda", "cuda:3", "b200", 5)] == [0, 0, 3, 0, 5]
    with pytest.raises(ValueError):
        resolve_sampling_device("coreml")


def test_cuda_self_play_hands_cache_size_to_the_engine(oracle, tmp_path):
    """`cache_size` (rust_self_play, sampling/bindings.rs:299) reaches `ar_engine_set_eval_cache` whenever an
    evaluator is loaded, and the statistics come straight from the engine (no host-side bookkeeping)."""
    calls = []

    class StubEngine:  # the surface cuda_self_play uses; play itself is the oracle's (CPU-only box)
        has_evaluator = True

        def set_eval_cache(self, n):
            calls.append(n)

        def selfplay(self, pods, cfg, seeds, progress=None):
            summ, pos, stride, st = oracle_selfplay(oracle, pods, cfg, seeds, n_threads=2)
            st.cache_hits, st.cache_misses = 7, 5
            return summ, pos, stride, st

        def close(self):
            calls.append("closed")

    eng = StubEngine()
    stats = ab.cuda_self_play(width=5, height=5, cheese_count=5, max_turns=20, num_games=3, simulations=30,
                              output_dir=str(tmp_path / "g"), cache_size=2048, seed=3, engine=eng)
    assert calls == [2048]  # a caller-owned engine is configured but not closed
    assert (stats.cache_hits, stats.cache_misses, round(stats.cache_hit_rate, 4)) == (7, 5, round(7 / 12, 4))
    assert len(list((tmp_path / "g").glob("bundle_*.npz"))) == 1
    eng.has_evaluator = False
    ab.cuda_self_play(width=5, height=5, cheese_count=5, max_turns=20, num_games=1, simulations=10,
                      output_dir=None, cache_size=2048, seed=3, engine=eng)
    assert calls == [2048]  # uniform priors: there is nothing to cache
    with pytest.raises(ValueError):
        ab.cuda_self_play(width=5, height=5, cheese_count=5, max_turns=20, num_games=1, simulations=10,
                          output_dir=None, cache_size=-1, engine=eng)
    with pytest.raises(ValueError, match="device"):  # an ORT provider name is not a GPU
        ab.cuda_self_play(width=5, height=5, cheese_count=5, max_turns=20, num_games=1, simulations=10,
                          output_dir=None, device="coreml", engine=eng)


REFERENCE = Path("/root/reference")


@pytest.mark.skipif(not (REFERENCE / "alpharat" / "data" / "loader.py").exists(),
                    reason="the reference tree is only mounted in the build container")
def test_bundles_load_with_the_reference_loader(oracle, tmp_path):
    """Drop-in check of the bundle format with the reference's own consumer: `load_game_bundle`
    (alpharat/data/loader.py:114-130, the loader behind sharding / training; the reference tests it against
    Rust-written bundles in tests/data/test_rust_bundle_parity.py) reads the files `write_bundles` produced and
    returns the games that were played.  Only loader.py and types.py are imported (the package __init__ needs the
    third-party engine)."""
    import importlib
    import sys
    import types

    saved = {k: sys.modules.get(k) for k in ("alpharat", "alpharat.data", "alpharat.data.types", "alpharat.data.loader")}
    try:
        for name, path in (("alpharat", REFERENCE / "alpharat"), ("alpharat.data", REFERENCE / "alpharat" / "data")):
            pkg = types.ModuleType(name)
            pkg.__path__ = [str(path)]
            sys.modules[name] = pkg
        loader = importlib.import_module("alpharat.data.loader")
        ref_types = importlib.import_module("alpharat.data.types")
        # the key names and outcome codes are the reference's own enums (alpharat/data/types.py:13-68)
        assert {k.value for k in ref_types.GameFileKey} - {"num_positions"} == set(BUNDLE_KEYS)
        assert [int(ref_types.CheeseOutcome[n]) for n in ("P1_WIN", "SIMULTANEOUS", "UNCOLLECTED", "P2_WIN")] == [0, 1, 2, 3]

        n = 5
        specs = (make_games(3, width=5, height=5, cheese_count=5, max_turns=30)
                 + make_games(2, width=7, height=5, cheese_count=6, max_turns=20, maze_type="classic", first_index=9))
        pods = pods_array(specs)
        summ, pos, stride, st = oracle_selfplay(oracle, pods, search_cfg(simulations=40), list(range(n)), n_threads=2)
        paths = write_bundles(tmp_path / "games", specs[:3], summ, pos, stride, max_games_per_bundle=8)
        assert len(paths) == 1 and loader.is_bundle_file(paths[0])
        games = loader.load_game_bundle(paths[0])
        assert len(games) == 3
        for g, game in enumerate(games):
            assert (game.width, game.height, game.max_turns) == (5, 5, 30)
            assert len(game.positions) == summ[g].n_positions
            assert (game.final_p1_score, game.final_p2_score) == (summ[g].final_p1_score, summ[g].final_p2_score)
            assert np.array_equal(game.maze, maze_array(specs[g]))
            assert sorted(map(tuple, np.argwhere(game.initial_cheese)[:, ::-1].tolist())) == sorted(specs[g].cheese)
            codes = np.frombuffer(bytes(summ[g].cheese_outcomes), dtype=np.uint8)[:25].reshape(5, 5)
            assert np.array_equal(np.asarray(game.cheese_outcomes), codes.astype(np.int8))
            collected = {int(ref_types.CheeseOutcome.P1_WIN): 1.0, int(ref_types.CheeseOutcome.SIMULTANEOUS): 0.5}
            p1_from_outcomes = sum(collected.get(int(c), 0.0) for c in codes[game.initial_cheese])
            assert p1_from_outcomes == game.final_p1_score  # the attribution adds up to the score
            for t, p in enumerate(game.positions):
                r = pos[g * stride + t]
                assert (tuple(p.p1_pos), tuple(p.p2_pos)) == ((r.p1_x, r.p1_y), (r.p2_x, r.p2_y))
                assert (p.action_p1, p.action_p2, p.turn) == (r.action_p1, r.action_p2, r.turn)
                assert np.array_equal(np.asarray(p.policy_p1, np.float32), np.asarray(list(r.search.policy_p1), np.float32))
                assert np.array_equal(np.asarray(p.visit_counts_p2, np.float32),
                                      np.asarray(list(r.search.visit_counts_p2), np.float32))
                assert np.float32(p.value_p1) == np.float32(r.search.value_p1)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not REFERENCE.exists(), reason="the reference tree is only mounted in the build container")
def test_interface_surface_matches_the_reference_sources():
    """The mirrored interface is compared with the reference's own sources (parsed, not imported: the Rust
    extension modules are absent): keywords and defaults of `rust_self_play` / `rust_mcts_search`, the fields and
    defaults of `RustMCTSConfig`, the fields of `SearchResult`, the getters of the `SelfPlayStats` pyclass."""
    import ast
    import dataclasses
    import re

    def pyo3_signature(path, fn):
        src = (REFERENCE / path).read_text()
        m = re.search(r"#\[pyo3\(signature = \((.*?)\)\)\]\s*(?:#\[[^\]]*\]\s*)*fn " + fn + r"\b", src, re.S)
        assert m, f"{fn} not found in {path}"
        out = {}
        for part in m.group(1).split(","):
            part = part.strip()
            if part and part != "*":
                name, _, default = part.partition("=")
                out[name.strip()] = default.strip() or None
        return out

    def same_default(rust, py):
        if rust in (None, "None"):
            return True  # required keyword, or None
        if rust in ("true", "false"):
            return py is (rust == "true")
        if rust.startswith('"'):
            return py == rust.strip('"') or (rust == '"auto"' and py == "cuda")  # device: a GPU, not an ORT provider
        return float(rust) == float(py)

    sig = inspect.signature(ab.cuda_self_play).parameters
    for name, default in pyo3_signature("crates/alpharat-sampling/src/bindings.rs", "rust_self_play").items():
        assert name in sig, f"cuda_self_play lacks {name}"
        py_default = sig[name].default
        if default is None:
            assert py_default is inspect.Parameter.empty or name == "output_dir", name
        else:
            assert same_default(default, py_default), (name, default, py_default)

    # CudaSearcher mirrors RustSearcher.__init__ (alpharat/mcts/searcher.py:43-59): same names, order and defaults
    # (predict_fn is replaced by checkpoint=), which in turn forwards to rust_mcts_search's keywords
    search = pyo3_signature("crates/alpharat-mcts/src/bindings.rs", "rust_mcts_search")
    tree = ast.parse((REFERENCE / "alpharat/mcts/searcher.py").read_text())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "RustSearcher")
    init = next(s for s in cls.body if isinstance(s, ast.FunctionDef) and s.name == "__init__")
    ref_args = [a.arg for a in init.args.args[1:]]
    ref_defaults = dict(zip(ref_args[len(ref_args) - len(init.args.defaults):], map(ast.literal_eval, init.args.defaults)))
    ctor = inspect.signature(ab.CudaSearcher.__init__).parameters
    ours = [n for n in ctor if n != "self"]
    shared = [a for a in ref_args if a != "predict_fn"]
    assert [n for n in ours if n in shared] == shared  # same relative order (positional call sites keep working)
    assert ours[:4] == ref_args[:4] == ["simulations", "c_puct", "force_k", "fpu_reduction"]
    for name in shared:
        if name in ref_defaults:
            assert ctor[name].default == ref_defaults[name], name
        else:
            assert ctor[name].default is inspect.Parameter.empty, name
        assert name in search  # every constructor keyword is a keyword of the Rust binding

    tree = ast.parse((REFERENCE / "alpharat/mcts/config.py").read_text())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "RustMCTSConfig")
    ref_fields = {s.target.id: ast.literal_eval(s.value) for s in cls.body
                  if isinstance(s, ast.AnnAssign) and s.value is not None}
    assert ref_fields.pop("backend") == "rust"
    ours = ab.CudaMCTSConfig()
    for name, default in ref_fields.items():
        assert getattr(ours, name) == default and getattr(ab.RustMCTSConfig(), name) == default, name
    for method in ("for_evaluation", "build_searcher", "build_agent"):
        assert any(isinstance(s, ast.FunctionDef) and s.name == method for s in cls.body) and hasattr(ours, method)

    tree = ast.parse((REFERENCE / "alpharat/mcts/result.py").read_text())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "SearchResult")
    assert [s.target.id for s in cls.body if isinstance(s, ast.AnnAssign)] == [f.name for f in dataclasses.fields(ab.SearchResult)]

    # run_cuda_sampling / CudaSamplingMetrics against alpharat/data/rust_sampling.py
    from alpharat_b200.sampling import CudaSamplingMetrics, run_cuda_sampling

    tree = ast.parse((REFERENCE / "alpharat/data/rust_sampling.py").read_text())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "run_rust_sampling")
    ref_kw = [a.arg for a in fn.args.kwonlyargs]
    ref_kw_defaults = {a.arg: ast.literal_eval(d) for a, d in zip(fn.args.kwonlyargs, fn.args.kw_defaults) if d is not None}
    ours_kw = inspect.signature(run_cuda_sampling).parameters
    assert [k for k in ref_kw if k not in ours_kw] == []
    for k, d in ref_kw_defaults.items():
        assert ours_kw[k].default == d or k == "device", k  # device: "cuda" here, an ORT provider name there
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "RustSamplingMetrics")
    assert [s.target.id for s in cls.body if isinstance(s, ast.AnnAssign)] == list(CudaSamplingMetrics.__dataclass_fields__)
    ref_props = [s.name for s in cls.body if isinstance(s, ast.FunctionDef)]
    assert ref_props and all(isinstance(getattr(CudaSamplingMetrics, p), property) for p in ref_props)

    src = (REFERENCE / "crates/alpharat-sampling/src/bindings.rs").read_text()
    stats_impl = src[src.index("impl PySelfPlayStats"):src.index("impl PySelfPlayProgress")] \
        if "impl PySelfPlayProgress" in src else src[src.index("impl PySelfPlayStats"):]
    getters = re.findall(r"#\[getter\]\s*fn (\w+)", stats_impl)
    assert len(getters) >= 20
    from alpharat_b200.selfplay import SelfPlayStats

    st = SelfPlayStats(N.Stats())
    missing = [g for g in getters if not hasattr(st, g)]
    assert missing == [], missing


@pytest.mark.skipif(not (REFERENCE / "alpharat" / "nn" / "extraction.py").exists(),
                    reason="the reference tree is only mounted in the build container")
def test_bundles_feed_the_reference_observation_builder(oracle, tmp_path):
    """The training side of the drop-in (tests/data/test_rust_bundle_parity.py::test_observation_building): a
    bundle written here, read back with the reference's loader, turned into `ObservationInput` by
    `from_game_arrays` (alpharat/nn/extraction.py:15-44) and encoded by `FlatObservationBuilder`
    (alpharat/nn/builders/flat.py:142-197) gives, position by position, the observation the encoder of the
    self-play path (oracle `orc_encode`, pinned to the Rust fixtures) produces for the replayed game."""
    import importlib
    import sys
    import types

    names = ("alpharat", "alpharat.data", "alpharat.nn", "alpharat.nn.builders", "alpharat.nn.training")
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "alpharat" or k.startswith("alpharat.")}
    try:
        for name in names:
            pkg = types.ModuleType(name)
            pkg.__path__ = [str(REFERENCE / name.replace(".", "/"))]
            sys.modules[name] = pkg
        loader = importlib.import_module("alpharat.data.loader")
        extraction = importlib.import_module("alpharat.nn.extraction")
        flat = importlib.import_module("alpharat.nn.builders.flat")

        specs = make_games(4, width=7, height=5, cheese_count=6, max_turns=25, maze_type="classic", first_index=3)
        pods = pods_array(specs)
        summ, pos, stride, st = oracle_selfplay(oracle, pods, search_cfg(simulations=40), [9, 8, 7, 6], n_threads=2)
        path = write_bundles(tmp_path / "games", specs, summ, pos, stride, max_games_per_bundle=8)[0]
        builder = flat.FlatObservationBuilder(width=7, height=5)
        dim = 7 * 35 + 6
        checked = 0
        for g, game in enumerate(loader.load_game_bundle(path)):
            replay = pods_array([specs[g]])
            for t, position in enumerate(game.positions):
                want = np.zeros(dim, dtype=np.float32)
                oracle.orc_encode(replay, 1, want.ctypes.data_as(C.POINTER(C.c_float)))
                got = np.asarray(builder.build(extraction.from_game_arrays(game, position)), dtype=np.float32)
                assert got.shape == want.shape and np.abs(got - want).max() <= 1e-6, (g, t)
                r = pos[g * stride + t]
                oracle.orc_game_make_move(replay, r.action_p1, r.action_p2)
                checked += 1
        assert checked == st.total_positions > 20
    finally:
        for k in [k for k in sys.modules if k == "alpharat" or k.startswith("alpharat.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
