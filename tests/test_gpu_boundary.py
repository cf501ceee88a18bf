"""The drop-in boundary on hardware: `CudaSearcher.search(PyRat)`, `.pt` checkpoint ->
GPU evaluator -> search / self-play, and the gates the north star states for NN-guided search (L1 distance on
visit-proportional policies, absolute error on values) for every evaluator at its configuration's real
simulation count; plus the Dirichlet sampler's moments and the loud failure on a non-finite evaluator output."""

from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np
import pytest

from alpharat_b200 import _native as N
from alpharat_b200.config import CudaMCTSConfig
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import GameSpec, make_games, pack_pod, pods_array
from alpharat_b200.searcher import CudaSearcher
from conftest import EVAL_CB, oracle_search, oracle_selfplay
from nn_ref import (cnn_forward, make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict, mlp_forward,
                    symmetric_forward)
from test_gpu_nn_search import gpu_eval_callback
from test_gpu_parity_uniform import assert_result_equal, compare_selfplay

pytestmark = pytest.mark.gpu


class FakePyRat:
    """The duck-typed surface `Searcher.search` reads from a `PyRat` (CLAUDE.md "PyRat Game API",
    crates/alpharat-mcts/src/bindings.rs:228-304): attributes, not the GameSpec the other tests use."""

    def __init__(self, spec: GameSpec) -> None:
        self.width, self.height = spec.width, spec.height
        self.turn, self.max_turns = spec.turn, spec.max_turns
        self.player1_position = SimpleNamespace(x=spec.p1[0], y=spec.p1[1])
        self.player2_position = SimpleNamespace(x=spec.p2[0], y=spec.p2[1])
        self.player1_score, self.player2_score = spec.p1_score, spec.p2_score
        self.player1_mud_turns, self.player2_mud_turns = spec.p1_mud, spec.p2_mud
        self._cheese = [SimpleNamespace(x=x, y=y) for x, y in spec.cheese]
        self._walls = [SimpleNamespace(pos1=SimpleNamespace(x=a[0], y=a[1]), pos2=SimpleNamespace(x=b[0], y=b[1]))
                       for a, b in spec.walls]
        self._mud = [SimpleNamespace(pos1=SimpleNamespace(x=a[0], y=a[1]), pos2=SimpleNamespace(x=b[0], y=b[1]), value=v)
                     for a, b, v in spec.mud]

    def cheese_positions(self):
        return list(self._cheese)

    def wall_entries(self):
        return list(self._walls)

    def mud_entries(self):
        return list(self._mud)


PYRAT_CASES = [
    # walls + mud on a non-square board, mid-game scores
    GameSpec(7, 5, 80, (1, 1), (5, 3), [(3, 2), (6, 0), (0, 4)], walls=[((1, 1), (1, 2)), ((4, 3), (5, 3))],
             mud=[((4, 3), (4, 4), 2), ((2, 2), (3, 2), 3)], turn=11, p1_score=1.5, p2_score=0.5),
    # a player stuck in mud: its only outcome is STAY
    GameSpec(5, 5, 100, (2, 3), (4, 4), [(0, 0), (4, 0)], mud=[((2, 2), (2, 3), 3)], p1_mud=2, turn=7, p1_score=1.0),
    # both players on one cell next to a cheese, a corridor of walls
    GameSpec(5, 5, 60, (2, 2), (2, 2), [(2, 3), (0, 0)], walls=[((x, 0), (x, 1)) for x in range(5)], turn=3),
    # the turn limit is one move away; a game beyond the default 120-turn engine
    GameSpec(5, 5, 10, (0, 0), (4, 4), [(1, 0), (3, 4)], turn=9),
    GameSpec(7, 7, 200, (0, 0), (6, 6), [(3, 3), (1, 5), (5, 1)], turn=150, p1_score=2.0, p2_score=2.0),
]


def test_searcher_on_pyrat_shaped_games_matches_oracle(oracle):
    """`CudaSearcher.search(game)`: pod_from_pyrat (wall / mud entry unpacking, mud timers, scores) + the search
    equal the oracle's search on the same position, and the returned policies are the f64 renormalisation of
    the f32 result (alpharat/mcts/searcher.py:97-117)."""
    cfgm = CudaMCTSConfig(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103, seed=99)
    searcher = cfgm.build_searcher()
    cfg = search_cfg(simulations=300, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
    for k, spec in enumerate(PYRAT_CASES):
        game = FakePyRat(spec)
        pod = pack_pod(spec)
        raw = searcher.search_many([game], seeds=[99])[0]
        rc, ref, clean = oracle_search(oracle, pod, cfg, 99)
        assert rc == 0 and clean
        assert_result_equal(raw, ref, f"case {k}")
        res = searcher.search(game)  # seed=99 from the config: the same search again
        for got, want in ((res.policy_p1, ref.policy_p1), (res.policy_p2, ref.policy_p2)):
            w = np.asarray(want[:], dtype=np.float64)
            w = w / w.sum() if w.sum() > 0 else w
            assert got.dtype == np.float64 and np.array_equal(got, w), f"case {k}"
            assert abs(got.sum() - 1.0) < 1e-12
        assert res.total_visits == ref.total_visits
        assert np.array_equal(res.visit_counts_p1, np.asarray(ref.visit_counts_p1[:], dtype=np.float64))
        if spec.p1_mud:  # stuck in mud: 100 % STAY (search.rs:2570-2588)
            assert res.policy_p1[4] == 1.0


def _write_checkpoint(path, arch: str, sd: dict, model_cfg: dict) -> None:
    """The keys nn/training/loop.py:395-421 writes and load_model_from_checkpoint reads
    (alpharat/config/checkpoint.py:24-104); tensor names / shapes are those of the reference's own model classes
    (tests/test_reference_checkpoints.py checks that against PyRatMLP / SymmetricMLP / PyRatCNN)."""
    import torch

    torch.save({"model_state_dict": {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()},
                "config": {"model": dict(model_cfg, architecture=arch)}, "width": 7, "height": 7}, str(path))


CHECKPOINTS = [
    ("mlp", N.AR_ARCH_MLP, lambda: make_mlp_state_dict(0, 349), {"hidden_dim": 256}),
    ("symmetric", N.AR_ARCH_SYMMETRIC, lambda: make_symmetric_state_dict(2, 7, 7), {"hidden_dim": 256}),
    ("cnn", N.AR_ARCH_CNN, lambda: make_cnn_state_dict(3, ("res", "res", "gpool")), {}),
]


@pytest.mark.parametrize("arch,arch_id,make_sd,model_cfg", CHECKPOINTS)
def test_checkpoint_to_selfplay_and_searcher(oracle, tmp_path, arch, arch_id, make_sd, model_cfg):
    """`.pt` -> `cuda_self_play(checkpoint=...)` and `build_searcher(checkpoint=...)`: records and search results
    equal the oracle driven by the same device evaluator (weights loaded directly), bit for bit."""
    from alpharat_b200.selfplay import cuda_self_play

    sd = make_sd()
    ck = tmp_path / f"{arch}.pt"
    _write_checkpoint(ck, arch, sd, model_cfg)
    n = 4
    specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=16, first_index=321)
    kw = dict(simulations=120, batch_size=16, c_puct=0.512, fpu_reduction=0.479, force_k=0.025)
    stats, summ, pos, stride = cuda_self_play(width=7, height=7, cheese_count=10, max_turns=16, num_games=n, games=specs,
                                              seed=40, checkpoint=str(ck), output_dir=None, concurrent_games=4,
                                              pool_nodes=8192, return_records=True, **kw)
    cfg = search_cfg(**kw)
    pods = pods_array(specs)
    with Engine(concurrent_games=4, max_turns=16, max_batch_size=16, max_simulations=120, pool_nodes=8192) as eng:
        eng.load_weights(arch_id, 7, 7, sd)
        cb = gpu_eval_callback(eng)
        cpu = oracle_selfplay(oracle, pods, cfg, [40 + i for i in range(n)], n_threads=1, eval_cb=cb)
        compare_selfplay((summ, pos, stride, None), cpu, n)
        assert stats.total_nn_evals == cpu[3].total_nn_evals > 0
        searcher = CudaMCTSConfig(seed=7, **kw).build_searcher(checkpoint=str(ck))
        raw = searcher.search_many([FakePyRat(specs[0])], seeds=[7])[0]
        rc, ref, clean = oracle_search(oracle, pods[0], cfg, 7, eval_cb=cb)
        assert rc == 0 and clean
        assert_result_equal(raw, ref, arch)


def test_checkpoint_without_architecture_is_refused(tmp_path):
    import torch

    ck = tmp_path / "legacy.pt"
    torch.save({"model_state_dict": {}, "config": {"model": {"hidden_dim": 256}}, "width": 7, "height": 7}, str(ck))
    with Engine(concurrent_games=4, max_turns=16) as eng:
        from alpharat_b200.weights import load_checkpoint_into

        with pytest.raises(ValueError):
            load_checkpoint_into(eng, str(ck))


# Stated tolerances of the NN-guided gates (bf16 tcgen05 evaluator vs the fp32 restatement of the reference model,
# both inside the same search).  A search is a discontinuous function of its priors: at a near-tie between two moves
# a 1e-3 change of a prior flips the visit-proportional policy (L1 up to 2), so the bound is on the median and the
# 90th percentile; the maximum is printed.  Value errors are absolute (values are expected cheese counts, 0..10).
GATES = [
    # arch, arch id, state dict, fp32 forward, search parameters, positions, (L1 median, L1 p90, |dv| median, |dv| p90)
    ("mlp", N.AR_ARCH_MLP, lambda: make_mlp_state_dict(0, 349), mlp_forward,
     dict(simulations=1897, fpu_reduction=0.459, force_k=0.103), 48, (0.02, 0.25, 0.05, 0.30)),
    ("symmetric", N.AR_ARCH_SYMMETRIC, lambda: make_symmetric_state_dict(2, 7, 7),
     lambda sd, o: symmetric_forward(sd, o, 7, 7), dict(simulations=2693, fpu_reduction=0.479, force_k=0.025), 24,
     (0.06, 0.30, 0.10, 0.60)),
    ("cnn", N.AR_ARCH_CNN, lambda: make_cnn_state_dict(3, ("res", "res", "gpool")),
     lambda sd, o: cnn_forward(sd, o, 7, 7), dict(simulations=600, fpu_reduction=0.479, force_k=0.025), 6,
     (0.06, 0.30, 0.05, 0.30)),
]


@pytest.mark.parametrize("arch,arch_id,make_sd,fwd,sp,n_pos,tol", GATES)
def test_nn_guided_policy_l1_and_value_error(oracle, arch, arch_id, make_sd, fwd, sp, n_pos, tol):
    """North star: "L1 distance on visit-proportional policies and absolute error on values".  MLP at 1897 sims
    (7x7_rust_tuned), SymmetricMLP at 2693 (7x7_rust_strong); the CNN at 600 sims on 6 positions, because its fp32
    numpy restatement costs 0.14 s per evaluated batch on the host (2693 sims would be 25 s per position)."""
    sd = make_sd()
    pods = pods_array(make_games(n_pos, width=7, height=7, cheese_count=10, max_turns=50, first_index=4242))
    cfg = search_cfg(batch_size=16, c_puct=0.512, **sp)
    seeds = list(range(n_pos))

    def cb(user, states, n, p1, p2, v1, v2):
        sub = (N.GamePod * n).from_address(C.addressof(states.contents))
        obs = np.zeros((n, 349), np.float32)
        oracle.orc_encode(sub, n, obs.ctypes.data_as(C.POINTER(C.c_float)))
        for dst, src in zip((p1, p2, v1, v2), fwd(sd, obs)):
            src = np.ascontiguousarray(src, np.float32)
            C.memmove(dst, src.ctypes.data, src.nbytes)
        return 0

    cbp = EVAL_CB(cb)
    with Engine(concurrent_games=max(n_pos, 4), max_turns=50, max_batch_size=16, max_simulations=sp["simulations"],
                pool_nodes=sp["simulations"] + 64) as eng:
        eng.load_weights(arch_id, 7, 7, sd)
        out = eng.search_batch(pods, cfg, seeds)
    l1, dv = [], []
    for i in range(n_pos):
        rc, ref, _ = oracle_search(oracle, pods[i], cfg, seeds[i], eval_cb=cbp)
        assert rc == 0
        assert out[i].total_visits == ref.total_visits == sp["simulations"]
        for a, b in ((out[i].policy_p1, ref.policy_p1), (out[i].policy_p2, ref.policy_p2)):
            l1.append(float(np.abs(np.asarray(a[:]) - np.asarray(b[:])).sum()))
        for a, b in ((out[i].value_p1, ref.value_p1), (out[i].value_p2, ref.value_p2)):
            dv.append(abs(a - b))
    print(f"{arch} @ {sp['simulations']} sims, {n_pos} positions: policy L1 median {np.median(l1):.4f} "
          f"p90 {np.percentile(l1, 90):.4f} max {np.max(l1):.4f}; |dv| median {np.median(dv):.4f} "
          f"p90 {np.percentile(dv, 90):.4f} max {np.max(dv):.4f}")
    assert np.median(l1) <= tol[0] and np.percentile(l1, 90) <= tol[1], (np.median(l1), np.percentile(l1, 90))
    assert np.median(dv) <= tol[2] and np.percentile(dv, 90) <= tol[3], (np.median(dv), np.percentile(dv, 90))


def test_dirichlet_noise_moments():
    """apply_dirichlet_noise (search.rs:400-429) over 12288 roots: prior' = 0.75 * uniform + 0.25 * Dir(alpha 1),
    alpha = 10.83 / n.  For an n-outcome player the noise component has mean 1/n and variance
    (n - 1) / (n^2 (n alpha + 1)) = (n - 1) / (n^2 * 11.83).  One 1-simulation search per root (the root is evaluated,
    noised, and its priors are returned); interior cells have n = 5, a corner has n = 3."""
    eps, conc = 0.25, 10.83
    n_roots = 12288
    base = [GameSpec(7, 7, 50, (3, 3), (0, 0), [(6, 6), (1, 5)])]  # P1 centre: 5 outcomes, P2 corner: 3 outcomes
    pods = pods_array(base * n_roots)
    cfg = search_cfg(simulations=1, batch_size=1, noise_epsilon=eps, noise_concentration=conc)
    with Engine(concurrent_games=4096, max_turns=50, max_batch_size=1, max_simulations=1, pool_nodes=64) as eng:
        out = eng.search_batch(pods, cfg, list(range(1, n_roots + 1)))
    for pl, n, acts in ((1, 5, [0, 1, 2, 3, 4]), (2, 3, [0, 1, 4])):
        pri = np.array([[getattr(out[i], f"prior_p{pl}")[a] for a in acts] for i in range(n_roots)], dtype=np.float64)
        assert np.allclose(pri.sum(1), 1.0, atol=1e-5)
        noise = (pri - (1 - eps) / n) / eps
        assert (noise > -1e-5).all()
        mean, var = noise.mean(0), noise.var(0)
        want_var = (n - 1) / (n * n * (conc + 1))
        # standard errors over 12288 draws: mean ~ sqrt(var / N) ~ 1e-3, variance ~ a few percent
        assert np.abs(mean - 1.0 / n).max() < 6e-3, (pl, mean)
        assert np.abs(var / want_var - 1.0).max() < 0.10, (pl, var, want_var)
        blocked = [a for a in range(5) if a not in acts]
        for a in blocked:
            assert all(getattr(out[i], f"prior_p{pl}")[a] == 0.0 for i in range(0, n_roots, 257))


def test_nonfinite_evaluator_output_is_a_loud_error():
    """onnx.rs:233-241 rejects non-finite outputs with a BackendError; here AR_ERR_NONFINITE -> RuntimeError, from the
    forward pass alone and from inside a search."""
    sd = make_mlp_state_dict(0, 349)
    bad = dict(sd)
    key = next(k for k in bad if k.endswith("weight") and bad[k].ndim == 2)
    w = bad[key].copy()
    w[0, 0] = np.inf
    bad[key] = w
    pods = pods_array(make_games(8, width=7, height=7, cheese_count=10, max_turns=50))
    cfg = search_cfg(simulations=32, batch_size=8)
    with Engine(concurrent_games=8, max_turns=50, max_batch_size=8, max_simulations=32, pool_nodes=256) as eng:
        eng.load_weights(N.AR_ARCH_MLP, 7, 7, bad)
        with pytest.raises(RuntimeError, match="non-finite|status 4"):
            eng.nn_forward(pods)
        with pytest.raises(RuntimeError, match="non-finite|status 4"):
            eng.search_batch(pods, cfg, list(range(8)))
        eng.load_weights(N.AR_ARCH_MLP, 7, 7, sd)  # the engine recovers with good weights
        p1, _, v1, _ = eng.nn_forward(pods)
        assert np.isfinite(p1).all() and np.isfinite(v1).all()
