"""NN-guided self-play throughput for BASELINE configs 3 and 4 (random-init checkpoints, synthetic games).

    python scripts/bench_nn_configs.py [mlp symmetric cnn]

config 3: 7x7_rust_tuned (1897 sims, c_puct 0.512, fpu 0.459, force_k 0.103), MLP, 16384 concurrent games
config 4: 7x7_rust_strong (2693 sims, c_puct 0.512, fpu 0.479, force_k 0.025), SymmetricMLP and CNN-gpool
AR_EVAL_CACHE=<entries per tree> turns the evaluation cache on (rust_self_play's cache_size).
Prints one JSON line per run (device time from CUDA events around the whole run).
Config 4 runs with 4096 resident trees: 2693-sim searches guided by a peaked value head keep most of the
tree across moves (tens of thousands of nodes per tree), and 16384 trees would leave only 29k nodes each."""
import json
import os
import sys

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
from nn_ref import make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict

CACHE = int(os.environ.get("AR_EVAL_CACHE", "0"))  # evaluation-cache entries per resident tree (cache_size)
FLOPS = {"mlp": 315_904, "symmetric": 976_896, "cnn": 22.2e6}
RUNS = {
    "mlp": dict(arch=N.AR_ARCH_MLP, sd=lambda: make_mlp_state_dict(0, 349), conc=16384, n=32768, sims=1897, fpu=0.459, fk=0.103),
    "symmetric": dict(arch=N.AR_ARCH_SYMMETRIC, sd=lambda: make_symmetric_state_dict(2, 7, 7), conc=4096, n=8192, sims=2693, fpu=0.479, fk=0.025),
    "cnn": dict(arch=N.AR_ARCH_CNN, sd=lambda: make_cnn_state_dict(3, ("res", "res", "gpool")), conc=4096, n=4096, sims=2693, fpu=0.479, fk=0.025),
}
for name in (sys.argv[1:] or ["mlp", "symmetric", "cnn"]):
    r = RUNS[name]
    specs = make_games(r["n"], width=7, height=7, cheese_count=10, max_turns=50)
    pods = pods_array(specs)
    cfg = search_cfg(simulations=r["sims"], batch_size=16, c_puct=0.512, fpu_reduction=r["fpu"], force_k=r["fk"])
    with Engine(concurrent_games=r["conc"], max_turns=50, max_batch_size=16, max_simulations=r["sims"]) as eng:
        eng.load_weights(r["arch"], 7, 7, r["sd"]())
        eng.set_eval_cache(CACHE)
        eng.selfplay_upload(pods, list(range(r["n"])))
        st = eng.selfplay_run_resident(cfg)
        summ, _ = eng.selfplay_download(r["n"], 50)
    nn = sum(summ[i].total_nn_evals for i in range(r["n"]))
    term = sum(summ[i].total_terminals for i in range(r["n"]))
    coll = sum(summ[i].total_collisions for i in range(r["n"]))
    npos = sum(summ[i].n_positions for i in range(r["n"]))
    sec = st.device_ms * 1e-3
    print(json.dumps({
        "evaluator": name, "eval_cache": CACHE, "cache_hits": int(st.cache_hits), "cache_misses": int(st.cache_misses), "concurrent_games": r["conc"], "games": r["n"], "simulations": r["sims"],
        "device_s": round(sec, 3), "positions": npos, "S_new_per_s": (nn + term) / sec, "nn_evals_per_s": nn / sec,
        "games_per_hour": r["n"] / sec * 3600, "collision_fraction": coll / max(nn + term + coll, 1),
        "evaluator_rows_per_s": (int(st.cache_misses) if CACHE else nn) / sec,
        "leaf_eval_tflops": (int(st.cache_misses) if CACHE else nn) / sec * FLOPS[name] / 1e12, "kernel_launches": int(st.kernel_launches),
        "path_nodes_per_sim": st.path_nodes / max(nn + term, 1)}), flush=True)
