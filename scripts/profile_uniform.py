"""Small fixed workload for ncu: 7x7 tuned config, short games (bounded kernel time)."""
import os
import sys
sys.path.insert(0, '.')
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
n = int(sys.argv[1]); conc = int(sys.argv[2]); mt = int(sys.argv[3])
specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=mt)
pods = pods_array(specs)
cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
eng = Engine(concurrent_games=conc, max_turns=mt, max_batch_size=16, max_simulations=1897, pool_nodes=int(sys.argv[4]) if len(sys.argv) > 4 else 0,
             tree_engine=os.environ.get("AR_TREE_ENGINE", "warp"))
eng.selfplay_upload(pods, list(range(n)))
st = eng.selfplay_run_resident(cfg)
summ, pos = eng.selfplay_download(n, mt)
npos = sum(summ[i].n_positions for i in range(n))
print(f"n={n} conc={conc} mt={mt} device_ms={st.device_ms:.1f} positions={npos} S_new/s={npos*1897/st.device_ms*1e3:.3e} path_nodes={st.path_nodes} new_nodes={st.new_nodes} algo_bytes={288*st.path_nodes+240*st.new_nodes}")
