"""Steady-state probe of the uniform self-play kernel: N games (32768 distinct layouts tiled, distinct seeds)
through CONC resident trees; prints device-timed S_new/s.  usage: quick_bench.py N CONC [REPEATS]"""
import ctypes as C
import os
import sys, time
sys.path.insert(0,'.')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
n=int(sys.argv[1]); conc=int(sys.argv[2])
base = min(n, 32768)
pods0 = pods_array(make_games(base, width=7, height=7, cheese_count=10, max_turns=50))
pods = (N.GamePod * n)()
sz = C.sizeof(N.GamePod)
for off in range(0, n, base):
    m = min(base, n - off)
    C.memmove(C.byref(pods, off * sz), pods0, m * sz)
cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
eng = Engine(concurrent_games=conc, max_turns=50, max_batch_size=16, max_simulations=1897,
             tree_engine=os.environ.get("AR_TREE_ENGINE", "warp"))  # AR_TREE_ENGINE=thread: the thread-per-tree engine
eng.selfplay_upload(pods, list(range(n)))
for it in range(int(sys.argv[3]) if len(sys.argv)>3 else 1):
    st = eng.selfplay_run_resident(cfg)
    snew = st.total_nn_evals + st.total_terminals if False else None
    summ, pos = eng.selfplay_download(n, 50)
    npos = sum(summ[i].n_positions for i in range(n))
    snew = npos*1897
    print(f"n={n} conc={conc} device_ms={st.device_ms:.1f} positions={npos} S_new/s={snew/st.device_ms*1e3:.3e} games/h={n/st.device_ms*3.6e6:.3e} path_nodes/sim={st.path_nodes/snew:.2f} new/sim={st.new_nodes/snew:.3f} GB/s={(288*st.path_nodes+240*st.new_nodes)/st.device_ms/1e6:.1f}", flush=True)
