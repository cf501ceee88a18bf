"""Per-tree latency of the warp engine as a function of load: the measurement behind (or against) the statement
"a tree is a serial chain of a few microseconds per simulation" (DESIGN.md section 5).

    python scripts/latency_probe.py

Plays 6 games per resident tree through T resident trees (T = 148 is one warp per SM: an unloaded machine; 4144 is
the resident capacity, 7 blocks of 4 warps per SM) and prints, per T: simulations/s, the time one tree needs per
simulation (T / throughput) and per node visit (path nodes are counted by the kernel)."""
import ctypes as C
import sys
sys.path.insert(0, '.')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array

cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
base = pods_array(make_games(8192, width=7, height=7, cheese_count=10, max_turns=50))
sz = C.sizeof(N.GamePod)
print("trees  warps/SM  sims/s      us/sim/tree  us/node-visit/tree  (6 games per tree, tail included)")
for trees in (148, 296, 592, 1184, 2368, 4096):
    n = trees * 6
    pods = (N.GamePod * n)()
    for off in range(0, n, 8192):
        C.memmove(C.byref(pods, off * sz), base, min(8192, n - off) * sz)
    with Engine(concurrent_games=trees, max_turns=50, max_batch_size=16, max_simulations=1897) as eng:
        eng.selfplay_upload(pods, list(range(n)))
        st = eng.selfplay_run_resident(cfg)
        summ, _ = eng.selfplay_download(n, 50)
    sims = sum(summ[i].total_nn_evals + summ[i].total_terminals for i in range(n))
    rate = sims / (st.device_ms * 1e-3)
    print(f"{trees:5d}  {trees / 148:7.1f}  {rate:.3e}  {trees / rate * 1e6:11.2f}  {trees / rate * 1e6 / (st.path_nodes / sims):18.3f}", flush=True)
