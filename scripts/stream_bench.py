"""Continuous-feed probe: K batches of G games through NBUF stream buffers (kernel-only: inputs uploaded by
ar_stream_submit, no record download), device time from the first launch to the last.  usage: stream_bench.py G K [CONC] [NBUF]"""
import ctypes as C
import os
import sys, time
sys.path.insert(0, '.')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
G = int(sys.argv[1]); K = int(sys.argv[2]); conc = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
nbuf = int(sys.argv[4]) if len(sys.argv) > 4 else 3
base = min(G, 32768)
pods0 = pods_array(make_games(base, width=7, height=7, cheese_count=10, max_turns=50))
pods = (N.GamePod * G)()
sz = C.sizeof(N.GamePod)
for off in range(0, G, base):
    C.memmove(C.byref(pods, off * sz), pods0, min(base, G - off) * sz)
cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
eng = Engine(concurrent_games=conc, max_turns=50, max_batch_size=16, max_simulations=1897,
             tree_engine=os.environ.get("AR_TREE_ENGINE", "warp"))
# blocking reference point
eng.selfplay_upload(pods, list(range(G)))
st = eng.selfplay_run_resident(cfg)
summ, _ = eng.selfplay_download(G, 50)
npos = sum(summ[i].n_positions for i in range(G))
print(f"blocking: G={G} device_ms={st.device_ms:.1f} S_new/s={npos * 1897 / st.device_ms * 1e3:.3e}", flush=True)
eng.stream_open(nbuf, G, 50)
sims = 0
t0 = time.perf_counter()
done = []  # (wall time at which the batch was complete, its simulations)
for i in range(K + nbuf):
    b = i % nbuf
    if i >= nbuf:
        s = eng.stream_wait(b)
        sims += s.total_nn_evals + s.total_terminals
        if os.environ.get("ALPHARAT_CUDA_LIB", "").endswith("_idle.so"):  # -DAR_HALF_IDLE build: cycles / 1024
            print(f"batch {i - nbuf}: of {s.new_nodes} half-kcycles: waiting at the loop top {s.path_nodes / max(s.new_nodes, 1):.3f}, "
                  f"after the descents {s.total_nn_evals / max(s.new_nodes, 1):.3f}, after the backups {s.total_terminals / max(s.new_nodes, 1):.3f}", flush=True)
        done.append((time.perf_counter() - t0, s.total_nn_evals + s.total_terminals))
    if i < K:
        eng.stream_submit(b, pods, cfg, [i * G + j for j in range(G)])
wall = time.perf_counter() - t0
print(f"stream: G={G} K={K} nbuf={nbuf} wall_s={wall:.2f} sims={sims} S_new/s(wall)={sims / wall:.3e}", flush=True)
if K >= 8:  # steady state: completions of the middle batches (the first ones fill the machine, the last ones drain it)
    lo, hi = 2, K - 3
    mid = sum(x[1] for x in done[lo + 1:hi + 1])
    print(f"steady: batches {lo + 1}..{hi}: {mid / (done[hi][0] - done[lo][0]):.3e} S_new/s", flush=True)
eng.stream_close()
