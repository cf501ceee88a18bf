import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import load_oracle, oracle_selfplay
from test_gpu_parity_uniform import compare_selfplay
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
n=int(sys.argv[1]); pool=int(sys.argv[2])
specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=50)
pods = pods_array(specs)
cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
seeds=list(range(n))
with Engine(concurrent_games=n, max_turns=50, max_batch_size=16, max_simulations=1897, pool_nodes=pool) as eng:
    gpu = eng.selfplay(pods, cfg, seeds)
s,p,stride,st = gpu
print('gpu max nodes', max(p[i*stride+t].search.node_count for i in range(n) for t in range(s[i].n_positions)), 'device_ms', st.device_ms)
cpu = oracle_selfplay(load_oracle(), pods, cfg, seeds)
compare_selfplay(gpu, cpu, n)
print('parity ok', n)
