import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
from nn_ref import make_mlp_state_dict
n=int(sys.argv[1]); conc=int(sys.argv[2]); sims=int(sys.argv[3]) if len(sys.argv)>3 else 1897
specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=50)
pods = pods_array(specs)
cfg = search_cfg(simulations=sims, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
eng = Engine(concurrent_games=conc, max_turns=50, max_batch_size=16, max_simulations=sims)
eng.load_weights(N.AR_ARCH_MLP, 7, 7, make_mlp_state_dict(0, 349))
eng.selfplay_upload(pods, list(range(n)))
t=time.time()
st = eng.selfplay_run_resident(cfg)
wall=time.time()-t
summ, pos = eng.selfplay_download(n, 50)
npos = sum(summ[i].n_positions for i in range(n))
nn = sum(summ[i].total_nn_evals for i in range(n)); term=sum(summ[i].total_terminals for i in range(n))
snew = nn+term
print(f"NN n={n} conc={conc} sims={sims} device_ms={st.device_ms:.1f} wall={wall:.2f}s positions={npos} S_new/s={snew/st.device_ms*1e3:.3e} nn_evals/s={nn/st.device_ms*1e3:.3e} launches={st.kernel_launches} steps~{st.kernel_launches//2} games/h={n/st.device_ms*3.6e6:.3e}", flush=True)
