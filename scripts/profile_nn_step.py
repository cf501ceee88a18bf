"""Config-3 style NN self-play for ncu captures of nn_step_kernel (use --launch-skip to land mid-run)."""
import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine, search_cfg
from alpharat_b200.games import make_games, pods_array
from nn_ref import make_mlp_state_dict
n = int(sys.argv[1]); conc = int(sys.argv[2]); mt = int(sys.argv[3])
specs = make_games(n, width=7, height=7, cheese_count=10, max_turns=mt)
cfg = search_cfg(simulations=1897, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103)
with Engine(concurrent_games=conc, max_turns=mt, max_batch_size=16, max_simulations=1897) as eng:
    eng.load_weights(N.AR_ARCH_MLP, 7, 7, make_mlp_state_dict(0, 349))
    eng.set_eval_cache(int(os.environ.get("AR_EVAL_CACHE", "0")))
    eng.selfplay_upload(pods_array(specs), list(range(n)))
    st = eng.selfplay_run_resident(cfg)
print(f"device_ms={st.device_ms:.1f} launches={st.kernel_launches} hits={st.cache_hits} misses={st.cache_misses}")
