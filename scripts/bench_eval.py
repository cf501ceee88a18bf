"""Leaf-evaluator kernels alone (run under `ncu --metrics gpu__time_duration.sum` for per-launch
device times): ar_nn_forward on n synthetic positions for each architecture."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from alpharat_b200 import _native as N
from alpharat_b200.engine import Engine
from alpharat_b200.games import pods_array
from nn_ref import make_cnn_state_dict, make_mlp_state_dict, make_symmetric_state_dict, random_positions

import os
sizes = [int(a) for a in sys.argv[1:]] or [128, 148 * 128, 148 * 128 * 8]
only = os.environ.get('EVAL_ARCH')
base = random_positions(256, 7, 7, seed=9)
with Engine(concurrent_games=4, max_turns=120) as eng:
    for name, arch, sd in (("mlp", N.AR_ARCH_MLP, make_mlp_state_dict(0, 349)),
                           ("symmetric", N.AR_ARCH_SYMMETRIC, make_symmetric_state_dict(2, 7, 7)),
                           ("cnn", N.AR_ARCH_CNN, make_cnn_state_dict(3, ("res", "res", "gpool")))):
        if only and name != only:
            continue
        eng.load_weights(arch, 7, 7, sd)
        for n in sizes:
            pods = pods_array([base[i % 256] for i in range(n)])
            for _ in range(2):
                out = eng.nn_forward(pods)
            print(name, n, float(out[2].mean()), flush=True)
