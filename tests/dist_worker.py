"""Worker of tests/test_gpu_distributed.py: run under torchrun, one rank per GPU.  Every rank plays its shard of
one seeded run through `cuda_self_play_distributed`; rank 0 writes the NCCL-gathered records to OUT."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch
import torch.distributed as dist

from alpharat_b200.parallel import cuda_self_play_distributed

out = sys.argv[1]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
KW = dict(width=7, height=5, cheese_count=6, max_turns=24, maze_type="classic", positions="random",
          simulations=200, batch_size=16, c_puct=0.512, fpu_reduction=0.459, force_k=0.103, output_dir=None)
total, payloads = cuda_self_play_distributed(num_games=int(sys.argv[2]), seed=int(sys.argv[3]), gather_records=True,
                                             concurrent_games=16, **KW)
if dist.get_rank() == 0:
    np.savez(out, summaries=np.concatenate([p[0] for p in payloads]), records=np.concatenate([p[1] for p in payloads]),
             per_rank_games=np.array([len(p[0]) for p in payloads]),
             total_games=total["total_games"], total_positions=total["total_positions"],
             total_simulations=total["total_simulations"])
dist.barrier()
dist.destroy_process_group()
