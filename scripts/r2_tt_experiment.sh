#!/bin/bash
# How the thread-per-tree numbers of profiles/r2_summary.md section 3 were taken (one B200).
#   AR_TREE_ENGINE=thread      quick_bench.py / profile_uniform.py create the engine with tree_engine="thread"
#   AR_TT_KERNEL=0|1           0 lane-bound kernel (default), 1 block-sorted kernel
#   AR_TT_WARPS_PER_SM=8|12|16|20   register budget of the lane-bound kernel
#   AR_TT_MAX_MOVES=k          profiling knob: stop every game after k moves (bounds the kernel time under ncu)
#   AR_TT_ARENA_GB=g           profiling knob: small node arena (keeps ncu's save / restore of device memory cheap)
TAG=${1:-tt}
mkdir -p gpurun_out
export AR_TREE_ENGINE=thread
# occupancy sweep, 6 games per resident tree (r2_tt_occupancy_sweep.log)
for w in 8 12 16 20; do
  conc=$((w*148*32))
  AR_TT_WARPS_PER_SM=$w timeout 400 python scripts/quick_bench.py $((conc*6)) $conc 1 >> gpurun_out/${TAG}_sweep.log 2>&1
done
# block-sorted kernel at 3 blocks of 128 per SM
conc=$((3*148*128))
AR_TT_KERNEL=1 timeout 400 python scripts/quick_bench.py $((conc*6)) $conc 1 >> gpurun_out/${TAG}_sweep.log 2>&1
# counters on the first four moves (r2_tt_lanebound_w12_counters.csv, r2_tt_blocksorted_counters.csv)
AR_TT_KERNEL=0 bash scripts/r2_ncu_light.sh ${TAG}_lane 12 4
AR_TT_KERNEL=1 bash scripts/r2_ncu_light.sh ${TAG}_sorted 12 4
