#!/bin/bash
# usage: r2_ncu_src.sh TAG KERNEL_VARIANT(0 lane-bound,1 block-sorted) — source-level capture, 1 block/SM, first move only
TAG=$1; V=$2
conc=$((148*128))
mkdir -p gpurun_out
AR_TT_KERNEL=$V AR_TT_ARENA_GB=12 AR_TT_MAX_MOVES=${MM:-1} timeout 600 ncu --set full --clock-control none --import-source on -k regex:selfplay_t -c 1 -o gpurun_out/${TAG}_prof -f \
    python scripts/profile_uniform.py $conc $conc 50 > gpurun_out/${TAG}_ncu.log 2>&1
echo "exit $?" >> gpurun_out/${TAG}_ncu.log
