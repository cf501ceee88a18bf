#!/bin/bash
# usage: r2_ncu.sh TAG WARPS_PER_SM MAX_MOVES  — one ncu capture of the tt kernel at full occupancy,
# real 50-turn games cut after MAX_MOVES moves (AR_TT_MAX_MOVES profiling knob) to bound the kernel time
TAG=$1; W=$2; MM=$3
conc=$((W*148*32))
mkdir -p gpurun_out
AR_TT_ARENA_GB=40 AR_TT_MAX_MOVES=$MM AR_TT_WARPS_PER_SM=$W timeout 60 python scripts/profile_uniform.py $conc $conc 50 > gpurun_out/${TAG}_plain.log 2>&1
AR_TT_ARENA_GB=40 AR_TT_MAX_MOVES=$MM AR_TT_WARPS_PER_SM=$W timeout 900 ncu --set full --clock-control none --import-source on -k regex:selfplay_tt -c 1 -o gpurun_out/${TAG}_prof -f \
    python scripts/profile_uniform.py $conc $conc 50 > gpurun_out/${TAG}_ncu.log 2>&1
