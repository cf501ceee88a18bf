#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_uniform.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
conc=$((3*148*128))
timeout 400 python scripts/quick_bench.py $((conc*6)) $conc 1 >> gpurun_out/${TAG}_sweep.log 2>&1
AR_TT_MAX_MOVES=4 timeout 100 python scripts/profile_uniform.py $conc $conc 50 >> gpurun_out/${TAG}_sweep.log 2>&1
