#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_uniform.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for w in 8 12 16 20; do
  conc=$((w*148*32))
  AR_TT_WARPS_PER_SM=$w timeout 400 python scripts/quick_bench.py $((conc*6)) $conc 1 >> gpurun_out/${TAG}_sweep.log 2>&1
done
