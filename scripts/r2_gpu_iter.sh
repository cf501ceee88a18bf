#!/bin/bash
# round 2 iteration helper: quick uniform parity, concurrency sweep, optional ncu capture
# usage: r2_gpu_iter.sh TAG [ncu]
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_uniform.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
for conc in ${CONCS:-8192 16384 32768}; do
  timeout 300 python scripts/quick_bench.py ${NGAMES:-131072} $conc 1 >> gpurun_out/${TAG}_sweep.log 2>&1
done
if [ "$2" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:selfplay_tt -c 1 -o gpurun_out/${TAG}_prof -f \
    python scripts/profile_uniform.py 4096 4096 12 > gpurun_out/${TAG}_ncu.log 2>&1
fi
