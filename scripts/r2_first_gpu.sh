#!/bin/bash
# round 2, first GPU contact of the thread-per-tree engine: parity, then a concurrency sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity_uniform.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/r2_pytest_uniform.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_uniform.log
for conc in 4096 8192 16384 32768 65536; do
  timeout 300 python scripts/quick_bench.py 32768 $conc 2 >> gpurun_out/r2_sweep.log 2>&1
done
timeout 300 python scripts/quick_bench.py 131072 32768 1 >> gpurun_out/r2_sweep.log 2>&1
