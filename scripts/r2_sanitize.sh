#!/bin/bash
# compute-sanitizer over the uniform thread-per-tree engine: multi-page trees, compaction, page reuse
TAG=$1
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  AR_TREE_ENGINE=thread AR_TT_ARENA_GB=2 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/profile_uniform.py 96 64 50 > gpurun_out/${TAG}_${tool}.log 2>&1
  echo "exit $?" >> gpurun_out/${TAG}_${tool}.log
done
