#!/bin/bash
# round 2 measurements kept under profiles/: bench lines (both arms), launch list, one ncu --set full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt; lscpu | grep "Model name" >> gpurun_out/r2_gpu.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
timeout 1500 python bench.py --steps ${STEPS:-6} --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench exit $?" >> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 3 --games-per-step 16384 --no-nn --no-cpu --config5-games 0 > gpurun_out/r2_bench_n1_16k_k20.json 2> gpurun_out/r2_bench_n1_16k_k20.err
timeout 600 python bench.py --steps 6 --warmup 3 --concurrent 4096 --no-nn --no-cpu --config5-games 0 > gpurun_out/r2_bench_n1_c4096.json 2> gpurun_out/r2_bench_n1_c4096.err
# launch list of a short bench command (after it exited 0 without ncu)
timeout 300 python bench.py --steps 2 --warmup 1 --games-per-step 8192 --no-nn --no-cpu --config5-games 0 > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --games-per-step 8192 --no-nn --no-cpu --config5-games 0 > gpurun_out/r2_ncu_launches.log 2>&1
# the dominant kernel, bounded launch of the same workload
timeout 120 python scripts/profile_uniform.py 4736 4736 8 > gpurun_out/r2_profile_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selfplay_uniform -c 1 -o gpurun_out/r2_prof_uniform -f \
  python scripts/profile_uniform.py 4736 4736 8 > gpurun_out/r2_ncu_uniform.log 2>&1
