mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt; lscpu | grep "Model name" >> gpurun_out/r2_gpu.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference_n1.json 2> gpurun_out/r2f_bench_reference_n1.err
timeout 1500 python bench.py --steps 8 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench exit $?" >> gpurun_out/r2f_bench_n1.err
timeout 600 python bench.py --steps 20 --warmup 3 --games-per-step 16384 --no-nn --no-cpu --config5-games 0 > gpurun_out/r2f_bench_n1_16k_k20.json 2> gpurun_out/r2f_bench_n1_16k_k20.err
timeout 600 python bench.py --steps 20 --warmup 3 --games-per-step 16384 --tree-engine warp --no-nn --no-cpu --config5-games 0 > gpurun_out/r2f_bench_n1_16k_k20_warp.json 2> gpurun_out/r2f_bench_n1_16k_k20_warp.err
