mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h9_pytest_all.log 2>&1; echo "rc $?" >> gpurun_out/h9_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h9_smoke.log 2>&1; echo "rc $?" >> gpurun_out/h9_smoke.log
/usr/bin/time -v timeout 1500 python bench.py --steps 6 --warmup 3 > gpurun_out/h9_bench_n1.json 2> gpurun_out/h9_bench_n1.err; echo "rc $?" >> gpurun_out/h9_bench_n1.err
