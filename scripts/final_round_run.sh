#!/bin/bash
# One GPU session that regenerates every measurement quoted in profiles/r1_summary.md (run through gpurun).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1
python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference_n1.json 2> $O/final_bench_reference_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launches_eval.csv python scripts/bench_eval.py 151552 > $O/final_bench_eval.log 2>&1
for a in mlp symmetric cnn; do
  EVAL_ARCH=$a ncu --set full --import-source on --clock-control none -k regex:_forward_kernel -c 1 -o $O/final_prof_$a python scripts/bench_eval.py 151552 > $O/final_ncu_$a.log 2>&1
done
: > $O/final_nn_configs.jsonl
python scripts/bench_nn_configs.py mlp cnn 2>/dev/null | grep '^{' >> $O/final_nn_configs.jsonl
AR_EVAL_CACHE=4096 python scripts/bench_nn_configs.py mlp symmetric cnn 2>/dev/null | grep '^{' >> $O/final_nn_configs.jsonl
tail -2 $O/final_pytest_gpu.log; cat $O/final_smoke.log | tail -1; head -c 300 $O/final_bench_n1.json; echo; wc -l $O/final_nn_configs.jsonl
