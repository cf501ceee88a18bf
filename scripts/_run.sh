mkdir -p gpurun_out
for i in 1 2; do
ALPHARAT_CUDA_LIB=alpharat_b200/libalpharat_cuda_prev.so AR_TREE_ENGINE=half timeout 200 python scripts/stream_bench.py 32768 10 9472 4 > gpurun_out/h12_stream_prev_$i.log 2>&1
AR_TREE_ENGINE=half timeout 200 python scripts/stream_bench.py 32768 10 9472 4 > gpurun_out/h12_stream_new_$i.log 2>&1
done
