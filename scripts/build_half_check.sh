#!/bin/bash
# Builds alpharat_b200/libalpharat_cuda_check.so: the same library with -DAR_HALF_CHECK (mcts_half.cuh traps when a
# collective of the two-trees-per-warp engine is reached by anything but one whole half or both halves).
# Use: ALPHARAT_CUDA_LIB=alpharat_b200/libalpharat_cuda_check.so python -m pytest tests/test_gpu_parity_uniform.py -m gpu -k half
set -e
cd "$(dirname "$0")/.."
C=alpharat_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -fmad=false ${AR_DEFS:--DAR_HALF_CHECK} -c $C/engine.cu -o /tmp/engine_check.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ${AR_OUT:-alpharat_b200/libalpharat_cuda_check.so} /tmp/engine_check.o $C/nn_kernels.o $C/nn_symmetric.o $C/nn_cnn.o -lcuda
echo ${AR_OUT:-alpharat_b200/libalpharat_cuda_check.so}
