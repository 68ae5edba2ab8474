#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
P=svo_pro_universal_b200
timeout 600 python -m pytest tests/test_gpu_sparse_align.py tests/test_gpu_reference_frontend.py -m gpu -q 2>&1 | tail -4
SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=148 timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=4096 timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
for lib in libsvo_cuda_r01.so libsvo_cuda.so; do
  SVO_CUDA_LIB=$PWD/$P/$lib timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
done
cat gpurun_out/${tag}_align.log
