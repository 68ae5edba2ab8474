#!/bin/bash
# Round-2 first GPU call: parity tests on the new matcher / seed / filter kernels, then A/B timings (r01 build vs current).
tag=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -15 gpurun_out/${tag}_tests.log
P=svo_pro_universal_b200
for lib in libsvo_cuda_r01.so libsvo_cuda.so libsvo_cuda_timing.so; do
  SVO_CUDA_LIB=$PWD/$P/$lib timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
done
SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=148 timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=592 timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=1 timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
cat gpurun_out/${tag}_align.log
for lib in libsvo_cuda_r01.so libsvo_cuda.so; do
  SVO_CUDA_LIB=$PWD/$P/$lib timeout 900 python tools/bench_kernels.py > gpurun_out/${tag}_paths_$lib.jsonl 2> gpurun_out/${tag}_paths_$lib.err
  cut -c1-420 gpurun_out/${tag}_paths_$lib.jsonl
done
