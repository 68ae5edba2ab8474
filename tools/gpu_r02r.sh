#!/bin/bash
tag=${1:-r02r}
mkdir -p gpurun_out
L=$PWD/svo_pro_universal_b200
for np_ in 1 2; do
  for B in 4096 148; do
    SVO_CUDA_LIB=$L/libsvo_cuda_timing.so SVO_ALIGN_PAIRS_PER_CTA=$np_ PROF_PAIRS=$B timeout 120 python tools/exp_align.py 2>&1 | tail -3 | sed "s/^/np=$np_ /"
  done
done | tee gpurun_out/${tag}_align.log
