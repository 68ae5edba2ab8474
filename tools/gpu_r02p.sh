#!/bin/bash
# stagger sweep of the alignment kernel (start delay per arrival slot on the SM)
tag=${1:-r02p}
mkdir -p gpurun_out
L=svo_pro_universal_b200
for lib in libsvo_cuda_t192f.so libsvo_cuda.so; do
  for st in 0 1000 2000 3000 4000 6000; do
    for mask in 3 7; do
      SVO_ALIGN_STAGGER=$st SVO_ALIGN_STAGGER_MASK=$mask SVO_CUDA_LIB=$PWD/$L/$lib PROF_PAIRS=4096 timeout 300 python tools/exp_align.py 2>&1 | tail -1 | sed "s/^/stagger=$st mask=$mask /"
      [ $st = 0 ] && break
    done
  done
done | tee gpurun_out/${tag}_align.log
