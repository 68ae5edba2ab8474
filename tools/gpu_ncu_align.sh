#!/bin/bash
# ncu --set full capture of sparse_align_kernel at PROF_PAIRS pairs (default 148: one CTA per SM, the kernel's latency floor)
tag=${1:-r02h}; pairs=${2:-148}
mkdir -p gpurun_out
PROF_PAIRS=$pairs timeout 900 ncu --set full --clock-control none --import-source on -k regex:sparse_align_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_ncu_align_$pairs \
   python tools/exp_align.py > gpurun_out/${tag}_ncu_align_$pairs.log 2>&1
tail -3 gpurun_out/${tag}_ncu_align_$pairs.log
