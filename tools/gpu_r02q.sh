#!/bin/bash
# ping-pong alignment kernel: parity first (short timeouts: a barrier mismatch would hang), then timing
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse_align.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${tag}_tests.log
for np_ in 1 2; do
  for B in 4096 592 148 1; do
    SVO_ALIGN_PAIRS_PER_CTA=$np_ PROF_PAIRS=$B timeout 120 python tools/exp_align.py 2>&1 | tail -1 | sed "s/^/pairs_per_cta=$np_ /"
  done
done | tee gpurun_out/${tag}_align.log
for B in 4096 592 148 1; do PROF_PAIRS=$B timeout 120 python tools/exp_align.py 2>&1 | tail -1 | sed "s/^/auto /"; done | tee -a gpurun_out/${tag}_align.log
timeout 900 python -m pytest tests/test_gpu_reference_frontend.py tests/test_gpu_frontend_chain.py tests/test_gpu_ref_swap.py tests/test_gpu_host_facade.py tests/test_gpu_multi_context.py -m gpu -q -x 2>&1 | tail -8 | tee -a gpurun_out/${tag}_tests.log
