#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ref_swap.py -m gpu -q 2>&1 | tail -40
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ref_swap.py > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log
