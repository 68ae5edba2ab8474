#!/bin/bash
# new API-surface entry points: corner lists, stand-alone epipolar scan, Matcher members through the facade and the swap
tag=${1:-r02n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_detect.py tests/test_gpu_matcher.py tests/test_gpu_reference_frontend.py tests/test_gpu_host_facade.py tests/test_gpu_ref_swap.py tests/test_gpu_depth_filter.py tests/test_gpu_stereo_triangulation.py -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1
tail -30 gpurun_out/${tag}_tests.log
