#!/bin/bash
# pose optimizer occupancy A/B (launch-bounds min blocks 1 / 3 / 4 / 5) through the front-end chain leg + its own parity tests
tag=${1:-r02u}
mkdir -p gpurun_out
L=$PWD/svo_pro_universal_b200
timeout 900 python -m pytest tests/test_gpu_sparse_align.py -m gpu -q -x 2>&1 | tail -3
for lib in libsvo_cuda.so libsvo_cuda_po3.so libsvo_cuda_po4.so libsvo_cuda_po5.so; do
  SVO_CUDA_LIB=$L/$lib timeout 300 python -m pytest tests/test_gpu_pose_optimizer.py -m gpu -q -x 2>&1 | tail -1
  SVO_CUDA_LIB=$L/$lib timeout 600 python bench.py --steps 5 --warmup 3 --paths frontend_8192 > gpurun_out/${tag}_bench_$lib.json 2> gpurun_out/${tag}_bench_$lib.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_$lib.json").read().strip().splitlines()[-1])
f = d["paths"]["frontend_8192"]; print("$lib", f["ms_per_step"], f["kernel_ms"]["pose_optimize"], f["parity_sampled"]["status"])
PY
done | tee gpurun_out/${tag}.log
