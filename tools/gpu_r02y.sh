#!/bin/bash
# A/B of alternative builds through the front-end chain leg: usage bash tools/gpu_r02y.sh <tag> lib1.so lib2.so ...
tag=$1; shift
mkdir -p gpurun_out
L=$PWD/svo_pro_universal_b200
for lib in "$@"; do
  SVO_CUDA_LIB=$L/$lib timeout 300 python -m pytest tests/test_gpu_frontend_chain.py tests/test_gpu_sparse_align.py -m gpu -q -x 2>&1 | tail -1
  SVO_CUDA_LIB=$L/$lib timeout 600 python bench.py --steps 5 --warmup 3 --paths frontend_8192 > gpurun_out/${tag}_bench_$lib.json 2> gpurun_out/${tag}_bench_$lib.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_$lib.json").read().strip().splitlines()[-1])
f = d["paths"]["frontend_8192"]; print("$lib", "headline align", d["roofline"]["kernel_ms"], "chain", f["ms_per_step"], f["kernel_ms"], f["parity_sampled"]["status"])
PY
done | tee gpurun_out/${tag}.log
