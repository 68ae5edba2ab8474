#!/bin/bash
tag=${1:-r03a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matcher.py tests/test_gpu_depth_filter.py tests/test_gpu_reference_frontend.py tests/test_gpu_stereo_triangulation.py tests/test_gpu_reprojector.py tests/test_gpu_frontend_chain.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/${tag}_tests.log
timeout 600 python tools/exp_match_order.py 2>&1 | tail -8 | tee gpurun_out/${tag}_order.log
timeout 600 python bench.py --steps 5 --warmup 3 --paths match_512k,seeds_50k_x64,frontend_8192 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
for k, v in d["paths"].items(): print(k, v["ms_per_step"], v.get("kernel_ms"), v["parity_sampled"]["status"])
PY
