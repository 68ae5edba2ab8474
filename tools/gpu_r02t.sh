#!/bin/bash
tag=${1:-r02t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_align.py tests/test_gpu_frontend_chain.py tests/test_gpu_reference_frontend.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --paths frontend_8192 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["roofline"]["kernel_ms"])
f = d["paths"]["frontend_8192"]; print(f["value"], f["ms_per_step"], f["kernel_ms"], f["parity_sampled"]["status"])
PY
