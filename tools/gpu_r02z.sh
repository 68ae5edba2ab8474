#!/bin/bash
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reprojector.py tests/test_gpu_host_facade.py tests/test_gpu_frontend_chain.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --paths frontend_8192 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
f = d["paths"]["frontend_8192"]; print("headline align", d["roofline"]["kernel_ms"], "chain", f["ms_per_step"], f["kernel_ms"], f["parity_sampled"]["status"], f["mean_matches_per_frame"])
PY
timeout 300 python tools/bench_kernels.py 2>/dev/null | grep -i "reproj" | cut -c1-400
