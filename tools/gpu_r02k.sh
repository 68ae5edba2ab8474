#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_facade.py tests/test_gpu_multi_context.py -m gpu -q 2>&1 | tail -8
for pin in torch wc; do
  timeout 600 python bench.py --steps 10 --warmup 3 --paths '' --pinned $pin > gpurun_out/${tag}_bench_$pin.json 2> gpurun_out/${tag}_bench_$pin.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_$pin.json").read().strip().splitlines()[-1])
print("$pin", "value", d["value"], "e2e", d["e2e"]["value"], "GB/s", d["e2e"]["h2d_GBs_per_rank"], "align_ms", d["roofline"]["kernel_ms"])
PY
done
