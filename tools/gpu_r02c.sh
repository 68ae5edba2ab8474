#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
./tools/fp64_lat > gpurun_out/${tag}_fp64_lat.txt 2>&1; cat gpurun_out/${tag}_fp64_lat.txt
timeout 600 python -m pytest tests/test_gpu_sparse_align.py tests/test_gpu_depth_filter.py -m gpu -q 2>&1 | tail -5
( time timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | tail -3
tail -c 1500 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "align_ms", d["roofline"]["kernel_ms"], "parity", d["parity_sampled"])
    for k, v in d["paths"].items():
        print(k, "value", v["value"], v["unit"], "ms", v["ms_per_step"], "kernel_ms", v["kernel_ms"], "e2e", v["e2e"]["value"], "frac", v["roofline"]["frac"],
              "cpu", (v.get("cpu_baseline") or {}).get("value"), v["parity_sampled"]["status"])
        if "filter_only" in v: print("   filter_only", v["filter_only"]["kernel_ms"], v["filter_only"]["roofline"]["frac"])
    print(d["cpu_baseline"]); print(d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
