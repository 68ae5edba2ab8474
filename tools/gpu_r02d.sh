#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_align.py tests/test_gpu_reference_frontend.py tests/test_gpu_frontend_chain.py tests/test_gpu_host_facade.py -m gpu -q 2>&1 | tail -15
P=svo_pro_universal_b200
for lib in libsvo_cuda_r01.so libsvo_cuda.so; do
  SVO_CUDA_LIB=$PWD/$P/$lib timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
done
for b in 148 4096; do
  SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=$b timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
done
cat gpurun_out/${tag}_align.log
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
    print(d["cpu_baseline"]); print(d["clocks"]); print(d["latency"])
except Exception as e:
    print("bench parse failed", e)
PY
