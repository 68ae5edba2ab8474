#!/bin/bash
# Round-2 second GPU call: the whole GPU test suite, sparse-align phase clocks, the new bench.py (all BASELINE configs).
tag=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
tail -25 gpurun_out/${tag}_tests.log
P=svo_pro_universal_b200
for b in 148 4096; do
  SVO_CUDA_LIB=$PWD/$P/libsvo_cuda_timing.so PROF_PAIRS=$b timeout 300 python tools/exp_align.py >> gpurun_out/${tag}_align.log 2>&1
done
cat gpurun_out/${tag}_align.log
( time timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | tail -3
tail -c 1500 gpurun_out/${tag}_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02b_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "align_ms", d["roofline"]["kernel_ms"], "parity", d["parity_sampled"])
    for k, v in d["paths"].items():
        print(k, "value", v["value"], v["unit"], "ms", v["ms_per_step"], "kernel_ms", v["kernel_ms"], "e2e", v["e2e"]["value"], "frac", v["roofline"]["frac"],
              "cpu", (v.get("cpu_baseline") or {}).get("value"), v["parity_sampled"]["status"])
    print(d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
