#!/bin/bash
# A/B of alternative builds of libsvo_cuda (make -C svo_pro_universal_b200/csrc variant ...) through bench.py legs on the GPU box:
#   bash tools/ab_bench.sh <tag> <comma-separated bench paths> libsvo_cuda.so libsvo_cuda_<variant>.so ...
tag=$1; paths=$2; shift; shift
mkdir -p gpurun_out
L=$PWD/svo_pro_universal_b200
for lib in "$@"; do
  SVO_CUDA_LIB=$L/$lib timeout 600 python bench.py --steps 5 --warmup 3 --paths $paths > gpurun_out/${tag}_bench_$lib.json 2> gpurun_out/${tag}_bench_$lib.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_$lib.json").read().strip().splitlines()[-1])
print("$lib", {k: (round(v["ms_per_step"], 4), v.get("kernel_ms") if isinstance(v.get("kernel_ms"), dict) else None, v["parity_sampled"]["status"]) for k, v in d["paths"].items()})
PY
done | tee gpurun_out/${tag}.log
