#!/bin/bash
# multi-GPU e2e A/B: page-locked vs write-combined upload buffers (usage: gpurun --gpus N -- bash tools/gpu_r02l.sh <tag> N)
tag=${1:-r02l}; n=${2:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name" >> gpurun_out/${tag}_topo.txt; free -g | head -2 >> gpurun_out/${tag}_topo.txt
for pin in torch wc torch wc; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --paths '' --pinned $pin \
     > gpurun_out/${tag}_bench_n${n}_$pin.json 2> gpurun_out/${tag}_bench_n${n}_$pin.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n${n}_$pin.json").read().strip().splitlines()[-1])
    print("$pin N=$n", "value", d["value"], "e2e", d["e2e"]["value"], "GB/s per rank", d["e2e"]["h2d_GBs_per_rank"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/${tag}_bench_n${n}_$pin.err").read()[-1500:])
PY
done
cat gpurun_out/${tag}_topo.txt | head -30
