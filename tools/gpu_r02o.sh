#!/bin/bash
# A/B of the alignment kernel variants: threads per pair x patch-cache type
tag=${1:-r02o}
mkdir -p gpurun_out
L=svo_pro_universal_b200
for lib in libsvo_cuda_t192d.so libsvo_cuda_r01.so libsvo_cuda_t192f.so libsvo_cuda_t128f.so libsvo_cuda.so libsvo_cuda_t96d.so; do
  for B in 4096 592 1; do
    SVO_CUDA_LIB=$PWD/$L/$lib PROF_PAIRS=$B timeout 300 python tools/exp_align.py 2>&1 | tail -1
  done
done | tee gpurun_out/${tag}_align.log
python - <<'PY' | tee -a gpurun_out/${tag}_align.log
import numpy as np, glob, os
base = np.load("gpurun_out/align_res_libsvo_cuda_t192d_B4096.npz")
for f in sorted(glob.glob("gpurun_out/align_res_*_B4096.npz")):
    r = np.load(f)
    dq = 2 * np.arccos(np.clip(np.abs((r["T"][:, :4] * base["T"][:, :4]).sum(1)), 0, 1))
    dt = np.abs(r["T"][:, 4:] - base["T"][:, 4:]).max(1)
    print(os.path.basename(f), "iters equal:", bool(np.array_equal(r["iters"], base["iters"])), "n equal:", bool(np.array_equal(r["n"], base["n"])),
          "max |dq| %.3e max |dt| %.3e" % (dq.max(), dt.max()), "iters differ in", int((r["iters"] != base["iters"]).any(1).sum()), "pairs")
PY
