#!/bin/bash
# One gpurun call: GPU parity tests, headline bench, per-path benches, ncu launch list and full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [tests|bench|paths|launches|ncu ...]
# Everything lands under gpurun_out/<tag>_*; summaries to keep are copied into profiles/ by hand afterwards.
tag=${1:-r01}; shift
what=${*:-tests bench paths launches ncu}
mkdir -p gpurun_out
has() { [[ " $what " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
  tail -5 gpurun_out/${tag}_tests.log
fi
if has smoke; then
  timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
fi
if has bench; then
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 3000 gpurun_out/${tag}_bench.json
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench_ref.json
fi
if has paths; then
  timeout 600 python tools/bench_kernels.py > gpurun_out/${tag}_paths.jsonl 2> gpurun_out/${tag}_paths.err; cat gpurun_out/${tag}_paths.jsonl | cut -c1-600
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --batch 1184 > gpurun_out/${tag}_launches_bench.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_paths.csv \
      python tools/prof_paths.py > gpurun_out/${tag}_launches_paths.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_frontend.csv \
      python tools/bench_frontend.py --pairs 8192 --steps 1 --warmup 1 > gpurun_out/${tag}_launches_frontend.log 2>&1
fi
if has ncu; then
  for k in ${NCU_KERNELS:-sparse_align_kernel pyr_down fast_level match_kernel..int.0 match_kernel..int.1 seed_step_kernel seed_match_kernel filter_seq_kernel scan_epipolar_kernel reproj_match reproj_sort pose_optimize_kernel edgelet_score edgelet_decode optimize_points_kernel stereo_commit corner_scatter_kernel}; do
    kf=$(echo "$k" | tr -c 'A-Za-z0-9_\n' '_')
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -s 1 -c 1 -f -o gpurun_out/${tag}_ncu_$kf \
        python tools/prof_paths.py > gpurun_out/${tag}_ncu_$kf.log 2>&1
  done
fi
ls -la gpurun_out | tail -30
