#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over every kernel of libsvo_cuda.so, driven by tools/prof_paths.py at
# small sizes (PROF_SMALL=1: a few CTAs per kernel, every code path of the hot-path entry points). Summaries -> gpurun_out/<tag>_sanitize_*.log
# usage (GPU box): [SAN_TOOLS="memcheck racecheck"] [PROF_UNTIL=detect] bash tools/sanitize.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no --padding 32"
  [ $tool = racecheck ] && extra="--racecheck-report all"
  PROF_SMALL=1 timeout 2400 $SAN --tool $tool $extra --print-limit 30 --error-exitcode 3 python tools/prof_paths.py > gpurun_out/${tag}_sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/${tag}_sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit|Error|hazard" gpurun_out/${tag}_sanitize_$tool.log | sort | uniq -c | head -12
done
