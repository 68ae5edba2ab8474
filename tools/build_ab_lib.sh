#!/bin/bash
# A/B helper: build libsvo_cuda of another git revision next to the current one (svo_pro_universal_b200/libsvo_cuda_<tag>.so, git-ignored,
# travels to the GPU box). Select it at run time with SVO_CUDA_LIB=... (capi.lib()).   usage: tools/build_ab_lib.sh <rev> <tag>
set -e
rev=${1:-HEAD}; tag=${2:-ab}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
git -C "$root" archive "$rev" svo_pro_universal_b200/csrc include | tar -x -C "$tmp"
make -C "$tmp/svo_pro_universal_b200/csrc" -j8 -s all
cp "$tmp/svo_pro_universal_b200/libsvo_cuda.so" "$root/svo_pro_universal_b200/libsvo_cuda_$tag.so"
rm -rf "$tmp"
echo "built svo_pro_universal_b200/libsvo_cuda_$tag.so from $rev"
