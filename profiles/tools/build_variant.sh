#!/bin/bash
# Builds a variant of the library for same-box A/B runs (profiles/tools/ab_slide.py):
#   profiles/tools/build_variant.sh OUT.so [-DFLAG ...]
# e.g. profiles/tools/build_variant.sh scratch/libd3d_L3.so -DD3D_PIPE_L=3 -DD3D_PIPE_PROF
set -e
HERE="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$1"; shift
nvcc -shared -Xcompiler -fPIC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     "$@" -o "$OUT" "$HERE/deconv3d_b200/csrc/d3d_api.cu"
echo "built $OUT $*"
