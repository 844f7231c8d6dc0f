#!/bin/bash
# tools/build_variant.sh <name> [extra nvcc flags, e.g. -DA3D_WAVE_STEPS=32]
# builds variant_<name>.so in the repo root from csrc/*.cu with the flags of __graft_entry__.build_cuda
# (for tools/bench_variants.sh: A/B of kernel tuning macros on one box)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
  --threads 0 "$@" -o variant_${name}.so mav_active_3d_planning_b200/csrc/*.cu
