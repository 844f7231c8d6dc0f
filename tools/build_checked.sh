#!/bin/bash
# Builds variant_check.so: the CUDA library with -DA3D_BOUNDS_CHECK (device-side checks of every computed
# index into the map mirror and the kernel scratch; a violation prints the expression and traps).
# Run the GPU suite against it with  A3D_LIB_PATH=$PWD/variant_check.so python -m pytest tests -m gpu
set -e
cd "$(dirname "$0")/.."
${NVCC:-/usr/local/cuda/bin/nvcc} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC -shared --threads 0 -DA3D_BOUNDS_CHECK -o variant_check.so mav_active_3d_planning_b200/csrc/*.cu
echo "built $PWD/variant_check.so"
