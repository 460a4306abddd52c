#!/bin/bash
# Builds tuning variants of libcsg_b200.so into build/variants/ (git-ignored, shipped by gpurun).
set -e
cd "$(dirname "$0")/../cuda-csg-tree-raycasting_b200"
mkdir -p ../build/variants
for v in "$@"; do
  T=${v%%x*}; B=${v##*x}
  /usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC,-ffp-contract=off -DCSG_THREADS=$T -DCSG_MIN_BLOCKS=$B -Xptxas -v -shared \
     -o ../build/variants/libcsg_b200_${T}x${B}.so csrc/csg_render.cu csrc/csg_scene.cpp 2>&1 | grep -E "csg_frame_kernelILi0ELb1|registers|spill" | grep -A2 "ILi0ELb1" | grep -E "registers|spill" | tr '\n' ' '
  echo " <- ${T}x${B}"
done
