#!/bin/bash
# round 2, visit J: per-warp timeline of the frame kernel and phase timeline of the pruning kernel (probe build), shard 0 of 8 and one GPU
export CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so
for N in 8 1; do
  timeout 300 python tools/gpu_frame_probe.py $N 2>&1 | tail -22
done
timeout 300 python tools/gpu_prune_probe.py 8 2>&1 | tail -14
