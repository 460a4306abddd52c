#!/bin/bash
# probe timelines (instrumented build) of the current kernels: shard 0 of 8 and one GPU
mkdir -p gpurun_out
export CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so
timeout 300 python tools/gpu_frame_probe.py 8 2>&1 | tail -20 | tee gpurun_out/frame_probe_8.txt
timeout 300 python tools/gpu_frame_probe.py 1 2>&1 | tail -20 | tee gpurun_out/frame_probe_1.txt
timeout 300 python tools/gpu_prune_probe.py 8 2>&1 | tail -12 | tee gpurun_out/prune_probe_8.txt
timeout 300 python tools/gpu_prune_probe.py 1 2>&1 | tail -12 | tee gpurun_out/prune_probe_1.txt
