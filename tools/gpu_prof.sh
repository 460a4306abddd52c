#!/bin/bash
# ncu full capture (with source counters) of the frame and pruning kernels of the current build: gpurun_out/prof.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/prof.ncu-rep
