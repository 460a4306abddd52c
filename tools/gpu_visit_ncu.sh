#!/bin/bash
# bench + full ncu capture (source-level) of the two per-frame kernels
mkdir -p gpurun_out
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-230
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/prof.ncu-rep
