#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-230
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "rank0"
CSG_B200_LIB=$PWD/cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so python tools/gpu_frame_probe.py 8 | head -14
