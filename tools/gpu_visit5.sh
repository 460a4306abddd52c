#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "pruning or root_primitive or flat_and or view_cache or odd_sizes or batch or full_size" 2>&1 | tail -2
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-230
for n in 1 8; do CSG_B200_LIB=$PWD/cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so python tools/gpu_prune_probe.py $n 2>&1 | tail -11; done
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "rank0"
for t in 128 256 512; do echo "flat threads $t"; CSG_B200_FLAT_THREADS=$t timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "rank0" | cut -c1-60; done
