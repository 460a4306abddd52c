#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-200
python tools/gpu_cfg5_modes.py
CSG_B200_PRUNE_FLAT=1 timeout 300 python -m pytest tests -x -q -m gpu -k "pruning or synthetic" 2>&1 | tail -2
