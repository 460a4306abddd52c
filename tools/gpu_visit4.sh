#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-230
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_launches_run.log 2>&1
grep "prune\|frame" gpurun_out/launches.csv | head -4 | cut -d, -f5,12-
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "rank0"
