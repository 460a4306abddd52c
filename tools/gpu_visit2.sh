#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "pruning or root_primitive or flat_and or view_cache or odd_sizes or batch" 2>&1 | tail -4
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-230
timeout 300 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/racecheck_full.log 2>&1; tail -3 gpurun_out/racecheck_full.log
timeout 300 python tools/gpu_stats.py 2>&1 | tail -2 | cut -c1-900
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_launches_run.log 2>&1
grep -c . gpurun_out/launches.csv
