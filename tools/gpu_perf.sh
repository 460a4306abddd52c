#!/bin/bash
# Perf iteration visit: parity+timing check, work statistics, bench line, ncu launch list + full capture.
mkdir -p gpurun_out
python tools/gpu_check.py > gpurun_out/gpu_check.log 2>&1; tail -2 gpurun_out/gpu_check.log | cut -c1-300
python tools/gpu_stats.py 2>&1 | tail -2 | cut -c1-700
timeout 600 python bench.py --no-baselines 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_full_run.log 2>&1
ls gpurun_out
