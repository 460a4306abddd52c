#!/bin/bash
# round 2, visit R: ncu launch list + full capture (source counters) of the current kernels on the bench frame
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log
ls -la gpurun_out/prof.ncu-rep gpurun_out/launches.csv
