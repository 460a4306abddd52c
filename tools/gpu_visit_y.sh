#!/bin/bash
# last visit of round 2 (second try): tests, balance of the tile hand-out, ncu capture, bench, launch list — most important first
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 100 python tools/gpu_shard_balance.py 12 4,8 2>gpurun_out/balance.err | tee gpurun_out/shard_balance_new.jsonl
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/prof.ncu-rep
timeout 200 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_launches_run.log 2>&1
ls -la gpurun_out/launches.csv
