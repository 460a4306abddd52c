#!/bin/bash
# last visit of round 2: records of the final build (tests, bench, ncu launch list + full capture), then the balance of the tile
# hand-out before / after the traced rectangle is widened to a width coprime to the GPU count
# (ab/lib_r02i.so: the previous build; results: profiles/r02j_shard_balance.jsonl — this first try was 16 us slower per shard)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_launches_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_full_run.log 2>&1
: > gpurun_out/shard_balance.jsonl
CSG_B200_LIB=$PWD/ab/lib_r02i.so timeout 120 python tools/gpu_shard_balance.py 12 4,8 2>gpurun_out/balance.err | tee -a gpurun_out/shard_balance.jsonl
timeout 120 python tools/gpu_shard_balance.py 12 4,8 2>>gpurun_out/balance.err | tee -a gpurun_out/shard_balance.jsonl
ls -la gpurun_out/prof.ncu-rep gpurun_out/launches.csv
