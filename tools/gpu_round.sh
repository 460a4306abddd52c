#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both arms), ncu launch list + full capture of the frame kernel.
set -x
mkdir -p gpurun_out
nvidia-smi -L
nproc; grep -m1 "model name" /proc/cpuinfo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_full_run.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/memcheck.log 2>&1; tail -2 gpurun_out/memcheck.log
timeout 300 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/racecheck.log 2>&1; tail -2 gpurun_out/racecheck.log
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank0
ls -la gpurun_out
