#!/bin/bash
# One GPU-box visit for the record: tests, smoke, bench (both arms), ncu launch list + full capture of the kernels, sanitizers,
# per-GPU share of an N-way sharded frame, probe timelines.  Outputs under gpurun_out/ (tools/ncu_summary.py turns them into profiles/).
set -x
mkdir -p gpurun_out
nvidia-smi -L
nproc; grep -m1 "model name" /proc/cpuinfo
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench.json; cut -c1-600 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/prof \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/ncu_full_run.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/memcheck.log 2>&1; tail -2 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/racecheck.log 2>&1; tail -2 gpurun_out/racecheck.log
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank
export CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so
timeout 300 python tools/gpu_frame_probe.py 8 2>&1 | tail -20 | tee gpurun_out/frame_probe_8.txt
timeout 300 python tools/gpu_frame_probe.py 1 2>&1 | tail -20 | tee gpurun_out/frame_probe_1.txt
timeout 300 python tools/gpu_prune_probe.py 8 2>&1 | tail -12 | tee gpurun_out/prune_probe_8.txt
timeout 300 python tools/gpu_prune_probe.py 1 2>&1 | tail -12 | tee gpurun_out/prune_probe_1.txt
unset CSG_B200_LIB
ls -la gpurun_out | tail -30
