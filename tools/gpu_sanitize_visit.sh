#!/bin/bash
# compute-sanitizer memcheck + racecheck over tools/gpu_sanitize.py (every output mode, both pruning kernels, supersampling, batch, flat scenes)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py > gpurun_out/racecheck.log 2>&1; tail -3 gpurun_out/racecheck.log
