#!/bin/bash
# One build, several settings of a tuning environment variable, timed round-robin inside one box visit:
#   tools/gpu_env_sweep.sh VAR rounds value1 value2 ...     (results: gpurun_out/sweep.jsonl)
VAR=$1; R=$2; shift 2
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for r in $(seq 1 $R); do
  for v in "$@"; do
    env $VAR=$v timeout 300 python tools/gpu_time_one.py 40 2>gpurun_out/sweep.err | sed "s#^{#{\"$VAR\": \"$v\", #" | tee -a gpurun_out/sweep.jsonl | cut -c1-300
  done
done
