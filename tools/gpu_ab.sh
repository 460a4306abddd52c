#!/bin/bash
# A/B timing of several builds of the library inside ONE box visit: tools/gpu_ab.sh rounds lib1.so lib2.so ...
# (builds alternate round by round, so that drift of the box hits them alike); results: gpurun_out/ab.jsonl
R=$1; shift
mkdir -p gpurun_out; : > gpurun_out/ab.jsonl
for r in $(seq 1 $R); do
  for L in "$@"; do
    CSG_B200_LIB=$PWD/$L timeout 300 python tools/gpu_time_one.py 40 2>gpurun_out/ab.err | sed "s#^{#{\"build\": \"$L\", #" | tee -a gpurun_out/ab.jsonl | cut -c1-400
  done
done
