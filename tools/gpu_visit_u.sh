#!/bin/bash
# 2-GPU visit: why a peer needs longer than the root for the same share (NVLink pixel stores? the per-thread system fence?)
# (ab/lib_exp*.so: builds of a scratch copy of csrc/ with an environment switch CSG_EXP_PEER_LOCAL — peers store into local memory, wrong
# frame, timing only — and -DCSG_EXP_JOIN1 = one system fence per CTA; results: profiles/r02i_peer_experiment*_2gpu.*)
mkdir -p gpurun_out; : > gpurun_out/peer_exp.jsonl
N=$(nvidia-smi -L | wc -l)
P=$PWD/cuda-csg-tree-raycasting_b200
for r in 1; do
  for V in "$P/libcsg_b200.so:0" "$PWD/ab/lib_expj.so:0" "$PWD/ab/lib_exp.so:1" "$PWD/ab/lib_expj.so:1"; do
    CSG_B200_LIB=${V%%:*} CSG_EXP_PEER_LOCAL=${V##*:} timeout 200 python tools/gpu_peer_time.py $N 60 2>gpurun_out/peer_exp.err | tee -a gpurun_out/peer_exp.jsonl
  done
done
for V in "$PWD/ab/lib_exp_probe.so:0" "$PWD/ab/lib_expj_probe.so:0" "$PWD/ab/lib_exp_probe.so:1"; do
  echo "== $(basename ${V%%:*}) peer_local=${V##*:}" | tee -a gpurun_out/peer_exp_probe.txt
  CSG_B200_LIB=${V%%:*} CSG_EXP_PEER_LOCAL=${V##*:} timeout 200 python tools/gpu_sync_probe.py $N 2>>gpurun_out/peer_exp.err | cut -c1-420 | tee -a gpurun_out/peer_exp_probe.txt
done
