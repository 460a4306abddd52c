#!/bin/bash
# A/B of the pruning kernel's latency chain (CTA-wide refit, early topology loads, L2 prefetch of the scene), one box visit
mkdir -p gpurun_out
bash tools/gpu_ab.sh 3 ab/lib_base.so ab/lib_r.so ab/lib_rt.so ab/lib_rtp.so
cp gpurun_out/ab.jsonl gpurun_out/ab_t.jsonl
CSG_B200_LIB=$PWD/ab/lib_rtp.so timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_rtp.log
