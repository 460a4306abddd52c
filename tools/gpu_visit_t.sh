#!/bin/bash
# A/B of the pruning kernel's latency chain (CTA-wide refit, early topology loads, L2 prefetch of the scene), one box visit
# (ab/lib_*.so: builds of csrc/ with -DCSG_EXP_REFIT / -DCSG_EXP_TOPO / -DCSG_EXP_PREFETCH, macros that existed only for this visit;
# what won — refit + prefetch — is unconditional now; results: profiles/r02h_ab_prune_variants.jsonl)
mkdir -p gpurun_out
bash tools/gpu_ab.sh 3 ab/lib_base.so ab/lib_r.so ab/lib_rt.so ab/lib_rtp.so
cp gpurun_out/ab.jsonl gpurun_out/ab_t.jsonl
CSG_B200_LIB=$PWD/ab/lib_rtp.so timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_rtp.log
