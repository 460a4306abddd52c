#!/bin/bash
# 8-GPU visit: where a sharded frame's time goes between the GPUs (in-process probe build), and the background fill dealt out over
# all shards against the root filling everything (CSG_B200_FILL_SHARED: a two-line knob in fill_params() that existed only for this
# visit — result in profiles/r02g_fill_experiment_8gpu.json, shared fill 27 % slower — and is not in the shipped library)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so timeout 300 python tools/gpu_sync_probe.py $N 2>&1 | tail -14 | tee gpurun_out/sync_probe_$N.txt | cut -c1-400
for F in 0 1 0 1; do
  CSG_B200_FILL_SHARED=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+F)) bench.py --gpus $N --steps 60 --warmup 5 --no-baselines --no-configs 2>&1 | tail -1 > gpurun_out/fill_$F.json
  python - <<PY
import json
d=json.load(open("gpurun_out/fill_$F.json"))
print("fill_shared=$F", $N, "gpus", round(d["ms_per_step"],4), "ms; idle", round(d["timing"]["ms_per_step_idle_start"],4), "parity", d["parity_n"]["mismatching_bytes"])
PY
done
