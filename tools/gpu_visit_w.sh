#!/bin/bash
# 8-GPU visit for the record: torchrun bench (every config, parity_n) and the in-process per-device timeline
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 --no-baselines 2>gpurun_out/bench_n.err | tail -1 > gpurun_out/bench_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${N}gpu.json"))
print("torchrun", $N, "gpus:", round(d["ms_per_step"],4), "ms; prequeued", d["timing"].get("ms_per_step_peers_prequeued"), "; e2e", d["e2e"]["ms_per_step"], d["e2e"].get("d2h_only_ms"), "; parity", d["parity_n"])
for k,v in d.get("configs",{}).items(): print(k, v["ms_per_step"], v.get("mismatching_bytes_vs_1gpu"))
PY
CSG_B200_LIB=$PWD/cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so timeout 100 python tools/gpu_sync_probe.py $N 2>>gpurun_out/bench_n.err | cut -c1-420 | tee gpurun_out/sync_probe_${N}.txt
