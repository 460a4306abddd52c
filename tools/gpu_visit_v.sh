#!/bin/bash
# 2-GPU visit: gate / join changes (first kernel only enters the gate, one poller per peer, relaxed start word, one fence per CTA)
# against the r02h build; in-process multi-GPU test; torchrun bench with parity_n; per-device timeline
# (ab/lib_r02h.so: the previous build of the library, kept beside the new one for the A/B; results: profiles/r02i_gate_join_ab_2gpu.jsonl)
mkdir -p gpurun_out; : > gpurun_out/peer_ab.jsonl
N=$(nvidia-smi -L | wc -l)
P=$PWD/cuda-csg-tree-raycasting_b200
for r in 1 2; do
  for L in "$PWD/ab/lib_r02h.so" "$P/libcsg_b200.so"; do
    CSG_B200_LIB=$L timeout 200 python tools/gpu_peer_time.py $N 80 2>gpurun_out/peer_ab.err | tee -a gpurun_out/peer_ab.jsonl
  done
done
timeout 300 python -m pytest tests/test_gpu_shards.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/pytest_shards_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 100 --warmup 5 --no-baselines --no-configs 2>gpurun_out/bench_n.err | tail -1 > gpurun_out/bench_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${N}gpu.json"))
print("torchrun", $N, "gpus:", round(d["ms_per_step"],4), "ms; prequeued", d["timing"].get("ms_per_step_peers_prequeued"), "; e2e", d["e2e"]["ms_per_step"], "; parity", d["parity_n"])
PY
CSG_B200_LIB=$P/libcsg_b200_probe.so timeout 200 python tools/gpu_sync_probe.py $N 2>>gpurun_out/peer_ab.err | cut -c1-420 | tee gpurun_out/sync_probe_${N}.txt
