#!/bin/bash
# 1/2/4/8-GPU scaling of the bench workload on one box (run with gpurun --gpus 8).
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for g in 1 2 4 8; do
  [ $g -gt $N ] && break
  if [ $g -eq 1 ]; then
    python bench.py --gpus 1 --steps 100 --warmup 5 --no-baselines 2>&1 | tail -1 > gpurun_out/scale_$g.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500+g)) bench.py --gpus $g --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_$g.json
  fi
  python - <<PY
import json
d=json.load(open("gpurun_out/scale_$g.json"))
print($g, "gpus", round(d["ms_per_step"],4), "ms", round(d["value"]/1e9,2), "Grays/s e2e", round(d["e2e"]["ms_per_step"],3), "ms")
PY
done
