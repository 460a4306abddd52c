#!/bin/bash
# multi-GPU visit (gpurun --gpus 8): multi-GPU parity tests, bench.py at 2/4/8 GPUs under torchrun, configs[4] at 1..8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests -x -q -m gpu -k "multi_gpu" 2>&1 | tail -2
N=$(nvidia-smi -L | wc -l)
for g in 2 4 8; do
  [ $g -gt $N ] && break
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500+g)) bench.py --gpus $g --steps 60 --warmup 5 2>&1 | tail -1 > gpurun_out/scale_$g.json
  python - <<PY
import json
d=json.load(open("gpurun_out/scale_$g.json"))
print($g, "gpus", round(d["ms_per_step"],4), "ms", round(d["value"]/1e9,2), "Grays/s e2e", round(d["e2e"]["ms_per_step"],3), "ms", d["clocks"])
PY
done
timeout 300 python tools/gpu_cfg5_full.py 2>&1 | cut -c1-160
