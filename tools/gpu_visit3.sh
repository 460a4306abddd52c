#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "pruning or root_primitive or flat_and or view_cache or odd_sizes or batch" 2>&1 | tail -2
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-230
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/ncu_launches_run.log 2>&1
grep "prune" gpurun_out/launches.csv | head -3 | cut -d, -f5,9,12-
for cfg in "768 1" "384 2" "384 1" "256 3" "256 2" "256 1"; do
  set -- $cfg
  echo "shape $1 ctas/sm $2"
  CSG_B200_SHAPE=$1 CSG_B200_CTAS_PER_SM=$2 timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "rank0"
done
