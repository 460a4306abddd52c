#!/bin/bash
# round 2, visit N: per-warp-tile sphere mask for the flat evaluation: parity (incl. the flat-evaluation scenes), bench per flat_leaves, emulation
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -4 gpurun_out/n_pytest.log
for FL in 0 15 12 8; do
  CSG_B200_FLAT_LEAVES=$FL timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines --no-configs > gpurun_out/n_bench_$FL.json 2> gpurun_out/n_bench_$FL.err
  python - <<P
import json
d=json.load(open('gpurun_out/n_bench_$FL.json'))
print('flat_leaves', $FL, 'ms', round(d['ms_per_step'],4), 'static', round(d['static_view']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity_n'])
P
done
for FL in 15; do
  echo "== emul flat_leaves $FL"
  CSG_B200_FLAT_LEAVES=$FL timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank
done
