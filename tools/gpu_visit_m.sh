#!/bin/bash
# round 2, visit M: flat evaluation v3 (nearest Enter found during the scan, lazy far class, composite flats up to 30 spheres)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -4 gpurun_out/m_pytest.log
for FL in 0 15 30 24; do
  CSG_B200_FLAT_LEAVES=$FL timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines --no-configs > gpurun_out/m_bench_$FL.json 2> gpurun_out/m_bench_$FL.err
  python - <<P
import json
d=json.load(open('gpurun_out/m_bench_$FL.json'))
print('flat_leaves', $FL, 'ms', round(d['ms_per_step'],4), 'static', round(d['static_view']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity_n'])
P
done
for FL in 30 15; do
  echo "== emul flat_leaves $FL"
  CSG_B200_FLAT_LEAVES=$FL timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank
done
timeout 300 python tools/gpu_tile_iters.py
