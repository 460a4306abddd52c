#!/bin/bash
# round 2, visit D: flat evaluation — parity tests, then bench with several CSG_B200_FLAT_LEAVES
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
for FL in 0 16 24 48 0 24; do
  CSG_B200_FLAT_LEAVES=$FL timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines --no-configs > gpurun_out/d_bench_$FL.json 2> gpurun_out/d_bench_$FL.err
  python - <<P
import json
d=json.load(open('gpurun_out/d_bench_$FL.json'))
print('flat_leaves', $FL, 'ms', round(d['ms_per_step'],4), 'static', round(d['static_view']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity_n'])
P
done
