#!/bin/bash
# round 2, visit S: parity + bench + 8-way emulation of the current build (quick check after a kernel change)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
tail -3 gpurun_out/s_pytest.log
for K in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines --no-configs > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/s_bench.json'))
print('ms', round(d['ms_per_step'],4), 'static', round(d['static_view']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity_n']['mismatching_bytes'])
P
done
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "flat [18]"
