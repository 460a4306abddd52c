#!/bin/bash
# round 2, visit Q: height-bounded re-balancing: parity, configs[4] with optimize 1 / 0, bench with configs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
timeout 300 python tools/gpu_cfg4_opt.py 2>&1 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print('ms', round(d['ms_per_step'],4), 'static', round(d['static_view']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'd2h only', d['e2e'].get('d2h_only_ms'))
print({k:round(v['ms_per_step'],4) for k,v in d['configs'].items()}, d['parity_n'])
P
