#!/bin/bash
# round 2, visit B: all GPU tests (no -x), then configs[4] timing via bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-baselines > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/b_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['static_view']['ms_per_step'])
print(json.dumps(d['configs'])); print(d['parity_n'])
P
tail -3 gpurun_out/b_bench.err
