#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/frame', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'static', d['static_view']['ms_per_step'])"
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep "rank0"
