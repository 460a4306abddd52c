#!/bin/bash
# round 2, visit G: full GPU tests, 1-GPU bench, split sweep on the shard emulation, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/g_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['static_view']['ms_per_step'], d['roofline']['frac'], d['roofline'].get('issue_frac'))
print({k:round(v['ms_per_step'],4) for k,v in d['configs'].items()}, d['parity_n'])
P
for SP in 0 1 2; do
  echo "== split $SP"
  CSG_B200_SPLIT=$SP timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank
  cp gpurun_out/shard_emul.json gpurun_out/g_shard_emul_split$SP.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/g_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/g_ncu_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel|csg_prune" -s 8 -c 4 -f -o gpurun_out/g_prof \
    python bench.py --steps 3 --warmup 3 --no-baselines --no-configs > gpurun_out/g_ncu_full_run.log 2>&1
ls -la gpurun_out | grep " g_"
