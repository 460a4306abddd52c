#!/bin/bash
# round 2, visit P: bench with the d2h-only diagnostic; ncu capture of the supersampling frame kernel on configs[4]
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-baselines --no-configs > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/p_bench.json'))
print('ms', round(d['ms_per_step'],4), 'static', round(d['static_view']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'd2h only', d['e2e'].get('d2h_only_ms'))
P
tail -2 gpurun_out/p_bench.err
timeout 300 python tools/gpu_cfg4_frames.py 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csg_frame_kernel" -s 2 -c 1 -f -o gpurun_out/p_cfg4_prof python tools/gpu_cfg4_frames.py 3 > gpurun_out/p_ncu_cfg4.log 2>&1
tail -3 gpurun_out/p_ncu_cfg4.log
