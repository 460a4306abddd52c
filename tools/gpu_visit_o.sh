#!/bin/bash
# round 2, visit O (2 GPUs): where a sharded frame's time goes between the GPUs (probe build, in-process context), and the same
# frame timed with the product library in-process and under torchrun
mkdir -p gpurun_out
CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so timeout 300 python tools/gpu_sync_probe.py 2 2>&1 | tail -8
timeout 300 python - <<'P'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import csg_b200 as g, bench
txt, _ = bench.scene_bytes()
for n in (1, 2):
    sc = g.Scene.parse(txt); ctx = sc.upload(bench.WIDTH, bench.HEIGHT, n_gpus=n)
    cam, light = g.Camera(), g.Light()
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
    ms = []
    for k in range(40):
        for d in range(n): flush[d].zero_()
        for d in range(n): torch.cuda.synchronize(d)
        ctx.enqueue(cam, light); ctx.sync()
        if k >= 8: ms.append(ctx.last_frame_ms())
    print(f"in-process, {n} GPU(s): frame {np.mean(ms)*1e3:.1f} us (min {np.min(ms)*1e3:.1f}, max {np.max(ms)*1e3:.1f})")
    ctx.close(); sc.close()
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 30 --warmup 5 --no-configs > gpurun_out/o_bench_2.json 2> gpurun_out/o_bench_2.err
python - <<'P'
import json
for line in open('gpurun_out/o_bench_2.json'):
    if line.startswith('{'):
        d=json.loads(line); t=d['timing']
        print('torchrun 2: ms', round(d['ms_per_step'],4), 'min', round(t['ms_per_step_min'],4), 'prequeued', t.get('ms_per_step_peers_prequeued'), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity_n'])
P
