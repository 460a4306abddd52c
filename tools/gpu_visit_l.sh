#!/bin/bash
# round 2, visit L (8 GPUs): topology, in-process multi-GPU shard tests, bench under torchrun at 8 (full line), 8 without flat evaluation, 4, 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/l_topo.txt 2>&1; head -12 gpurun_out/l_topo.txt | cut -c1-150
timeout 600 python -m pytest tests/test_gpu_shards.py -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -3 gpurun_out/l_pytest.log
run() {  # N tag extra-args env
  N=$1; TAG=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --warmup 5 "$@" > gpurun_out/l_bench_$TAG.json 2> gpurun_out/l_bench_$TAG.err; echo "bench $TAG rc=$?"
  python - <<P
import json
for line in open('gpurun_out/l_bench_$TAG.json'):
    if line.startswith('{'):
        d=json.loads(line)
        t=d['timing']
        print('$TAG', 'ms', round(d['ms_per_step'],4), 'min', round(t['ms_per_step_min'],4), 'max', round(t['ms_per_step_max'],4), 'prequeued', t.get('ms_per_step_peers_prequeued'), 'e2e', round(d['e2e']['ms_per_step'],4), 'own', round(d['e2e']['ms_per_step_slowest_rank_own_call'],4), 'wall', round(t['wall_ms_per_step_incl_flush_and_barriers'],3))
        if d.get('configs'): print({k:(round(v['ms_per_step'],4), v.get('mismatching_bytes_vs_1gpu')) for k,v in d['configs'].items()})
        print(d['parity_n'])
P
  tail -2 gpurun_out/l_bench_$TAG.err | cut -c1-300
}
run 8 8 --steps 20
run 4 4 --steps 20
run 2 2 --steps 20

