#!/bin/bash
# round 2, visit C (N GPUs): shard tests incl. in-process multi-GPU, bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_shards.py -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -8 gpurun_out/c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c_bench_$N.json 2> gpurun_out/c_bench_$N.err; echo "bench rc=$?"
python - <<P
import json
for line in open('gpurun_out/c_bench_$N.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_slowest_rank_own_call'])
        print(d['timing'])
        print(json.dumps(d['configs'])); print(d['parity_n'])
P
tail -5 gpurun_out/c_bench_$N.err
