#!/bin/bash
# round 2, visit E (8 GPUs): in-process multi-GPU shard tests, then bench under torchrun at 8, 4 (and 2) ranks
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_shards.py -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench_$N.json 2> gpurun_out/e_bench_$N.err; echo "bench $N rc=$?"
python - <<P
import json
for line in open('gpurun_out/e_bench_$N.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('N', $N, 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'own', round(d['e2e']['ms_per_step_slowest_rank_own_call'],4), 'wall', round(d['timing']['wall_ms_per_step_incl_flush_and_barriers'],3), 'maxranks', round(d['timing']['ms_per_step_max_over_ranks_own_spans'],4))
        print({k:(round(v['ms_per_step'],4), v.get('mismatching_bytes_vs_1gpu')) for k,v in d['configs'].items()}, d['parity_n'])
P
tail -2 gpurun_out/e_bench_$N.err | cut -c1-300
done
