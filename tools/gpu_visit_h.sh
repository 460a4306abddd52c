#!/bin/bash
# round 2, visit H: split sweep on the shard emulation (after the n_tiles fix), parity tests with split forced
mkdir -p gpurun_out
for SP in 0 1 2; do
  echo "== split $SP"
  CSG_B200_SPLIT=$SP timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank
  cp gpurun_out/shard_emul.json gpurun_out/h_shard_emul_split$SP.json
done
echo "== default split"
timeout 300 python tools/gpu_shard_emul.py 30 flat 2>&1 | grep rank
CSG_B200_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/h_pytest_split1.log 2>&1; tail -3 gpurun_out/h_pytest_split1.log
CSG_B200_SPLIT=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/h_pytest_split2.log 2>&1; tail -3 gpurun_out/h_pytest_split2.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
