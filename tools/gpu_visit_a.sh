#!/bin/bash
# round 2, visit A: GPU tests + 1-GPU bench (both arms)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json
tail -5 gpurun_out/a_bench.err
