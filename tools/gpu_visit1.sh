#!/bin/bash
# Visit: parity of the new pruning kernel, then timings (flat vs walk), shard emulation.
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -x -q -m gpu -k "pruning or root_primitive or flat_and or view_cache or smoke or odd_sizes or batch" 2>&1 | tail -8 | tee gpurun_out/pytest_prune.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_flat.json | cut -c1-330
CSG_B200_PRUNE_WALK=1 timeout 300 python bench.py --no-baselines --steps 100 2>&1 | tail -1 | tee gpurun_out/bench_walk.json | cut -c1-330
timeout 400 python tools/gpu_shard_emul.py 40 2>&1 | tail -30
timeout 300 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py 2>&1 | tail -6 | tee gpurun_out/racecheck.log
