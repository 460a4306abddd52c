#!/bin/bash
# bench.py (headline + every BASELINE config, no baselines) for several builds of the library inside one box visit:
#   tools/gpu_bench_variants.sh lib1.so lib2.so ...      -> gpurun_out/bv_<name>.json
mkdir -p gpurun_out
for L in "$@"; do
  n=$(basename $L .so)
  CSG_B200_LIB=$PWD/$L timeout 600 python bench.py --steps 40 --warmup 5 --no-baselines > gpurun_out/bv_$n.json 2> gpurun_out/bv_$n.err
  python - "$n" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bv_{n}.json").read().strip().splitlines()[-1])
    c = {k: round(v.get("ms_per_step", -1), 4) for k, v in d.get("configs", {}).items() if isinstance(v, dict)}
    print(n, "ms", round(d["ms_per_step"], 4), "idle", round(d["timing"]["ms_per_step_idle_start"], 4), "static", round(d["static_view"]["ms_per_step"], 4),
          "e2e", round(d["e2e"]["ms_per_step"], 4), "parity", d["parity_n"]["mismatching_bytes"], c)
except Exception as e:
    print(n, "failed:", e, open(f"gpurun_out/bv_{n}.err").read()[-600:])
P
done
