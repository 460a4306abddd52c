#!/bin/bash
# Instruction-cache view of several builds inside one box visit: A/B timing (tools/gpu_ab.sh) + ncu counters of the frame kernel
# (requests of the SMs' instruction caches, their hit rate, requests that reach the GPC-level cache).
#   tools/gpu_icache.sh rounds lib1.so lib2.so ...
R=$1; shift
bash tools/gpu_ab.sh $R "$@"
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__cache_requests_type_instruction_lookup_miss.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
for L in "$@"; do
  n=$(basename $L .so)
  CSG_B200_LIB=$PWD/$L timeout 300 ncu --metrics $M --clock-control none -k regex:csg_frame_kernel -s 6 -c 2 --csv --log-file gpurun_out/icache_$n.csv python tools/gpu_time_one.py 6 > /dev/null 2>&1
  python - "$n" <<'P'
import csv, sys
n = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/icache_{n}.csv")) if len(r) > 10]
h = rows[0]
out = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    if d.get("ID") == rows[-1][0]:
        out[d["Metric Name"]] = d["Metric Value"]
print(n, {k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("gcc__cache_requests_type_instruction", "gcc_instr"): v for k, v in out.items()})
P
done
