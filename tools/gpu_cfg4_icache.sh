#!/bin/bash
# configs[4] on one GPU: time, then ncu instruction-cache counters of its frame kernel
mkdir -p gpurun_out
python tools/gpu_cfg4_one.py 6 2>&1 | tail -1 | cut -c1-400
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:"csg_frame_kernel|csg_prune" -s 2 -c 2 --csv --log-file gpurun_out/icache_cfg4.csv python tools/gpu_cfg4_one.py 3 > /dev/null 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/icache_cfg4.csv")) if len(r) > 10]
h = rows[0]
out = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    out.setdefault(d["Kernel Name"][:40], {})[d["Metric Name"].replace("smsp__average_warps_issue_stalled_", "stall_").replace("gcc__cache_requests_type_instruction", "gcc_instr")] = d["Metric Value"]
for k, v in out.items(): print(k, v)
P
