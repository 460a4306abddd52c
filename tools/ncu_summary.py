"""Summarises gpurun_out/prof.ncu-rep (ncu --set full capture of csg_frame_kernel) and gpurun_out/launches.csv into
profiles/<tag>_*.  Run here (no GPU needed): python tools/ncu_summary.py r01a"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
rep = os.path.join(src, "prof.ncu-rep")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__cycles_elapsed.max", "lts__t_sectors_op_write.sum",
        "lts__t_bytes.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        # instruction fetch: the SMs' instruction caches and what reaches the GPC-level cache (DESIGN 4.2 "Instruction cache")
        "sm__icc_requests.sum", "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
summ = []
for r in data:
    d = {}
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": r[i], "unit": units[i]}
    summ.append(d)


def num(d, k):
    try:
        v = float(d[k]["value"].replace(",", ""))
    except Exception:
        return None
    u = d[k]["unit"]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
    return v * mult


out = {"tag": tag, "command": "ncu --set full --clock-control none --import-source on -k regex:'csg_frame_kernel|csg_prune' -s 8 -c 4 python bench.py --steps 3 --warmup 3 --no-baselines",
       "launches": summ}
frame = [d for d in summ if "frame" in d["Kernel Name"]["value"]]
prune = [d for d in summ if "prune" in d["Kernel Name"]["value"]]
if frame:
    d = frame[-1]
    cyc = num(d, "sm__cycles_elapsed.avg") or 0
    fadd, fmul, ffma = ((num(d, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") or 0) * cyc for k in ("fadd", "fmul", "ffma"))
    out["frame_kernel"] = {"us": num(d, "gpu__time_duration.sum"), "warp_instructions": num(d, "smsp__inst_executed.sum"),
                           "executed_fp32_flop": fadd + fmul + 2 * ffma,
                           "fp32_thread_instructions": {"fadd": fadd, "fmul": fmul, "ffma": ffma},
                           "fma_pipe_pct": num(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                           "alu_pipe_pct": num(d, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                           "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                           "branch_uniform_pct": num(d, "smsp__sass_average_branch_targets_threads_uniform.pct"),
                           "threads_per_instruction": num(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                           "fp32_flop_per_cycle": (fadd + fmul + 2 * ffma) / cyc if cyc else None, "fp32_flop_per_cycle_peak": 148 * 128 * 2,
                           "local_load_instructions": num(d, "sass__inst_executed_local_loads"),
                           "local_store_instructions": num(d, "sass__inst_executed_local_stores"),
                           "registers": num(d, "launch__registers_per_thread"),
                           "sm_instruction_cache_hit_pct": num(d, "sm__icc_request_hit_rate.pct"),
                           "gpc_cache_instruction_requests_pct_of_peak": num(d, "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed"),
                           "no_instruction_stall_per_issue": num(d, "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")}
if prune:
    d = prune[-1]
    out["prune_kernel"] = {"us": num(d, "gpu__time_duration.sum"), "warp_instructions": num(d, "smsp__inst_executed.sum"),
                           "registers": num(d, "launch__registers_per_thread")}
if frame:
    d = frame[-1]
    rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
    out["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
    out["note"] = ("dram bytes are per launch of csg_frame_kernel; the 33 MB RGBA8 framebuffer is written into the 126 MB L2 and is "
                   "not evicted to HBM within the launch, so DRAM traffic is far below the 33 MB of algorithmic output bytes")
# the kernel sources this capture was taken from (bench.py compares it with the sources it runs: roofline.executed.profile_matches_source)
sys.path.insert(0, ROOT)
import bench  # noqa: E402
out["source_hash"] = bench.source_hash()
with open(os.path.join(dst, f"{tag}_ncu_summary.json"), "w") as f:
    json.dump(out, f, indent=1)
shutil.copy(os.path.join(dst, f"{tag}_ncu_summary.json"), os.path.join(dst, "ncu_summary.json"))
if os.path.exists(os.path.join(src, "launches.csv")):
    shutil.copy(os.path.join(src, "launches.csv"), os.path.join(dst, f"{tag}_launches.csv"))

# instruction mix + hottest SASS from the source page
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
with open(os.path.join(dst, f"{tag}_instruction_mix.txt"), "w") as f:
    if secs:
        s = secs[-1]
        ix = {n: i for i, n in enumerate(s["hdr"])}
        tot, per_op, stall = 0, collections.Counter(), collections.Counter()
        stall_cols = [c for c in s["hdr"] if c.startswith("stall_")]
        for r in s["rows"]:
            try:
                n = int(r[ix["Instructions Executed"]])
            except Exception:
                continue
            tot += n
            t = r[ix["Source"]].strip().split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            per_op[op] += n
            for c in stall_cols:
                try:
                    stall[c] += int(r[ix[c]])
                except Exception:
                    pass
        f.write(f"kernel: {s['name']}\nwarp instructions executed: {tot}\n\nopcode mix (warp instructions):\n")
        for op, n in per_op.most_common(40):
            f.write(f"  {op:12s} {n:12d} {100 * n / tot:5.1f}%\n")
        ts = sum(stall.values()) or 1
        f.write("\nwarp stall samples by reason:\n")
        for c, n in stall.most_common(12):
            f.write(f"  {c:28s} {n:9d} {100 * n / ts:5.1f}%\n")
# hottest source lines of both kernels (instructions executed, stall samples)
for kern, name in (("csg_frame_kernel", "frame"), ("csg_prune", "prune")):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kern, "45"], capture_output=True, text=True).stdout
    with open(os.path.join(dst, f"{tag}_{name}_kernel_lines.txt"), "w") as f:
        f.write(txt)
print(json.dumps({k: v["value"] for k, v in (summ[-1] if summ else {}).items()}, indent=1))
print(open(os.path.join(dst, f"{tag}_instruction_mix.txt")).read()[-700:])
