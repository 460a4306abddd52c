"""Where the warp instructions of the last csg_frame_kernel launch in an ncu capture were executed: the kernel's own body (split at
the traversal loop) and each out-of-line device function behind it (cube_isect, shade_pixel, flat_eval, ... in link order).
   python tools/ncu_regions.py [gpurun_out/prof.ncu-rep]"""
import csv, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:csg_frame"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
launches, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        launches.append(cur)
    elif r and r[0] == "Address":
        hdr = r
    elif cur is not None and r and r[0].startswith("0x"):
        cur.append(r)
L = launches[-1]
iI = hdr.index("Instructions Executed")
c = [int(r[iI]) for r in L]
src = [r[1].strip() for r in L]
tot = sum(c)
exits = [k for k, s in enumerate(src) if s.split()[-1] == "EXIT" or s == "EXIT"]
main_end = exits[-1]
hot = [k for k in range(main_end) if c[k] > 2.2 * max(c[:main_end][600:1100] or [1])]          # executed far more often than once per warp and ticket: the traversal loop
def line(name, a, b):
    n = sum(c[a:b + 1])
    print(f"{name:34s} instr {a:5d}-{b:5d}  {n:11d} warp instructions  {100 * n / tot:5.1f} %")
print(f"total {tot} warp instructions, {len(L)} SASS instructions")
if hot:
    line("kernel body before the loop", 0, hot[0] - 1)
    line("traversal loop (operator machine)", hot[0], hot[-1])
    line("kernel body after the loop", hot[-1] + 1, main_end)
else:
    line("kernel body", 0, main_end)
prev = main_end + 1
for k, s in enumerate(src):
    if k > main_end and s.startswith("RET"):
        line(f"device function ({c[prev]} calls): {src[prev][:22]}", prev, k)
        prev = k + 1
