"""Instruction-cache footprint of the last csg_frame_kernel launch in an ncu capture: executed warp instructions per 128-byte line of
SASS (8 instructions), how many lines carry 90 / 99 / 99.9 % of the executed instructions, and the lines of the hot range that are
(nearly) never executed — cold code sitting between hot code.
   python tools/ncu_icache.py [gpurun_out/prof.ncu-rep]"""
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
lines = [sum(c[k:k + 8]) for k in range(0, len(c), 8)]
order = sorted(range(len(lines)), key=lambda k: -lines[k])
print(f"{len(c)} SASS instructions = {len(lines)} lines of 128 B = {len(lines) / 8:.1f} KB; {tot} warp instructions executed")
acc, marks = 0, [0.5, 0.9, 0.99, 0.999, 0.9999]
for n, k in enumerate(order):
    acc += lines[k]
    while marks and acc >= marks[0] * tot:
        print(f"  {100 * marks[0]:7.2f} % of the executed instructions lie in {n + 1:4d} lines = {(n + 1) / 8:5.1f} KB")
        marks.pop(0)
thr = tot / 96e6 * 20000      # a line executed fewer than ~20 k times per frame (of ~70 k tickets): cold
hot = [k for k in range(len(lines)) if lines[k] >= thr]
print(f"lines executed >= {thr:.0f} times: {len(hot)} = {len(hot) / 8:.1f} KB, spanning lines {hot[0]}..{hot[-1]} ({(hot[-1] - hot[0] + 1) / 8:.1f} KB)")
# cold runs inside the hot span
k = hot[0]
while k <= hot[-1]:
    if lines[k] < thr:
        j = k
        while j <= hot[-1] and lines[j] < thr:
            j += 1
        if j - k >= 2:
            print(f"  cold run: lines {k}..{j - 1} ({(j - k) * 128} B) at instr {8 * k}: {src[8 * k][:60]}")
        k = j
    else:
        k += 1
