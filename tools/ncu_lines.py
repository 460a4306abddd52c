"""Per-source-line instruction / sample shares of the last csg_frame_kernel launch in gpurun_out/prof.ncu-rep."""
import csv, subprocess, sys, collections
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof.ncu-rep"
kern = sys.argv[2] if len(sys.argv) > 2 else "csg_frame_kernel"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# sections: (File Path, Function Name) repeated per launch; keep the LAST launch of the kernel
launches = []; cur = None; fname = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fname = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ixI = hdr.index("Instructions Executed"); ixS = hdr.index("# Samples")
        cur = {"file": fpath, "fn": fname, "rows": []}; launches.append(cur); continue
    if cur is not None and r[0].isdigit(): cur["rows"].append(r)
sel = [l for l in launches if kern in l["fn"]]
# group launches: a new launch starts when a file repeats
groups = []; seen = set(); g = []
for l in sel:
    if l["file"] in seen: groups.append(g); g = []; seen = set()
    seen.add(l["file"]); g.append(l)
if g: groups.append(g)
g = groups[-1]
agg = {}
for l in g:
    for r in l["rows"]:
        try: agg[(l["file"], int(r[0]), r[1].strip()[:105])] = (int(r[ixI]), int(r[ixS]))
        except Exception: pass
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("launches of", kern, ":", len(groups), " total warp instr:", tot, " samples:", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.2f}% inst {100*v[1]/max(ts,1):5.2f}% smp  {k[0]}:{k[1]}  {k[2]}")
