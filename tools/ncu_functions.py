"""Executed warp instructions of the last csg_frame_kernel launch of an ncu capture, by SOURCE function (inlined code included):
nvdisasm's line table of the shipped library gives every SASS instruction its source line, the capture its execution count.
   python tools/ncu_functions.py [gpurun_out/prof.ncu-rep] [kernel-name substring, default ILi0ELi768ELb0ELb0ELb1]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
kname = sys.argv[2] if len(sys.argv) > 2 else "csg_frame_kernelILi0ELi768ELb0ELb0ELb1"
lib = os.path.join(ROOT, "cuda-csg-tree-raycasting_b200", "libcsg_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "scene" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l][0]
end = [i for i, l in enumerate(dis) if i > start and l.startswith("//--------------------- .text")][0]
cur, srcmap = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", l):
        srcmap.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:csg_frame"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
launches, cu, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cu = []
        launches.append(cu)
    elif r and r[0] == "Address":
        hdr = r
    elif cu is not None and r and r[0].startswith("0x"):
        cu.append(r)
L = launches[-1]
c = [int(r[hdr.index("Instructions Executed")]) for r in L]
assert len(c) == len(srcmap), (len(c), len(srcmap), "the capture is not of this build")
src = {f: open(os.path.join(ROOT, "cuda-csg-tree-raycasting_b200", "csrc", f)).read().split("\n") for f in ("csg_frame.cuh", "csg_kernel.cuh")}


def fn_of(loc):
    if loc is None:
        return "(no line)"
    f, n = loc
    if f not in src:
        return f
    for i in range(min(n, len(src[f])) - 1, -1, -1):
        l = src[f][i]
        m = re.match(r"(?:template.*>\s*)?__(?:device|global)__.*?(\w+)\(", l)
        if m and not l.startswith(" "):
            return m.group(1)
        m = re.match(r"\s+auto (\w+) = \[", l)          # lambdas of the kernel body (make_ray, ordered_tile)
        if m and i < n - 1 and n - i < 25:
            return m.group(1)
    return f


tot = sum(c)
agg = collections.Counter()
for k, loc in enumerate(srcmap):
    agg[fn_of(loc)] += c[k]
print(f"{kname}: {tot} warp instructions executed, {len(c)} SASS instructions")
for fn, n in agg.most_common(25):
    print(f"  {fn:32s} {n:11d}  {100 * n / tot:5.1f} %")
