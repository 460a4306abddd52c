"""Phase timeline of csg_prune_flat_kernel (instrumented build: make -C cuda-csg-tree-raycasting_b200 libcsg_b200_probe.so).
   CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so python tools/gpu_prune_probe.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csg_b200 as g
import bench
txt, _ = bench.scene_bytes()
count = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sc = g.Scene.parse(txt); ctx = sc.upload_shard(bench.WIDTH, bench.HEIGHT, 0, 0, count)
fb = torch.empty(bench.WIDTH * bench.HEIGHT * 4, dtype=torch.uint8, device="cuda")
ctx.set_gather_target(fb.data_ptr())   # a plain local buffer: no start gate / join with the (absent) other shards
cam, light = g.Camera(), g.Light()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["start", "frustum", "leaf tests", "scan1+counts", "scan2", "emission", "refit", "op records", "desc+done atomic", "concat (last CTA)"]
acc = []
for k in range(8):
    flush.zero_(); torch.cuda.synchronize()
    ctx.enqueue(cam, light); ctx.sync()
    buf = np.zeros((4096, 16), np.uint64)
    g.lib.csg_debug_prune_probe.argtypes = [C.c_void_p, C.c_size_t]
    assert g.lib.csg_debug_prune_probe(buf.ctypes.data_as(C.c_void_p), buf.nbytes) == 0
    n = ctx.prune_stats()["traced_tiles"]
    t = buf[:n, :10].astype(np.int64)
    t0 = t[:, 0].min()
    if k >= 3:
        for i in range(1, 10):     # phases a tile skipped (empty tile, not the last CTA) take no time
            t[:, i] = np.where(t[:, i] == 0, t[:, i - 1], t[:, i])
        acc.append(t - t0)
a = np.mean(acc, axis=0)
print("tiles:", a.shape[0], " CTA start: min %.2f mean %.2f max %.2f us" % (a[:, 0].min() / 1e3, a[:, 0].mean() / 1e3, a[:, 0].max() / 1e3))
for i in range(1, 9):
    d = (a[:, i] - a[:, i - 1]) / 1e3
    print(f"{names[i]:22s} mean {d.mean():6.2f} us  max {d.max():6.2f}   (ends at mean {a[:, i].mean()/1e3:6.2f}, max {a[:, i].max()/1e3:6.2f})")
print("kernel (first CTA start to last probe): %.2f us" % (a.max() / 1e3))
