"""Per-warp timeline of csg_frame_kernel (instrumented build libcsg_b200_probe.so): when warps start, pass the grid dependency,
take their first ticket and finish; the longest single warp tile.
   CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so python tools/gpu_frame_probe.py [shard_count]"""
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
g.lib.csg_debug_frame_probe.argtypes = [C.c_void_p, C.c_size_t]
for k in range(6):
    flush.zero_(); torch.cuda.synchronize()
    ctx.enqueue(cam, light); ctx.sync()
    buf = np.zeros((8192, 8), np.uint64)
    assert g.lib.csg_debug_frame_probe(buf.ctypes.data_as(C.c_void_p), buf.nbytes) == 0
ms = ctx.last_frame_ms()
b = buf[buf[:, 0] > 0].astype(np.int64)
t0 = b[:, 0].min()
us = lambda x: x / 1e3
print(f"shards {count}: frame {ms*1e3:.1f} us; warps {len(b)}")
print("warp start        : min %.1f mean %.1f max %.1f" % (us(b[:, 0].min() - t0), us(b[:, 0].mean() - t0), us(b[:, 0].max() - t0)))
print("phase 1 done      : mean %.1f max %.1f" % (us(b[:, 1].mean() - t0), us(b[:, 1].max() - t0)))
print("grid dependency ok: mean %.1f max %.1f" % (us(b[:, 2].mean() - t0), us(b[:, 2].max() - t0)))
w = b[b[:, 3] > 0]
print("first ticket      : mean %.1f max %.1f   (%d warps got tiles)" % (us(w[:, 3].mean() - t0), us(w[:, 3].max() - t0), len(w)))
print("warp end          : min %.1f mean %.1f max %.1f" % (us(b[:, 4].min() - t0), us(b[:, 4].mean() - t0), us(b[:, 4].max() - t0)))
print("tiles per warp    : mean %.1f max %d" % (w[:, 6].mean(), w[:, 6].max()))
print("longest tile      : mean %.1f us, max %.1f us (ticket %d of %d)" % (us(w[:, 5].mean()), us(w[:, 5].max()), w[np.argmax(w[:, 5]), 7], w[:, 6].sum()))
last = np.argsort(b[:, 4])[-5:]
for i in last:
    print("  late warp: end %.1f, tiles %d, its longest tile %.1f us (ticket %d), first ticket at %.1f" % (us(b[i, 4] - t0), b[i, 6], us(b[i, 5]), b[i, 7], us(b[i, 3] - t0)))
hist = np.histogram(us(w[:, 5]), bins=[0, 2, 5, 10, 15, 20, 30, 40, 60, 100])[0]
print("longest-tile histogram over warps (us bins 0,2,5,10,15,20,30,40,60,100):", hist.tolist())
