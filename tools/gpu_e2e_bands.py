"""End-to-end csg_render() into pinned host memory for different band counts (CSG_B200_BANDS)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import csg_b200 as g
import bench
txt, _ = bench.scene_bytes()
sc = g.Scene.parse(txt); ctx = sc.upload(bench.WIDTH, bench.HEIGHT)
cam, light = g.Camera(), g.Light()
host = torch.empty(bench.WIDTH * bench.HEIGHT * 4, dtype=torch.uint8).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for bands in (1, 2, 3, 4, 5, 6, 8):
    os.environ["CSG_B200_BANDS"] = str(bands)
    ts = []
    for k in range(25):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); ctx.render(cam, light, host.data_ptr()); torch.cuda.synchronize(); t1 = time.perf_counter()
        if k >= 5: ts.append(t1 - t0)
    print(bands, "bands: e2e %.1f us (min %.1f)" % (np.mean(ts) * 1e6, np.min(ts) * 1e6), flush=True)
