import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import csg_b200 as g
from oracle_py import scene_text
sc = g.Scene.parse(scene_text("testCheese512")); ctx = sc.upload(3840, 2160); cam, light = g.Camera(), g.Light()
host = torch.empty(3840*2160*4, dtype=torch.uint8).pin_memory()
for _ in range(5): ctx.render(cam, light, host.data_ptr())
ts = []
for _ in range(20):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.render(cam, light, host.data_ptr()); ts.append(time.perf_counter() - t0)
print("bands", os.environ.get("CSG_B200_BANDS"), "e2e ms", round(1e3*float(np.median(ts)), 4), "kernel span ms", round(ctx.last_frame_ms(), 4))
