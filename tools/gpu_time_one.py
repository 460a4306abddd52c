"""Timing of ONE build of the library (CSG_B200_LIB) for A/B comparisons inside one box visit (tools/gpu_ab.sh):
Cheese512 @ 4K on one GPU — frame, frame with the view cache (frame kernel alone), shard 0 of 8 — medians over `frames` frames with
the L2 flushed in between, and a CRC of the frame (the builds under comparison must agree on it).
   CSG_B200_LIB=... python tools/gpu_time_one.py [frames]"""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csg_b200 as g
import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
txt, _ = bench.scene_bytes()
cam, light = g.Camera(), g.Light()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(ctx, n):
    ms = []
    for k in range(n + 5):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.enqueue(cam, light)
        ctx.sync()
        if k >= 5:
            ms.append(ctx.last_frame_ms())
    return float(np.median(ms)), float(np.min(ms))


out = {"lib": os.path.basename(g.binding.LIB_PATH) if hasattr(g, "binding") else os.environ.get("CSG_B200_LIB", "default")}
sc = g.Scene.parse(txt)
ctx = sc.upload(bench.WIDTH, bench.HEIGHT)
out["frame"], out["frame_min"] = timed(ctx, frames)
ctx.set_view_cache(True)
out["static"], _ = timed(ctx, frames // 2)
ctx.set_view_cache(False)
img = ctx.read_framebuffer(np.empty((bench.HEIGHT, bench.WIDTH, 4), np.uint8))
out["crc"] = zlib.crc32(img.tobytes())
ctx.close()
fb = torch.empty(bench.WIDTH * bench.HEIGHT * 4, dtype=torch.uint8, device="cuda")
c8 = sc.upload_shard(bench.WIDTH, bench.HEIGHT, 0, 0, 8)
c8.set_gather_target(fb.data_ptr())
out["shard0of8"], out["shard0of8_min"] = timed(c8, frames)
c8.set_view_cache(True)
out["shard0of8_static"], _ = timed(c8, frames // 2)
c8.close(); sc.close()
print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in out.items()}))
