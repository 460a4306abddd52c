"""Frame time (root's events) of an N-GPU in-process context for ONE build of the library (CSG_B200_LIB) — Cheese512 @ 4K, L2 of every
device flushed in front of every frame — and whether the gathered frame equals the one-GPU frame.
   CSG_B200_LIB=... python tools/gpu_peer_time.py [n_gpus] [frames]"""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csg_b200 as g
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt, _ = bench.scene_bytes()
cam, light = g.Camera(), g.Light()
sc = g.Scene.parse(txt)
one = sc.upload(bench.WIDTH, bench.HEIGHT)
one.enqueue(cam, light); one.sync()
ref = one.read_framebuffer(np.empty((bench.HEIGHT, bench.WIDTH, 4), np.uint8)).copy()
one.close()
ctx = sc.upload(bench.WIDTH, bench.HEIGHT, n_gpus=n)
flush = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
ms = []
for k in range(frames + 8):
    for d in range(n):
        flush[d].zero_()
    for d in range(n):
        torch.cuda.synchronize(d)
    ctx.enqueue(cam, light); ctx.sync()
    if k >= 8:
        ms.append(ctx.last_frame_ms())
img = ctx.read_framebuffer(np.empty((bench.HEIGHT, bench.WIDTH, 4), np.uint8))
print(json.dumps({"lib": os.path.basename(os.environ.get("CSG_B200_LIB", "shipped")), "peer_local": os.environ.get("CSG_EXP_PEER_LOCAL", "0"), "n_gpus": n,
                  "frame_us": round(float(np.median(ms)) * 1e3, 2), "min_us": round(float(np.min(ms)) * 1e3, 2),
                  "mismatching_bytes": int(np.count_nonzero(img != ref))}))
ctx.close(); sc.close()
