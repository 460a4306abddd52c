"""BASELINE.json configs[4] at full size: synthetic balanced tree of 4096 primitives, 7680x4320, 16 rays/pixel, on 1..N GPUs of one box
(in-process sharding, NVLink peer stores).  Prints ms/frame, primary rays/s and whether the N-GPU frame equals the 1-GPU frame."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g

W, H, SS = 7680, 4320, 4
txt = g.Scene.generate_text(4096, 1234)
sc = g.Scene.parse(txt)
cam, light = g.Camera(), g.Light()
base = None
ngpu = torch.cuda.device_count()
for n in [1, 2, 4, 8]:
    if n > ngpu:
        break
    ctx = sc.upload(W, H, n)
    ctx.set_supersampling(SS)
    img = ctx.render(cam, light).copy()
    ms = []
    for _ in range(5):
        ctx.enqueue(cam, light); ctx.sync(); ms.append(ctx.last_frame_ms())
    m = float(np.median(ms))
    if base is None:
        base = img
    print(json.dumps({"gpus": n, "ms_per_frame": m, "primary_rays_per_s": W * H * SS * SS / (m * 1e-3),
                      "identical_to_1gpu": bool(np.array_equal(img, base)), "info": ctx.info()}), flush=True)
    ctx.close()
