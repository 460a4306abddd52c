"""configs[4] (4096-primitive synthetic tree @ 7680x4320 x 16 rays/pixel) on one GPU: a few frames, times printed (for ncu or A/B runs).
   python tools/gpu_cfg4_one.py [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csg_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sc = g.Scene.parse(g.Scene.generate_text(4096, seed=1234)); ctx = sc.upload(7680, 4320); ctx.set_supersampling(4)
cam, light = g.Camera(), g.Light()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = []
for k in range(n):
    flush.zero_(); torch.cuda.synchronize()
    ctx.enqueue(cam, light); ctx.sync()
    ms.append(ctx.last_frame_ms())
print("configs[4] ms", [round(m, 3) for m in ms], "median", round(float(np.median(ms[1:])), 4), ctx.info())
