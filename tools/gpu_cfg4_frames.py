"""configs[4] (4096-primitive synthetic tree @ 7680x4320 x 16 rays/pixel) on one GPU: a few frames, for ncu / timing."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sc = g.Scene.parse(g.Scene.generate_text(4096, 1234))
ctx = sc.upload(7680, 4320)
ctx.set_supersampling(4)
cam, light = g.Camera(), g.Light()
ms = []
for _ in range(n):
    ctx.enqueue(cam, light); ctx.sync(); ms.append(ctx.last_frame_ms())
print("configs[4] frame ms:", [round(m, 3) for m in ms], ctx.prune_stats(), ctx.info())
