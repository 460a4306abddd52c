"""configs[4] (4096-primitive synthetic tree, 7680x4320 x 16 rays/pixel) on one GPU with each pruning kernel; also at 4K x 1."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g
txt = g.Scene.generate_text(4096, 1234)
sc = g.Scene.parse(txt)
cam, light = g.Camera(), g.Light()
for (w, h, ss) in ((7680, 4320, 4), (3840, 2160, 1)):
    ctx = sc.upload(w, h)
    ctx.set_supersampling(ss)
    for mode, name in ((1, "flat"), (2, "walk")):
        ctx.set_pruning(mode)
        ctx.render(cam, light)
        ms = []
        for _ in range(5):
            ctx.enqueue(cam, light); ctx.sync(); ms.append(ctx.last_frame_ms())
        ctx.set_view_cache(True)
        fk = []
        for _ in range(4):
            ctx.enqueue(cam, light); ctx.sync(); fk.append(ctx.last_frame_ms())
        ctx.set_view_cache(False)
        print(w, h, ss, name, "frame %.3f ms, frame kernel only %.3f ms" % (float(np.median(ms)), float(np.median(fk[1:]))), ctx.prune_stats(), flush=True)
    ctx.close()
