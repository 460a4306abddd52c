"""Config-5-like check on a GPU box: synthetic 4096-primitive tree; pruning statistics, pruned vs unpruned frame equality, timing."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g

def run(nprims, w, h, ss, frames=5):
    txt = g.Scene.generate_text(nprims, 1234)
    sc = g.Scene.parse(txt)
    ctx = sc.upload(w, h)
    ctx.set_supersampling(ss)
    cam, light = g.Camera(), g.Light()
    out = {}
    for prune in (1, 0):
        ctx.set_pruning(prune)
        img = ctx.render(cam, light).copy()
        ms = []
        for _ in range(frames):
            ctx.enqueue(cam, light); ctx.sync(); ms.append(ctx.last_frame_ms())
        out[prune] = (img, float(np.median(ms)), ctx.prune_stats() if prune else None)
    same = bool((out[1][0] == out[0][0]).all())
    print(json.dumps({"prims": nprims, "w": w, "h": h, "ss": ss, "ms_pruned": out[1][1], "ms_unpruned": out[0][1], "identical": same,
                      "stats": out[1][2], "info": ctx.info()}), flush=True)
    ctx.close(); sc.close()

run(512, 1920, 1080, 1)
run(4096, 1920, 1080, 1)
run(4096, 3840, 2160, 1)
run(4096, 1920, 1080, 4, frames=3)
