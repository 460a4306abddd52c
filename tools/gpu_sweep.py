"""Times each build/variants/libcsg_b200_*.so (and the default library) on the bench workloads.  GPU box only."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, os, json, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import csg_b200 as g
from oracle_py import scene_text, View, oblique_view
res = {}
for name, view, key in [("testCheese512", View(3840, 2160), "c512"), ("testCheese512", oblique_view(3840, 2160), "c512o"),
                        ("testCheese256", View(3840, 2160), "c256"), ("testWikipedia", View(1920, 1080), "wiki"),
                        ("testSphereCutByCubesAndCylinder", View(3840, 2160), "scut")]:
    sc = g.Scene.parse(scene_text(name), optimize=1)
    ctx = sc.upload(view.width, view.height)
    cam = g.Camera(pos=view.pos, pitch=view.pitch, yaw=view.yaw); light = g.Light()
    ms = []
    for i in range(25):
        ctx.enqueue(cam, light); ctx.sync()
        if i >= 5: ms.append(ctx.last_frame_ms())
    res[key] = round(float(np.median(ms)), 4)
    info = ctx.info()
    ctx.close()
res["ctas"] = info["ctas"]; res["smem"] = info["smem_bytes_per_cta"]; res["prune"] = info["prune"]
print(json.dumps(res))
'''
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["CSG_B200_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT], capture_output=True, text=True, env=env)
    print(os.path.basename(lib) if lib else "default", r.stdout.strip() or r.stderr[-300:], flush=True)
