"""Small renders of every code path for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import csg_b200 as g
import scenes
from oracle_py import View, orbit_view

def go(txt, w, h, ss=1, marks=False, opt=1, walk=False):
    os.environ["CSG_B200_MARKS_FIRST"] = "1" if marks else "0"
    sc = g.Scene.parse(txt, optimize=opt); ctx = sc.upload(w, h); ctx.set_supersampling(ss)
    ctx.set_pruning(2 if (walk or marks) else 1)   # 1: csg_prune_flat_kernel (default), 2: csg_prune_kernel (walk / leaf marks)
    v = orbit_view(w, h, 7, radius=6.0)
    cam, light = g.Camera(pos=v.pos, pitch=v.pitch, yaw=v.yaw), g.Light()
    a = ctx.render(cam, light).copy()
    if ss == 1:
        ctx.render_aov(cam)
    ctx.render_f32(cam, light)
    ctx.render_batch([cam, g.Camera()], light)
    ctx.close(); sc.close()
    return int(a.sum())

print(go(scenes.INLINE["nested"], 200, 120))
print(go(scenes.INLINE["nested"], 200, 120, ss=4))
print(go(scenes.INLINE["deep_left_chain"], 130, 70, ss=2, marks=True))
print(go(g.Scene.generate_text(300, 5), 256, 144))
print(go(g.Scene.generate_text(300, 5), 256, 144, walk=True))
print(go(scenes.INLINE["nested"], 200, 120, walk=True))
print(go(g.Scene.generate_text(300, 5), 256, 144, marks=True, opt=0))
for name in ("sphere_chain_12", "sphere_chain_40", "sphere_union_root", "spheres_minus_sphere"):   # flat evaluation of sphere unions (flat_eval)
    print(go(scenes.INLINE[name], 256, 144))
print(go(scenes.INLINE["sphere_chain_26"], 128, 72, ss=2))
print(go(scenes.DUP_CHAIN, 160, 90, opt=0))
if "testCheese256" in scenes.corpus_names():
    print(go(scenes.text_of("corpus:testCheese256"), 320, 180))
print("done")
