"""Renders a few frames of one scene (for ncu): python tools/gpu_prof_scene.py <scene> <w> <h> [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g
from oracle_py import scene_text, orbit_view
name, w, h = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 6
sc = g.Scene.parse(scene_text(name)); ctx = sc.upload(w, h); light = g.Light()
for k in range(n):
    v = orbit_view(w, h, 9 + k)
    ctx.enqueue(g.Camera(pos=v.pos, pitch=v.pitch, yaw=v.yaw), light); ctx.sync()
    print(k, ctx.last_frame_ms(), ctx.prune_stats())
