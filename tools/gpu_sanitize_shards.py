"""Sharded frames for compute-sanitizer (no torch: starts in a second): every shard of a 4-way sharded Cheese512 frame renders into
one context's framebuffer — the widened traced rectangle, the pruning kernel's clearing passes (tiles that see the spheres but
not the cube) — and the result is compared with the unsharded frame.
   compute-sanitizer --tool racecheck python tools/gpu_sanitize_shards.py [width height count]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import csg_b200 as g
import bench
w, h, count = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (960, 540, 4)
txt, _ = bench.scene_bytes()
cam, light = g.Camera(), g.Light()
sc = g.Scene.parse(txt)
full = sc.upload(w, h)
full.enqueue(cam, light); full.sync()
want = full.read_framebuffer(np.empty((h, w, 4), np.uint8)).copy()
ptr = full.framebuffer()
full.enqueue(g.Camera(pos=(0.0, 0.0, 50.0)), light); full.sync()   # something else into the buffer before the shards fill it
for rank in range(count):
    ctx = sc.upload_shard(w, h, 0, rank, count)
    ctx.set_gather_target(ptr)
    ctx.enqueue(cam, light); ctx.sync()
    print("shard", rank, ctx.prune_stats(), flush=True)
    ctx.close()
got = full.read_framebuffer(np.empty((h, w, 4), np.uint8))
print("mismatching bytes:", int(np.count_nonzero(got != want)))
full.close(); sc.close()
print("done")
