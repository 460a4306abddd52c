"""Per-warp-tile traversal rounds (max over the tile's 32 rays, by state) of the bench frame with and without flat evaluation;
saves gpurun_out/tile_iters.npz for study on the CPU (which tiles are the heavy ones, and what the model says they do)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g
from oracle_py import scene_text
W, H = 3840, 2160
out = {}
for fl in (0, 30):
    os.environ["CSG_B200_FLAT_LEAVES"] = str(fl)
    sc = g.Scene.parse(scene_text("testCheese512"), optimize=1)
    ctx = sc.upload(W, H)
    packed = ctx.render_stats(g.Camera()).reshape(H, W)
    for name, a in (("search", packed >> 20), ("enter", (packed >> 10) & 1023), ("other", packed & 1023)):
        out[f"{name}{fl}"] = a.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3)).astype(np.int16)
    tot = (packed >> 20) + ((packed >> 10) & 1023) + (packed & 1023)
    t = tot.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
    out[f"total{fl}"] = t.astype(np.int16)
    print("flat_leaves", fl, "warp rounds total", int(t.sum()), "max", int(t.max()), "at (tile y, x)", np.unravel_index(np.argmax(t), t.shape),
          "pixel rounds max", int(tot.max()), "at", np.unravel_index(np.argmax(tot), tot.shape), flush=True)
    ctx.close(); sc.close()
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tile_iters.npz"), **out)
