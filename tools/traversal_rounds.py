"""CPU study of the frame kernel's traversal on per-tile trees (no GPU): the Python model of traverse() (tests/test_traversal_model.py)
run on the trees the prefix-sum pruning produces for the 64x32-pixel tiles of the bench frame (Cheese512 @ 3840x2160), on a sparse
sample of its rays.  Prints rounds per ray by state, which can be held against tools/gpu_stats.py (the kernel's own counters), and
is the place to try a change of the traversal against the reference semantics before a GPU is involved.
   python tools/traversal_rounds.py [pixel step, default 24]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
import csg_b200 as g
from oracle_py import Oracle, View, scene_text
import test_traversal_model as tm
import test_flat_tree as ft

STEP = int(sys.argv[1]) if len(sys.argv) > 1 else 24
W, H = 3840, 2160
orc = Oracle()
txt = scene_text("testCheese512")
sc = g.Scene.parse(txt, optimize=1)
rec, par, depth = sc.flatten()
_, prims = sc.dump()
sc.close()
rec = rec.reshape(-1, 8)
meta = rec[:, 7].astype(np.int64)
kind = meta & 7
end = ft.subtree_ends(kind, meta)
view = View(W, H)
cam = orc.camera(view)
pos = np.array([cam.x, cam.y, cam.z], np.float64)
tan_half = float(np.tan(np.float32(cam.fov) * np.float32(0.5)))
leaf_nodes = np.nonzero(kind >= 3)[0]
boxes = {int(n): ft.cull_box(rec[n]) for n in leaf_nodes}


def pruned_tree(mx, my):
    """Records of the tile's tree (world space, boxes and flags refitted), or None when nothing is reachable."""
    return tm.pruned_tile_tree(rec, view, cam, tan_half, mx * 64, my * 32, min(mx * 64 + 64, W), min(my * 32 + 32, H))


class Counting(tm.TileModel):
    def __init__(self, *a):
        super().__init__(*a)
        self.n = collections.Counter()

    def eval_child(self, c, d, dd, tmin, gated):
        self.n["leaf" if (int(self.meta[c]) & 7) >= 3 else "box"] += 1
        return super().eval_child(c, d, dd, tmin, gated)


trees = {}
stats = []
out3 = (C.c_float * 3)()
ref_rows = {}
for y in range(STEP // 2, H, STEP):
    for x in range(STEP // 2, W, STEP):
        key = (x // 64, y // 32)
        if key not in trees:
            t = pruned_tree(*key)
            trees[key] = None if t is None else Counting(orc, t, np.asarray(prims).reshape(-1, 48), view.pos)
        model = trees[key]
        if model is None:
            stats.append((0, 0, 0, 0))
            continue
        orc.lib.orc_raygen(C.byref(cam), W, H, x, y, C.c_float(tan_half), out3)
        model.n.clear()
        hit = model.traverse((float(out3[0]), float(out3[1]), float(out3[2])))
        stats.append((model.n["leaf"], model.n["box"], 0 if hit.miss else 1, len(model.meta)))
a = np.array(stats)
hit = a[:, 2] == 1
print(f"rays {len(a)}, hit {hit.mean():.3f}; tiles with a tree {sum(1 for t in trees.values() if t is not None)} of {len(trees)}, "
      f"mean nodes per tile tree {np.mean([len(t.meta) for t in trees.values() if t is not None]):.1f}")
print(f"leaf tests per ray {a[:, 0].mean():.2f} (hit rays {a[hit, 0].mean():.2f}, max {a[:, 0].max()}); box tests per ray {a[:, 1].mean():.2f} "
      f"(hit rays {a[hit, 1].mean():.2f}, max {a[:, 1].max()})")
