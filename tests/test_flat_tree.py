"""CPU tests of the flattened GPU tree (csg_scene_flatten) and of the claim behind per-tile pruning, the latter
checked with the oracle: dropping the primitives a tile's frustum cannot reach and collapsing the operators left with one
operand does not change a single pixel of that tile."""
import numpy as np
import pytest

import scenes
from oracle_py import View, orbit_view

K_UNION, K_DIFF, K_INTER, K_SPHERE, K_CYL, K_CUBE = range(6)
LEFT_LEAF, RIGHT_LEAF, BOUNDED, PURE = 1 << 3, 1 << 4, 1 << 5, 1 << 6


def flat(csg, txt, optimize):
    sc = csg.Scene.parse(txt, optimize=optimize)
    rec, par, depth = sc.flatten()
    nn, npr, _ = sc.counts()
    sc.close()
    return rec, par, depth, nn, npr


def cull_box(rec_row):
    f = rec_row[:7].view(np.float32)
    kind = int(rec_row[7]) & 7
    if kind == K_SPHERE:   # float32, operation for operation like leaf_cull_box (csg_scene.cpp)
        c, r = f[0:3], np.abs(f[3])
        pad = r * np.float32(1e-4) + np.abs(c) * np.float32(4e-7) + np.float32(1e-30)
        return (c - r - pad).astype(np.float64), (c + r + pad).astype(np.float64)
    return f[0:3].astype(np.float64), f[3:6].astype(np.float64)


@pytest.mark.parametrize("optimize", [0, 1])
@pytest.mark.parametrize("scene_id", scenes.all_scene_ids() + ["synthetic:200"])
def test_flattened_tree_invariants(scene_id, optimize, csg):
    txt = csg.Scene.generate_text(200, seed=11) if scene_id.startswith("synthetic:") else scenes.text_of(scene_id)
    rec, par, depth, nn, npr = flat(csg, txt, optimize)
    n = len(rec)
    assert n == nn == 2 * npr - 1
    meta = rec[:, 7]
    kind = meta & 7
    assert par[0] == -1
    seen_prims = sorted(int(m >> 8) for m, k in zip(meta, kind) if k >= 3)
    assert seen_prims == list(range(npr))                                   # every primitive exactly once, ids kept
    info = {}

    def walk(i, d):   # returns (end index, pure, bounded, operator depth)
        k = int(kind[i])
        if k >= 3:
            info[i] = (k in (K_SPHERE, K_CUBE), k != K_CYL)
            return i + 1, info[i][0], info[i][1], d
        l, r = i + 1, int(meta[i] >> 8)
        assert par[l] == i and par[r] == i
        end_l, pl, bl, dl = walk(l, d + 1)
        assert end_l == r                                                   # preorder: the right child follows the left subtree
        end_r, pr, br, dr = walk(r, d + 1)
        assert bool(meta[i] & LEFT_LEAF) == (kind[l] >= 3) and bool(meta[i] & RIGHT_LEAF) == (kind[r] >= 3)
        pure, bounded = (k == K_UNION and pl and pr), (bl and br)
        assert bool(meta[i] & PURE) == pure and bool(meta[i] & BOUNDED) == bounded
        lo, hi = cull_box(rec[i])
        (llo, lhi), (rlo, rhi) = cull_box(rec[l]), cull_box(rec[r])
        if k == K_UNION:
            assert (lo <= np.minimum(llo, rlo)).all() and (hi >= np.maximum(lhi, rhi)).all()
        elif k == K_DIFF:
            assert (lo <= llo).all() and (hi >= lhi).all()
        else:
            assert ((lo <= llo).all() and (hi >= lhi).all()) or ((lo <= rlo).all() and (hi >= rhi).all())
        return end_r, pure, bounded, max(dl, dr)
    end, _, _, dmax = walk(0, 0)
    assert end == n
    assert depth == dmax if n > 1 else depth == 0


# ---- the pruning claim, on the oracle ----------------------------------------------------------------------------------
ARGS = {"Sphere": 5, "Cube": 5, "Cylinder": 9}


def parse_tokens(text):
    toks = text.decode().split() if isinstance(text, bytes) else text.split()
    pos = 0
    leaves = []

    def rec():
        nonlocal pos
        kw = toks[pos]
        pos += 1
        if kw in ARGS:
            node = ("leaf", kw, toks[pos:pos + ARGS[kw]], len(leaves))
            leaves.append(node)
            pos += ARGS[kw]
            return node
        left = rec()
        right = rec()
        return ("op", kw, left, right)
    root = rec()
    assert pos == len(toks)
    return root, leaves


def collapse(node, alive):
    """The tile's tree: None if it can only miss."""
    if node[0] == "leaf":
        return node if alive[node[3]] else None
    _, kw, l, r = node
    a, b = collapse(l, alive), collapse(r, alive)
    if kw == "Union":
        return ("op", kw, a, b) if a and b else (a or b)
    if kw == "Difference":
        return None if not a else (("op", kw, a, b) if b else a)
    return ("op", kw, a, b) if a and b else None


def emit(node, out, ids):
    if node[0] == "leaf":
        ids.append(node[3])
        out.append(node[1] + " " + " ".join(node[2]))
    else:
        out.append(node[1])
        emit(node[2], out, ids)
        emit(node[3], out, ids)


def tile_planes(view, cam, tan_half, x0, y0, x1, y1):
    """Inward normals of the tile's frustum (pixel range + 1 pixel of margin), as csg_prune_kernel builds them."""
    w, h = view.width, view.height
    fwd, right, up = (np.array(cam.forward, np.float64), np.array(cam.right, np.float64), np.array(cam.up, np.float64))
    corners = []
    for fx, fy in [(x0 - 1, y0 - 1), (x1 + 1, y0 - 1), (x1 + 1, y1 + 1), (x0 - 1, y1 + 1)]:
        u, v = fx / (w - 1), fy / (h - 1)
        nx, ny = (w / h) * (2 * u - 1) * tan_half, (1 - 2 * v) * tan_half
        corners.append(fwd + right * nx + up * ny)
    dc = sum(corners)
    planes = []
    for c in range(4):
        nrm = np.cross(corners[c], corners[(c + 1) & 3])
        planes.append(nrm if nrm @ dc >= 0 else -nrm)
    planes.append(fwd)
    return planes


@pytest.mark.parametrize("scene_id", ["inline:nested", "inline:deep_left_chain", "inline:rotated_cylinder_union", "corpus:testWikipediaMult",
                                      "corpus:testCubeCutEdges", "synthetic:60"])
def test_pruned_tile_tree_gives_the_same_pixels(scene_id, csg, oracle):
    if scene_id.startswith("corpus:") and scene_id[7:] not in scenes.corpus_names():
        pytest.skip("scene corpus not staged")
    txt = csg.Scene.generate_text(60, seed=5) if scene_id.startswith("synthetic:") else scenes.text_of(scene_id)
    root, leaves = parse_tokens(txt)
    rec, _, _, _, npr = flat(csg, txt, 0)
    boxes = {int(m >> 8): cull_box(r) for r, m in zip(rec, rec[:, 7]) if (int(m) & 7) >= 3}
    w, h, tw, th = 96, 64, 32, 16
    for view in (View(w, h), orbit_view(w, h, 11, radius=6.0, pitch_deg=25.0)):
        if scene_id.startswith("synthetic:"):
            view = View(w, h, pos=(0.0, 0.0, 5.0)) if view.pitch == 0 else View(w, h, pos=(30.0, 10.0, -10.0), pitch=-0.2, yaw=1.2)
        full = oracle.render(txt, view)
        cam = oracle.camera(view)
        pos = np.array([cam.x, cam.y, cam.z], np.float64)
        tan_half = float(np.tan(np.float32(cam.fov) * np.float32(0.5)))
        pruned_somewhere = False
        for ty in range(0, h, th):
            for tx in range(0, w, tw):
                planes = tile_planes(view, cam, tan_half, tx, ty, tx + tw, ty + th)
                alive = []
                for k in range(npr):
                    lo, hi = boxes[k][0] - pos, boxes[k][1] - pos
                    outside = any((np.maximum(p * lo, p * hi)).sum() < 0 for p in planes)
                    alive.append(not outside)
                pruned_somewhere |= not all(alive)
                tree = collapse(root, alive)
                sl = np.s_[ty:ty + th, tx:tx + tw]
                fh, ft, fp = (full.hit.reshape(h, w)[sl], full.t.reshape(h, w)[sl], full.prim.reshape(h, w)[sl])
                if tree is None:
                    assert not fh.any(), f"{scene_id}: tile ({tx},{ty}) has hits but its pruned tree is empty"
                    continue
                out, ids = [], []
                emit(tree, out, ids)
                if tree[0] == "leaf" and tree[1] == "Cylinder" and root[0] == "op":
                    # A cylinder that became the root by collapse is still reached through an operator in the scene, so it keeps
                    # the reference's non-conservative gating box (Q6) — a root primitive would be intersected without it (Q7).
                    # The kernel carries a flag for this (root_gated); the oracle needs an operator: a Union with a sphere no ray hits.
                    out = ["Union"] + out + ["Sphere 1e6 1e6 1e6 000000 0.001"]
                part = oracle.render("\n".join(out).encode(), view, rows=(ty, ty + th))
                ph, pt, pp = (part.hit.reshape(h, w)[sl], part.t.reshape(h, w)[sl], part.prim.reshape(h, w)[sl])
                assert np.array_equal(fh, ph), f"{scene_id}: hit mask differs in tile ({tx},{ty})"
                m = fh == 1
                assert np.array_equal(ft[m].view(np.uint32), pt[m].view(np.uint32))            # every bit of t
                assert np.array_equal(fp[m], np.array(ids, np.int32)[pp[m]])                 # same primitive (ids renumbered)
                assert np.array_equal(full.rgba8().reshape(h, w, 4)[sl], part.rgba8().reshape(h, w, 4)[sl])
        assert pruned_somewhere


# ---- csg_prune_flat_kernel's arithmetic: the tile's tree out of prefix sums over the preorder layout ---------------------
def subtree_ends(kind, meta):
    n = len(kind)
    end = np.zeros(n, np.int64)

    def walk(i):
        if kind[i] >= 3:
            end[i] = i + 1
        else:
            walk(i + 1)
            end[i] = walk(int(meta[i] >> 8))
        return int(end[i])
    assert walk(0) == n
    return end


def collapse_flat(kind, meta, alive):
    """Recursive definition: returns the surviving nodes as a nested tuple, or None."""
    def rec(i):
        k = int(kind[i])
        if k >= 3:
            return (i,) if alive[i] else None
        a, b = rec(i + 1), rec(int(meta[i] >> 8))
        if k == K_UNION:
            return (i, a, b) if a and b else (a or b)
        if k == K_DIFF:
            return None if not a else ((i, a, b) if b else a)
        return (i, a, b) if a and b else None
    return rec(0)


def preorder_records(tree):
    """[(node, right operand's record index or -1)] in preorder."""
    out = []

    def rec(t):
        me = len(out)
        out.append([t[0], -1])
        if len(t) == 3:
            rec(t[1])
            out[me][1] = len(out)
            rec(t[2])
    if tree:
        rec(tree)
    return [tuple(x) for x in out]


def prune_by_prefix_sums(kind, meta, end, alive, par=None):
    """What csg_prune_flat_kernel does, step for step, in numpy.  par = the parent array the kernel walks when it clears the
    primitives below operators that are gone (None: clear the preorder range of every such operator, the definition)."""
    n = len(kind)
    is_leaf = kind >= 3
    right = (meta >> 8).astype(np.int64)
    flags = np.where(is_leaf, alive, 0).astype(np.int64)
    rounds = 0
    while True:
        rounds += 1
        A = np.cumsum(flags)
        ops = np.nonzero(~is_leaf)[0]
        hl = np.zeros(n, bool)
        hr = np.zeros(n, bool)
        hl[ops] = A[right[ops] - 1] != A[ops]
        hr[ops] = A[end[ops] - 1] != A[right[ops] - 1]
        surv = np.where(is_leaf, flags, (hl & hr).astype(np.int64))
        gone = (~is_leaf) & (((kind == K_DIFF) & ~hl & hr) | ((kind == K_INTER) & (hl != hr)))
        if not gone.any():
            break
        if par is None:
            for i in np.nonzero(gone)[0]:
                flags[i + 1:end[i]] = 0
        else:
            # the kernel's two parallel passes: gone operators are marked, every primitive still alive walks up its ancestors
            for i in np.nonzero(is_leaf & (flags != 0))[0]:
                a = int(par[i])
                while a >= 0:
                    if gone[a]:
                        flags[i] = 0
                        break
                    a = int(par[a])
    S = np.cumsum(surv)
    recs = []
    for i in np.nonzero(surv)[0]:
        recs.append((int(i), -1 if is_leaf[i] else int(S[right[i] - 1])))
    return recs, rounds


@pytest.mark.parametrize("optimize", [0, 1])
@pytest.mark.parametrize("scene_id", scenes.all_scene_ids() + ["synthetic:300"])
def test_prefix_sum_pruning_equals_the_recursive_collapse(scene_id, optimize, csg):
    txt = csg.Scene.generate_text(300, seed=3) if scene_id.startswith("synthetic:") else scenes.text_of(scene_id)
    rec, par, depth, nn, npr = flat(csg, txt, optimize)
    meta = rec[:, 7].astype(np.int64)
    kind = meta & 7
    end = subtree_ends(kind, meta)
    rng = np.random.default_rng(17)
    saw_rounds = 0
    for trial in range(40):
        p = [0.0, 0.1, 0.5, 0.9, 1.0][trial % 5] if trial < 10 else rng.uniform(0.02, 0.98)
        alive = rng.uniform(size=len(kind)) < p
        want = preorder_records(collapse_flat(kind, meta, alive))
        got, rounds = prune_by_prefix_sums(kind, meta, end, alive)
        assert got == want
        got_walk, rounds_walk = prune_by_prefix_sums(kind, meta, end, alive, par)
        assert (got_walk, rounds_walk) == (got, rounds)   # clearing by the walk up the parents == clearing the preorder ranges
        saw_rounds = max(saw_rounds, rounds)
        # record i+1 is the left operand of operator record i; a record's subtree is contiguous
        for i, (node, ri) in enumerate(got):
            if ri >= 0:
                assert i + 1 < ri < len(got)
    if scene_id == "inline:nested":
        assert saw_rounds >= 2    # an Intersection / Difference that went away took primitives from its ancestors


def _operator_levels(rec):
    meta = rec[:, 7]
    kind = meta & 7

    def walk(i):
        if int(kind[i]) >= 3:
            return 0, [int(meta[i] >> 8)]
        dl, pl = walk(i + 1)
        dr, pr = walk(int(meta[i] >> 8))
        return 1 + max(dl, dr), pl + pr
    return walk(0)


@pytest.mark.parametrize("n_prims,seed", [(4096, 1234), (1000, 7), (200, 13), (37, 3)])
def test_rebalancing_respects_the_height_budget(n_prims, seed, csg):
    """Load-time re-balancing (csg_scene.cpp build_union) rebuilds Union subtrees spatially under a height budget: the frame
    kernel keeps one 16-byte frame per operator level in shared memory, so a tall tree costs resident warps (the 4096-primitive
    synthetic tree: 30 levels and 12 warps per SM with count-balanced splits, 12 levels and 24 warps now).  The rebuilt tree never
    needs more than 13 levels when the parsed one fits in 13, keeps every primitive, and only re-associates Unions."""
    import sys
    sys.setrecursionlimit(20000)
    txt = csg.Scene.generate_text(n_prims, seed=seed)
    rec0, _, depth0, nn, npr = flat(csg, txt, 0)
    rec1, _, depth1, _, _ = flat(csg, txt, 1)
    lv0, prims0 = _operator_levels(rec0)
    lv1, prims1 = _operator_levels(rec1)
    assert (lv0, lv1) == (depth0, depth1)
    assert sorted(prims0) == sorted(prims1) == list(range(npr))
    assert depth1 <= max(13, depth0), f"re-balanced tree has {depth1} levels, the parsed one {depth0}"
    kinds0 = np.bincount(rec0[:, 7] & 7, minlength=6)
    kinds1 = np.bincount(rec1[:, 7] & 7, minlength=6)
    assert (kinds0 == kinds1).all()                                         # same operators, same primitives


def test_rebalancing_keeps_a_chain_of_unions_short(csg):
    """A left-deep chain of 300 Unions over spheres (depth 300 as parsed) comes out balanced: ceil(log2 301) + slack levels."""
    import sys
    sys.setrecursionlimit(20000)
    leaves = [f"Sphere {0.1 * k:.3f} 0 0 FF0000 0.3" for k in range(301)]
    txt = ("Union\n" * 300 + leaves[0] + "\n" + "\n".join(leaves[1:]) + "\n").encode()
    rec0, _, depth0, _, _ = flat(csg, txt, 0)
    rec1, _, depth1, _, _ = flat(csg, txt, 1)
    assert depth0 == 300
    assert 9 <= depth1 <= 11
