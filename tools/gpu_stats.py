"""Work statistics of our traversal (loop iterations per pixel) on the bench workload; run on a GPU box."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import csg_b200 as g  # noqa: E402
from oracle_py import scene_text  # noqa: E402

W, H = 3840, 2160
out = {}
for name in ["testCheese512"]:
    for opt in (0, 1):
        sc = g.Scene.parse(scene_text(name), optimize=opt)
        ctx = sc.upload(W, H)
        cam = g.Camera()
        packed = ctx.render_stats(cam).reshape(H, W)
        it_s, it_g, it_o = packed >> 20, (packed >> 10) & 1023, packed & 1023
        it = it_s + it_g + it_o
        hit, _, _ = ctx.render_aov(cam)
        # warp tiles are 8x4
        tiles = it.reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
        tmax = tiles.max(axis=1)
        out[f"{name}/opt{opt}"] = dict(mean_iters_per_ray=float(it.mean()), max_iters=int(it.max()),
                                       mean_iters_hit_rays=float(it[hit.reshape(H, W) == 1].mean()),
                                       p50=float(np.percentile(it, 50)), p99=float(np.percentile(it, 99)),
                                       warp_iters_total=int(tmax.sum()), simt_efficiency_on_iters=float(it.sum() / (32.0 * tmax.sum())),
                                       tiles_with_work=int((tmax > 2).sum()), tiles=int(tmax.size),
                                       search_visits=float(it_s.mean()), frame_visits=float(it_g.mean()), other_iters=float(it_o.mean()),
                                       tile_cost_hist=np.histogram(tmax[tmax > 0], bins=[1, 4, 8, 16, 32, 64, 128, 256, 512, 100000])[0].tolist(),
                                       tile_cost_sum_by_bin=[int(tmax[(tmax >= a) & (tmax < b)].sum()) for a, b in
                                                             zip([1, 4, 8, 16, 32, 64, 128, 256, 512], [4, 8, 16, 32, 64, 128, 256, 512, 100000])])
        if name == "testCheese512" and opt == 1:
            np.save(os.path.join(ROOT, "gpurun_out", "iters_c512.npy"), packed[::4, ::4].astype(np.int32))
        print(name, opt, json.dumps(out[f"{name}/opt{opt}"]), flush=True)
        ctx.close()
with open(os.path.join(ROOT, "gpurun_out", "stats.json"), "w") as f:
    json.dump(out, f, indent=1)
