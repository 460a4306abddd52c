"""Development check run on a GPU box: CUDA path vs the reference CUDA kernel itself (oracle/_ref/libref_gpu.so),
on the whole scene corpus, plus kernel timings of both.  Not part of the product or the test suite."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import csg_b200 as g  # noqa: E402
from oracle_py import RefGPU, View, oblique_view, orbit_view, scene_text, SCENES_DIR  # noqa: E402


def cam_of(v):
    return g.Camera(pos=v.pos, pitch=v.pitch, yaw=v.yaw, fov=v.fov)


def light_of(v):
    return g.Light() if v.polar > 1e9 else g.Light(v.polar, v.azimuth)


def compare(name, v, ref, optimize, timing_iters=0):
    txt = g.Scene.generate_text(int(name.split(":")[1]), 1234) if name.startswith("synthetic:") else scene_text(name)
    fr = ref.render(txt, v, warmup=1 if timing_iters else 0, iters=max(1, timing_iters), shipped=bool(timing_iters))
    sc = g.Scene.parse(txt, optimize=optimize)
    ctx = sc.upload(v.width, v.height)
    cam, light = cam_of(v), light_of(v)
    hit, prim, t = ctx.render_aov(cam)
    f32 = ctx.render_f32(cam, light).reshape(-1)
    rgba8 = ctx.render(cam, light).reshape(-1)
    n = hit.size
    hm = int((hit != fr.hit).sum())
    both = (hit == 1) & (fr.hit == 1)
    pm = int((both & (prim != fr.prim)).sum())
    same = both & (prim == fr.prim)
    tb = int((t[same].view(np.uint32) != fr.t[same].view(np.uint32)).sum())
    rel = np.abs(t[same] - fr.t[same]) / np.maximum(np.abs(fr.t[same]), 1e-30)
    ref8 = fr.rgba8()
    q8 = (np.clip(f32, 0, 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    cd_f32 = int(np.abs(q8.astype(int) - ref8.astype(int)).max())
    cd_u8 = int(np.abs(rgba8.astype(int) - ref8.astype(int)).max())
    nb = int((np.abs(rgba8.astype(int) - ref8.astype(int)).reshape(-1, 4).max(axis=1) > 1).sum())
    out = dict(scene=name, w=v.width, h=v.height, opt=optimize, hit_mismatch=hm, prim_mismatch=pm, t_bits_differ=tb,
               max_rel_t=float(rel.max()) if rel.size else 0.0, max_rgba8_diff=cd_u8, max_f32q_diff=cd_f32,
               px_over_1lsb=nb, hits=int(fr.hit.sum()), agree=1.0 - (hm + pm) / n)
    if timing_iters:
        ms = []
        for _ in range(3):
            ctx.enqueue(cam, light)
            ctx.sync()
        for _ in range(timing_iters):
            ctx.enqueue(cam, light)
            ctx.sync()
            ms.append(ctx.last_frame_ms())
        out.update(ms_ours=float(np.median(ms)), ms_ref_kernels=float(np.median(fr.ms_kernels)),
                   ms_ref_shipped=float(np.median(fr.ms_shipped)), info=ctx.info())
        out["speedup_vs_ref_kernels"] = out["ms_ref_kernels"] / out["ms_ours"]
    ctx.close()
    sc.close()
    return out


def main():
    ref = RefGPU()
    res = []
    W, H = 640, 360
    names = sorted(fn[:-4] for fn in os.listdir(SCENES_DIR) if fn.endswith(".txt"))
    for name in names:
        views = [View(W, H), oblique_view(W, H) if "Cheese" in name else orbit_view(W, H, 7)]
        for v in views:
            for opt in (0, 1):
                r = compare(name, v, ref, opt)
                res.append(r)
                print(json.dumps(r), flush=True)
    for name, v in [("testWikipedia", View(1920, 1080)), ("testSphereCutByCubesAndCylinder", orbit_view(3840, 2160, 9)),
                    ("testCheese256", View(3840, 2160)), ("testCheese512", View(3840, 2160)),
                    ("testCheese512", oblique_view(3840, 2160)), ("synthetic:4096", View(3840, 2160))]:
        for opt in (0, 1):
            r = compare(name, v, ref, opt, timing_iters=10)
            res.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
