"""BASELINE.json configs[1]: testSphereCutByCubesAndCylinder @ 3840x2160, 64-frame orbit, one B200: frame-by-frame vs csg_render_batch."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import csg_b200 as g
from oracle_py import scene_text, orbit_view

W, H, N = 3840, 2160, 64
for name in ["testSphereCutByCubesAndCylinder", "testCheese512"]:
    sc = g.Scene.parse(scene_text(name))
    ctx = sc.upload(W, H)
    light = g.Light()
    tgt = (0.0, 0.0, -20.0) if "Cheese" in name else (0.0, 0.0, 0.0)
    rad = 30.0 if "Cheese" in name else 5.0
    views = [orbit_view(W, H, k, n=N, radius=rad, target=tgt) for k in range(N)]
    cams = [g.Camera(pos=v.pos, pitch=v.pitch, yaw=v.yaw) for v in views]
    dev = torch.empty(N * W * H * 4, dtype=torch.uint8, device="cuda")
    host = torch.empty(N * W * H * 4, dtype=torch.uint8).pin_memory()
    res = {"scene": name, "frames": N}
    for _ in range(2):   # second round = warm
        ms = []
        for c in cams:
            ctx.enqueue(c, light); ctx.sync(); ms.append(ctx.last_frame_ms())
        res["frame_by_frame_ms_per_frame"] = float(np.mean(ms))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.render_batch(cams, light, dev.data_ptr())
        torch.cuda.synchronize(); t1 = time.perf_counter()
        res["batch_device_ms_per_frame"] = ctx.last_frame_ms() / N
        res["batch_device_wall_ms_per_frame"] = (t1 - t0) * 1e3 / N
        t0 = time.perf_counter()
        ctx.render_batch(cams, light, host.data_ptr())
        t1 = time.perf_counter()
        res["batch_to_pinned_host_ms_per_frame"] = (t1 - t0) * 1e3 / N
    # equality with frame-by-frame
    ref = ctx.render(cams[5], light).copy().reshape(-1)
    n1 = W * H * 4
    res["frame5_identical"] = bool(np.array_equal(ref, host[5 * n1:6 * n1].numpy())) and bool(np.array_equal(ref, dev[5 * n1:6 * n1].cpu().numpy()))
    res["Mrays_per_s_batch_device"] = W * H / res["batch_device_ms_per_frame"] / 1e3
    print(json.dumps(res), flush=True)
    ctx.close()
