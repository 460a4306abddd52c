"""Per-GPU frame time of an N-way sharded frame, measured on ONE GPU: shard `rank` of `count` renders its own tiles into a local
framebuffer (csg_upload_shard), so T(count) here = what each GPU of an N-GPU box does per frame, minus NVLink.  Decomposes the
fixed part of a frame (pruning kernel latency, longest warp tile, launch) without spending N GPUs.
   python tools/gpu_shard_emul.py [frames]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import csg_b200 as g
import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
only = sys.argv[2] if len(sys.argv) > 2 else None
txt, _ = bench.scene_bytes()
cam, light = g.Camera(), g.Light()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for mode, name in ((1, "flat"), (2, "walk"), (0, "off")):
    if only and name != only:
        continue
    for count in (1, 2, 4, 8):
        for rank in sorted({0, count - 1}):
            sc = g.Scene.parse(txt)
            ctx = sc.upload_shard(bench.WIDTH, bench.HEIGHT, 0, rank, count)
            ctx.set_pruning(mode)
            fb = torch.empty(bench.WIDTH * bench.HEIGHT * 4, dtype=torch.uint8, device="cuda")
            ctx.set_gather_target(fb.data_ptr())   # a plain local buffer: no start gate / join with the (absent) other shards
            ms = []
            for k in range(frames + 5):
                flush.zero_()
                torch.cuda.synchronize()
                ctx.enqueue(cam, light)
                ctx.sync()
                if k >= 5:
                    ms.append(ctx.last_frame_ms())
            ctx.set_view_cache(True)      # trees kept: the frame kernel alone
            fk = []
            for k in range(frames // 2 + 3):
                flush.zero_()
                torch.cuda.synchronize()
                ctx.enqueue(cam, light)
                ctx.sync()
                if k >= 3:
                    fk.append(ctx.last_frame_ms())
            out[f"{name}/{count}/rank{rank}"] = {"frame_ms": round(float(np.median(ms)), 4), "min": round(float(np.min(ms)), 4),
                                                 "frame_kernel_only_ms": round(float(np.median(fk)), 4)}
            print(name, count, f"rank{rank}", out[f"{name}/{count}/rank{rank}"], flush=True)
            ctx.close(); sc.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "shard_emul.json"), "w"), indent=1)
