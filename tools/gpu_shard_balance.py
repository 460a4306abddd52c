"""Every shard of an N-way sharded frame on ONE GPU, one after the other, all rendering into one buffer: each shard's frame time
(the balance of the tile hand-out) and whether the shards together produce the single-GPU frame.
   CSG_B200_LIB=... python tools/gpu_shard_balance.py [frames] [counts, comma separated]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csg_b200 as g
import bench

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 15
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8]
txt, _ = bench.scene_bytes()
cam, light = g.Camera(), g.Light()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sc = g.Scene.parse(txt)
one = sc.upload(bench.WIDTH, bench.HEIGHT)
one.enqueue(cam, light); one.sync()
ref = torch.from_numpy(one.read_framebuffer(np.empty((bench.HEIGHT, bench.WIDTH, 4), np.uint8)).copy()).cuda().reshape(-1)
one.close()
out = {"lib": os.path.basename(os.environ.get("CSG_B200_LIB", "shipped"))}
for count in counts:
    fb = torch.zeros(bench.WIDTH * bench.HEIGHT * 4, dtype=torch.uint8, device="cuda")
    per = []
    for rank in range(count):
        ctx = sc.upload_shard(bench.WIDTH, bench.HEIGHT, 0, rank, count)
        ctx.set_gather_target(fb.data_ptr())
        ms = []
        for k in range(frames + 4):
            flush.zero_(); torch.cuda.synchronize()
            ctx.enqueue(cam, light); ctx.sync()
            if k >= 4:
                ms.append(ctx.last_frame_ms())
        per.append(round(float(np.median(ms)) * 1e3, 2))
        ctx.close()
    torch.cuda.synchronize()
    out[f"shards_{count}"] = {"per_shard_us": per, "max_us": max(per), "mean_us": round(float(np.mean(per)), 2),
                              "mismatching_bytes": int((fb != ref).sum().item())}
sc.close()
print(json.dumps(out))
