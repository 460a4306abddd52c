import os, sys
import numpy as np
sys.path.insert(0, "oracle"); sys.path.insert(0, ".")
import csg_b200 as g
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for opt in (1, 0):
    sc = g.Scene.parse(g.Scene.generate_text(4096, 1234), optimize=opt)
    ctx = sc.upload(7680, 4320)
    ctx.set_supersampling(4)
    cam, light = g.Camera(), g.Light()
    ms = []
    for k in range(6):
        flush.zero_(); torch.cuda.synchronize()
        ctx.enqueue(cam, light); ctx.sync(); ms.append(ctx.last_frame_ms())
    print("optimize", opt, "configs[4] frame ms:", [round(m, 3) for m in ms], ctx.prune_stats(), ctx.info(), flush=True)
    ctx.close(); sc.close()
