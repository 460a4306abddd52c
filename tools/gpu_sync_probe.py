"""Where the time of a sharded frame goes between the GPUs (instrumented build libcsg_b200_probe.so, in-process multi-GPU context):
every device's own globaltimer at the gate of its two kernels, when its last CTA has finished the shard's own work, and when the
join is over.  Differences are taken per device (the devices' timers are not assumed to agree).
   CSG_B200_LIB=cuda-csg-tree-raycasting_b200/libcsg_b200_probe.so python tools/gpu_sync_probe.py [n_gpus]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import csg_b200 as g
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
txt, _ = bench.scene_bytes()
sc = g.Scene.parse(txt); ctx = sc.upload(bench.WIDTH, bench.HEIGHT, n_gpus=n)
cam, light = g.Camera(), g.Light()
flush = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
g.lib.csg_debug_sync_probe.argtypes = [C.c_int, C.c_void_p]
rows, ms = [], []
for k in range(12):
    for d in range(n):
        flush[d].zero_()
    for d in range(n):
        torch.cuda.synchronize(d)
    ctx.enqueue(cam, light); ctx.sync()
    if k >= 4:
        ms.append(ctx.last_frame_ms())
        r = []
        for d in range(n):
            buf = np.zeros(8, np.uint64)
            assert g.lib.csg_debug_sync_probe(d, buf.ctypes.data_as(C.c_void_p)) == 0
            r.append(buf.astype(np.int64))
        rows.append(r)
a = np.array(rows, dtype=np.float64) / 1e3     # [frame, device, slot] in us
print(f"{n} GPUs in one process: frame (root's events) {np.mean(ms)*1e3:.1f} us (min {np.min(ms)*1e3:.1f})")
names = ["prune gate wait", "prune gate -> frame gate open", "frame gate wait", "frame gate -> own work done (last CTA)", "join (root: wait for peers; peer: fence + flag)"]
for d in range(n):
    t = a[:, d, :]
    parts = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]]
    print(f"device {d}: " + "; ".join(f"{nm} {p.mean():.1f}" for nm, p in zip(names, parts)) + f"; prune gate -> join done {np.mean(t[:, 5] - t[:, 1]):.1f} us")
# cross-device view, if the timers agree to within a few us (printed for what it is worth): peers' gate-open time after the root's
print("peer prune-gate-open minus root prune-gate-open (raw timers):", [round(float(np.mean(a[:, d, 1] - a[:, 0, 1])), 1) for d in range(n)])
print("peer own-work-done minus root join-done (raw timers):", [round(float(np.mean(a[:, d, 4] - a[:, 0, 5])), 1) for d in range(n)])
