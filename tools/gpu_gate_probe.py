"""Why does a peer-first enqueue on ONE device time out?  Host timestamps around every call."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import csg_b200 as g
import scenes

txt = scenes.INLINE["nested"]
w, h = 257, 129
sc = g.Scene.parse(txt)
cam, light = g.Camera(), g.Light()
one = sc.upload(w, h); want = one.render(cam, light).copy(); one.close()
for order in ("root_first", "peer_first", "peer_first_prewarmed"):
    root, peer = sc.upload_shard(w, h, 0, 0, 2), sc.upload_shard(w, h, 0, 1, 2)
    peer.set_gather_root(root)
    if order == "peer_first_prewarmed":   # frame 0 in the safe order: tan cached, kernels loaded in both contexts
        root.enqueue(cam, light); peer.enqueue(cam, light); root.sync(); peer.sync()
    for frame in range(2):
        seq = (root, peer) if order == "root_first" else (peer, root)
        t = [time.perf_counter()]
        for c in seq:
            c.enqueue(cam, light); t.append(time.perf_counter())
        err = []
        for c in (root, peer):
            try:
                c.sync()
            except g.CsgError as e:
                err.append(str(e)[:40])
            t.append(time.perf_counter())
        ok = np.array_equal(root.read_framebuffer(), want)
        print(order, "frame", frame, "enqueue1 %.3f ms enqueue2 %.3f ms sync_root %.3f ms sync_peer %.3f ms" % tuple((b - a) * 1e3 for a, b in zip(t, t[1:])), "ok" if ok else "DIFF", err, flush=True)
    peer.close(); root.close()
