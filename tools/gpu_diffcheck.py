"""Diagnostic: where do float colour differences against the reference CUDA kernel come from?  Needs a shading-debug build
(CSG_B200_LIB=build/variants/libcsg_b200_dbg.so) that returns intermediate vectors through csg_render_f32."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import csg_b200 as g
import oracle_py
from oracle_py import RefGPU, View

import sys as _s
txt = g.Scene.generate_text(4096, 1234)
v = View(3840, 2160)
rg = RefGPU()
ref = rg.render(txt, v)
rg.lib.refgpu_render_details.argtypes = [C.c_char_p, C.POINTER(oracle_py.RefView), C.c_void_p, C.c_char_p, C.c_int]
det = np.zeros((v.width * v.height, 6), np.float32)
err = C.create_string_buffer(256)
rv = v.ref()
assert rg.lib.refgpu_render_details(txt, C.byref(rv), det.ctypes.data_as(C.c_void_p), err, 256) == 0
lib = C.CDLL(os.environ["CSG_B200_LIB"])
sc = g.Scene.parse(txt); ctx = sc.upload(v.width, v.height)
cam, light = g.Camera(), g.Light()
out = {}
for mode in range(6):
    lib.csg_dbg_mode(mode)
    out[mode] = ctx.render_f32(cam, light).reshape(-1, 4).copy()
lib.csg_dbg_mode(0)
hit, prim, t = ctx.render_aov(cam)
rf = ref.rgba.reshape(-1, 4)
h = hit == 1
col_bad = h & (np.abs(out[0] - rf).max(axis=1) > 0)
both_nan = np.isnan(out[1][:, :3]) & np.isnan(det[:, 3:6])
nrm_bad = h & ((out[1][:, :3].view(np.uint32) != det[:, 3:6].view(np.uint32)) & ~both_nan & ~((out[1][:, :3] == 0) & (det[:, 3:6] == 0))).any(axis=1)
pos_bad = h & (out[2][:, :3].view(np.uint32) != det[:, 0:3].view(np.uint32)).any(axis=1)
print("hit pixels", int(h.sum()), "colour differs", int(col_bad.sum()), "normal differs", int(nrm_bad.sum()), "position differs", int(pos_bad.sum()))
print("colour differs but normal and position equal:", int((col_bad & ~nrm_bad & ~pos_bad).sum()))
print("normal differs but colour equal:", int((nrm_bad & ~col_bad).sum()))
nodes, _ = sc.dump()
rows = nodes.view(np.int32).reshape(-1, 11)
ptype = {int(r[1]): int(r[0]) for r in rows if r[1] >= 0}
import collections
print("normal-diff by type:", dict(collections.Counter(ptype[int(prim[i])] for i in np.nonzero(nrm_bad)[0])))
print("position-diff by type:", dict(collections.Counter(ptype[int(prim[i])] for i in np.nonzero(pos_bad)[0])))
for i in np.nonzero(nrm_bad)[0][:4]:
    print("n ours", out[1][i, :3], "ref", det[i, 3:6], "pos ours", out[2][i, :3], "ref", det[i, :3], "type", ptype[int(prim[i])], "t", t[i])
for i in np.nonzero(col_bad & ~nrm_bad & ~pos_bad)[0][:4]:
    print("clean-input colour diff: ours", out[0][i], "ref", rf[i], "diff,sb,spec,k", out[5][i], "type", ptype[int(prim[i])])
