"""Generates tests/golden/ref_gpu_golden.npz on a GPU box.

Runs the UNMODIFIED reference CUDA kernels (oracle/_ref/libref_gpu.so = the reference's own
RaycastKernel + LightningKernel built for sm_100 by oracle/Makefile) on the reference's scene corpus
and stores hit mask / primitive id / t / RGBA8 per (scene, view).  The fixtures pin both the C oracle
(tests -m "not gpu") and the CUDA path (tests -m gpu) to the reference itself.

    gpurun -- python tests/golden/make_golden.py         # writes gpurun_out/ref_gpu_golden.npz
    cp gpurun_out/ref_gpu_golden.npz tests/golden/
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
from oracle_py import RefGPU, View, oblique_view, orbit_view, scene_text, SCENES_DIR  # noqa: E402

W, H = 192, 108


def views_for(name):
    v = [("default", View(W, H))]
    if "Cheese" in name:
        v.append(("oblique", oblique_view(W, H)))
    else:
        v.append(("orbit5", orbit_view(W, H, 5)))
        v.append(("orbit23_fov60", orbit_view(W, H, 23, pitch_deg=35.0)))
        v[-1][1].fov = 60.0 * 3.14159 / 180.0
    return v


def view_params(v):
    return np.array([v.width, v.height, *v.pos, v.pitch, v.yaw, v.fov, v.polar, v.azimuth], dtype=np.float64)


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    ref = RefGPU()
    data = {}
    for fn in sorted(os.listdir(SCENES_DIR)):
        if not fn.endswith(".txt"):
            continue
        name = fn[:-4]
        txt = scene_text(name)
        for vname, v in views_for(name):
            fr = ref.render(txt, v)
            key = f"{name}/{vname}"
            data[key + "/view"] = view_params(v)
            data[key + "/hit"] = np.packbits(fr.hit)
            data[key + "/prim"] = fr.prim.astype(np.int16)
            data[key + "/t"] = fr.t
            data[key + "/rgba8"] = fr.rgba8()
            print(key, int(fr.hit.sum()))
    np.savez_compressed(os.path.join(out_dir, "ref_gpu_golden.npz"), **data)
    print("wrote", os.path.join(out_dir, "ref_gpu_golden.npz"))


if __name__ == "__main__":
    main()
