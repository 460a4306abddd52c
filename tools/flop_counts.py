"""Algorithmic flop/ray of the REFERENCE algorithm per (scene, view, resolution) — SURVEY.md §8(d).

Runs the instrumented C oracle (event counts of the reference's state machine: AABB tests, primitive tests,
Compute steps) and applies the per-event weights of §8(d):
  AABB 26, sphere 36, cube 60, cylinder 80, Compute 10, ray-gen 40 per ray, hit details + Phong 130 per hit pixel.
Writes profiles/flop_per_ray.json, which bench.py reads for roofline.achieved."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oracle_py import Oracle, View, oblique_view, orbit_view, scene_text  # noqa: E402

W = dict(aabb=26, sphere=36, cube=60, cylinder=80, compute_calls=10)


def flop_per_ray(c):
    n = c["rays"]
    f = sum(c[k] * w for k, w in W.items()) + 40 * n + 130 * c["hits"]
    return f / n


def main():
    orc = Oracle()
    cfgs = {
        "testWikipedia@1920x1080/default": ("testWikipedia", View(1920, 1080)),
        "testSphereCutByCubesAndCylinder@3840x2160/orbit0": ("testSphereCutByCubesAndCylinder", orbit_view(3840, 2160, 0)),
        "testCheese256@3840x2160/default": ("testCheese256", View(3840, 2160)),
        "testCheese512@3840x2160/default": ("testCheese512", View(3840, 2160)),
        "testCheese512@3840x2160/oblique": ("testCheese512", oblique_view(3840, 2160)),
    }
    out = {}
    for key, (name, v) in cfgs.items():
        fr = orc.render(scene_text(name), v, want_rgba=False)
        c = fr.counters
        n = c["rays"]
        out[key] = dict(flop_per_ray=round(flop_per_ray(c), 1), hit_fraction=round(c["hits"] / n, 4),
                        aabb_per_ray=round(c["aabb"] / n, 2), sphere_per_ray=round(c["sphere"] / n, 2),
                        cube_per_ray=round(c["cube"] / n, 2), cylinder_per_ray=round(c["cylinder"] / n, 2),
                        compute_per_ray=round(c["compute_calls"] / n, 2), iters_per_ray=round(c["iters"] / n, 1),
                        max_iters=c["max_iters"], rays=n)
        print(key, out[key], flush=True)
    # configs[4]: the 4096-primitive synthetic tree at 7680x4320 x 16 rays/pixel is 531 M rays of the reference algorithm — hours on
    # the host.  Estimate from a sample: every 16th scanline of the 7680x4320 frame at one ray per pixel (the 16 samples of a pixel
    # see the same geometry statistics); the per-ray averages are what roofline figures for that config should use.
    sys.path.insert(0, ROOT)
    import csg_b200 as g  # noqa: E402  (scene generator only: host code, no GPU needed)
    txt = g.Scene.generate_text(4096, 1234)
    v = View(7680, 4320)
    tot = None
    for y in range(0, 4320, 16):
        c = orc.render(txt, v, rows=(y, y + 1), want_rgba=False).counters
        tot = c if tot is None else {k: (max(tot[k], c[k]) if k == "max_iters" else tot[k] + c[k]) for k in c}
    n = tot["rays"]
    out["synthetic4096@7680x4320x16spp/default (estimate: every 16th scanline at 1 ray/pixel)"] = dict(
        flop_per_ray=round(flop_per_ray(tot), 1), hit_fraction=round(tot["hits"] / n, 4), aabb_per_ray=round(tot["aabb"] / n, 2),
        sphere_per_ray=round(tot["sphere"] / n, 2), cube_per_ray=round(tot["cube"] / n, 2), cylinder_per_ray=round(tot["cylinder"] / n, 2),
        compute_per_ray=round(tot["compute_calls"] / n, 2), iters_per_ray=round(tot["iters"] / n, 1), max_iters=tot["max_iters"], rays=n)
    print("synthetic4096", out[list(out)[-1]], flush=True)
    out["_weights"] = dict(W, raygen_per_ray=40, details_phong_per_hit=130)
    with open(os.path.join(ROOT, "profiles", "flop_per_ray.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
