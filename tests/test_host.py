"""Host-side logic of libcsg_b200 (no GPU needed): C-ABI exports, parser parity with the oracle, error behaviour,
camera/light math, writer/generator, multi-rank tile sharding logic (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import scenes
from oracle_py import View, ParseError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(csg):
    header = open(os.path.join(ROOT, "include", "csg_b200.h")).read()
    declared = set(re.findall(r"\b(csg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = C.CDLL(csg.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/csg_b200.h but not exported"
    assert declared == set(csg.EXPORTS)


def test_library_does_not_link_the_oracle(csg):
    out = subprocess.run(["ldd", csg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ref_cpu" not in out and "ref_gpu" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", csg.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms and "refcpu_" not in syms


@pytest.mark.parametrize("scene_id", scenes.all_scene_ids())
def test_parser_matches_oracle(scene_id, csg, oracle):
    """csg_parse_scene == CSGTree::Parse: reference-layout node and primitive arrays byte-identical to the oracle's
    (which tests/test_oracle_vs_reference.py pins to the reference parser itself)."""
    txt = scenes.text_of(scene_id)
    s = csg.Scene.parse(txt)
    n1, p1 = s.dump()
    sc = oracle.parse(txt)
    n2, p2 = oracle.tree_arrays(sc)
    oracle.free(sc)
    assert np.array_equal(n1, n2)
    types = n1.view(np.int32)[:, 0]
    prim = n1.view(np.int32)[:, 1]
    for t, pi in zip(types, prim):
        if pi >= 0:
            nb = 48 if t == 4 else 32
            assert np.array_equal(p1[pi, :nb], p2[pi, :nb])
    nn, npr, depth = s.counts()
    assert nn == 2 * npr - 1 and depth >= (0 if npr == 1 else 1)


BAD_INPUTS = [
    ("", "Cannot parse - number of primitives do not match number of nodes"),
    ("Union Sphere 0 0 0 FF00FF 1", "Cannot parse - number of primitives do not match number of nodes"),
    ("Sphere 0 0 0 FF00FF 1 Sphere 0 0 0 FF00FF 1", "Cannot parse"),
    ("Blob 0 0 0", "Cannot parse - Unrecognized keyword: Blob"),
    ("Sphere 0 0 0 FF00F 1", "Cannot parse color FF00F"),
    ("Sphere x 0 0 FF00FF 1", "stof"),
    ("Sphere 0 0 0 GG00FF 1", "stoi"),
    ("Cylinder 0 0 0 FF00FF 1 2 361 0 0", "Invalid roation rotX should be in range [0, 360] deg"),
    ("Cylinder 0 0 0 FF00FF 1 2 0 -1 0", "Invalid roation rotY should be in range [0, 360] deg"),
    ("Cylinder 0 0 0 FF00FF 1 2 0 0 400", "Invalid roation rotZ should be in range [0, 360] deg"),
    ("Cylinder 0 0 0 FF00FF 1 2 0 0 q", "stod"),
    ("Union Sphere 0 0 0 FF00FF", "Cannot parse - unexpected end of input"),  # reference: out-of-range read (UB)
]


@pytest.mark.parametrize("text,message", BAD_INPUTS)
def test_parse_errors(text, message, csg, oracle):
    with pytest.raises(csg.CsgError) as e:
        csg.Scene.parse(text)
    assert e.value.code == csg.CSG_ERR_PARSE and e.value.message == message
    with pytest.raises(ParseError) as e2:
        oracle.parse(text)
    assert str(e2.value) == message


def test_load_scene_io_error(csg):
    with pytest.raises(csg.CsgError) as e:
        csg.Scene.load("/nonexistent/scene.txt")
    assert e.value.code == csg.CSG_ERR_IO


def test_camera_and_light_match_oracle(csg, oracle):
    rng = np.random.default_rng(11)
    for i in range(500):
        v = View(64, 36, pos=rng.uniform(-10, 10, 3), pitch=rng.uniform(-2, 2), yaw=rng.uniform(-7, 7),
                 fov=rng.uniform(0.2, 2.5) if i % 2 else -1, polar=rng.uniform(-3, 3), azimuth=rng.uniform(-3, 3))
        cam = csg.Camera(pos=v.pos, pitch=v.pitch, yaw=v.yaw, fov=v.fov)
        assert np.array_equal(cam.as_array().view(np.uint32), oracle.camera(v).as_array().view(np.uint32))
        li = csg.Light(v.polar, v.azimuth)
        assert np.array_equal(li.direction().view(np.uint32), np.array(oracle.light_dir(v), np.float32).view(np.uint32))
    d = View(8, 8)
    assert np.array_equal(csg.Camera().as_array().view(np.uint32), oracle.camera(d).as_array().view(np.uint32))
    assert np.array_equal(csg.Light().direction().view(np.uint32), np.array(oracle.light_dir(d), np.float32).view(np.uint32))


def test_generator_and_writer_roundtrip(csg):
    txt = csg.Scene.generate_text(4096, seed=1234)
    s = csg.Scene.parse(txt)
    nn, npr, depth = s.counts()
    assert (nn, npr, depth) == (8191, 4096, 12)
    assert csg.Scene.generate_text(4096, seed=1234) == txt          # deterministic
    assert csg.Scene.generate_text(4096, seed=1235) != txt
    small = csg.Scene.parse(csg.Scene.generate_text(37, seed=5))
    out = small.write()
    again = csg.Scene.parse(out)
    n1, p1 = small.dump()
    n2, p2 = again.dump()
    assert np.array_equal(n1.view(np.int32)[:, :5], n2.view(np.int32)[:, :5])   # same tree shape
    a = p1.view(np.float32).reshape(len(p1), 12)
    b = p2.view(np.float32).reshape(len(p2), 12)
    types = {pi: t for t, pi in zip(n1.view(np.int32)[:, 0], n1.view(np.int32)[:, 1]) if pi >= 0}
    for i in range(len(a)):
        assert np.allclose(a[i, 1:4], b[i, 1:4], rtol=0, atol=0)   # positions exact (%.9g)
        assert a[i, 7] == b[i, 7]
        if types[i] == 4:
            assert np.allclose(a[i, 9:12], b[i, 9:12], atol=2e-6)    # axis re-derived from Euler angles


def test_no_gpu_means_error_not_fallback(csg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    s = csg.Scene.parse(scenes.INLINE["nested"])
    with pytest.raises(csg.CsgError) as e:
        s.upload(64, 36)
    assert e.value.code == csg.CSG_ERR_NO_DEVICE


def test_tile_sharding_covers_frame_once():
    """The kernel renders macro tile m on shard m % S; local ticket i maps to macro (i>>6)*S + rank.  Check the host-side
    arithmetic the launcher uses (number of local warp tiles per shard) covers every macro tile exactly once."""
    for (w, h) in [(3840, 2160), (1920, 1080), (7680, 4320), (100, 50), (64, 32), (65, 33)]:
        mx, my = (w + 63) // 64, (h + 31) // 32
        total = mx * my
        for S in (1, 2, 3, 4, 8):
            seen = np.zeros(total, int)
            for r in range(S):
                mine = (total - r + S - 1) // S
                for j in range(mine):
                    seen[j * S + r] += 1
            assert (seen == 1).all()


GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
from oracle_py import Oracle, View
import csg_b200 as g
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# Host-side model of the multi-rank frame, with the oracle standing in for the kernels and the library's own tile hand-out
# (csg_shard_tile = what the kernels compile):
#   tiles: each rank renders the 64x32 macro tiles it is dealt and "stores" them into rank 0's framebuffer (gloo gather);
#   rows : each rank renders the tile rows it is dealt and writes them into ONE shared host frame (bench.SharedHost, the
#          /dev/shm buffer + spin barrier bench.py's N > 1 end-to-end leg uses; not page-locked here).
# Either way the result must equal the single-rank frame byte for byte.
W, H = 200, 100
txt = "Difference\n Cube 0 0 0 FF0000 2\n Union\n  Sphere 1 1 1 00FF00 0.8\n  Cylinder 0 0 0 0000FF 0.5 3 0 0 0\n"
orc = Oracle()
full = orc.render(txt, View(W, H, pos=(1.5, 1.0, 4.0), pitch=-0.2, yaw=0.3))
rgba = full.rgba8().reshape(H, W, 4)
mx, my = (W + 63) // 64, (H + 31) // 32
rect = (0, 0, mx, my)

def my_tiles(mode):
    n = g.shard_tile(mx, my, rect, mode, rank, world, -1)[3]
    return [g.shard_tile(mx, my, rect, mode, rank, world, t)[:2] for t in range(n)]

mine = np.zeros((H, W, 4), np.uint8)
mask = np.zeros((H, W), np.uint8)
for (tx, ty) in my_tiles(0):
    x0, y0 = tx * 64, ty * 32
    mine[y0:y0 + 32, x0:x0 + 64] = rgba[y0:y0 + 32, x0:x0 + 64]
    mask[y0:y0 + 32, x0:x0 + 64] += 1
parts = [torch.zeros(H, W, 4, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
masks = [torch.zeros(H, W, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
dist.gather(torch.from_numpy(mine), parts, dst=0)
dist.gather(torch.from_numpy(mask), masks, dst=0)
if rank == 0:
    cover = sum(m.numpy().astype(int) for m in masks)
    assert (cover == 1).all(), "tiles must partition the frame"
    fb = sum(p.numpy().astype(int) for p in parts).astype(np.uint8)
    assert np.array_equal(fb, rgba)

host = bench.SharedHost(g, W * H * 4, rank, world, dist, "gloo_test", pin=False)
frame = host.arr.reshape(H, W, 4)
for step in range(3):
    host.spin_barrier()
    for (tx, ty) in my_tiles(1):
        assert ty % world == rank
        x0, y0 = tx * 64, ty * 32
        frame[y0:y0 + 32, x0:x0 + 64] = rgba[y0:y0 + 32, x0:x0 + 64]
    host.spin_barrier()
    if rank == 0:
        assert np.array_equal(frame, rgba), "rows must partition the frame"
    host.spin_barrier()
    if rank == 0:
        frame[:] = 0
frame = None
host.close()
if rank == 0:
    print("GLOO_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_gloo(tmp_path):
    """world_size-2 CPU run of the N>1 host logic (gloo backend)."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout


def test_cube_normal_thresholds_reproduce_the_reference_expression(csg):
    """The kernel replaces (float)(int)((pc / half) * 1.00001f) (RaycastingKernels.cu:422-424) by two comparisons against
    host-computed thresholds; checked here against the expression itself in float32, around the steps and at random."""
    import numpy as np
    f32 = np.float32
    rng = np.random.default_rng(5)

    def ref(pc, half):
        with np.errstate(all="ignore"):
            v = (pc / half) * f32(1.00001)            # float32 operations, one rounding each
            return np.trunc(v).astype(np.float32) + f32(0.0)   # (float)(int): -0 becomes +0

    halves = [f32(x) for x in (10.0, 1.0, 0.5, 0.35, 3.0, 1e-3, 123.456, 2.0 ** -20, 7e8)] + [f32(h) for h in rng.uniform(0.01, 50, 40)]
    for half in halves:
        a1, a2 = f32(csg.cube_normal_threshold(half, 1)), f32(csg.cube_normal_threshold(half, 2))
        assert 0 < a1 < a2
        pts = [rng.uniform(-2.5, 2.5, 4000).astype(np.float32) * half]
        for a in (a1, a2):
            bits = np.array([a], np.float32).view(np.uint32)[0]
            near = (np.arange(-300, 301, dtype=np.int64) + int(bits)).astype(np.uint32).view(np.float32)
            pts += [near, -near]
        pts.append(np.array([0.0, -0.0, half, -half], np.float32))
        pc = np.concatenate(pts)
        want = ref(pc, half)
        m = np.abs(pc)
        got = np.where(m < a1, f32(0.0), np.where(m < a2, np.copysign(f32(1.0), pc), want))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"half={half}"
        # the thresholds are tight: one ulp below a1 still gives 0
        below = np.nextafter(a1, f32(0.0))
        assert ref(np.array([below], np.float32), half)[0] == 0 and ref(np.array([a1], np.float32), half)[0] == 1
    for bad in (0.0, -1.0, float("inf"), float("nan"), 1e-45):
        assert csg.cube_normal_threshold(bad, 1) == 0.0


def test_optional_viewer_compiles_against_sdl2_headers():
    """host/csg_viewer.cpp (SDL2 + OpenGL + CUDA-GL interop, off by default) is syntax-checked against the SDL2 headers the
    reference vendors; it cannot be linked or run in this image (no SDL2, no OpenGL)."""
    sdl = "/root/reference/CSGRayCasting/Libraries/include"
    if not os.path.exists(os.path.join(sdl, "SDL.h")):
        pytest.skip("SDL2 headers not available (reference checkout absent)")
    src = os.path.join(ROOT, "cuda-csg-tree-raycasting_b200", "host", "csg_viewer.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + sdl, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU implementation, or the oracle port when oracle/_ref is absent) runs
    without a GPU and prints ONE JSON line with the keys the driver reads."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "primary rays/s" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "testCheese512" in d["config"]["workload"] and "3840x2160" in d["config"]["workload"]


def test_bench_ours_refuses_to_run_without_a_gpu():
    """No CPU fallback: our arm of bench.py fails loudly when there is no CUDA device."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pytest.skip("torch missing")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-baselines"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)


def test_sass_has_no_local_memory_in_any_kernel(csg):
    """The traversal stack lives in shared memory and nothing spills: no LDL/STL in the SASS of any of the 48 frame kernel
    instantiations (three CTA shapes x three output modes x {one ray per pixel with tickets of one / of two warp tiles,
    supersampling} x {scenes with, without cylinders}; the AOV mode is per primary ray and has no supersampling variant), in
    eval_flat_union and the other device functions they call, nor in the pruning kernels; everything is built for sm_100a."""
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", csg.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in out
    assert not re.search(r"arch = sm_(?!100a)", out), "only sm_100a code is shipped"
    fn, per_fn = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = [0, 0]
            continue
        if fn and re.search(r"\b(LDL|STL)\b", line):
            per_fn[fn][0] += 1
        if fn and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            per_fn[fn][1] += 1
    frame = {f: v for f, v in per_fn.items() if "csg_frame_kernel" in f}
    assert len(frame) == 48                                            # 3 CTA shapes x (3 modes x 2 ticket sizes at 1 ray + 2 modes with supersampling) x (with, without cylinder code)
    assert len([f for f in frame if re.search(r"ELb0ELb[01]ELb[01]E", f)]) == 36 and len([f for f in frame if re.search(r"ELb1ELb[01]ELb0E", f)]) == 12
    for f, (local_ops, n) in frame.items():
        assert local_ops == 0, f"{f}: {local_ops} local-memory instructions"
        assert n > 1000
    prune = {f: v for f, v in per_fn.items() if "csg_prune" in f}
    assert len(prune) == 4                                             # the walk + the prefix-sum kernel at 128/256/512 threads
    for f, (local_ops, n) in prune.items():
        assert local_ops == 0, f"{f}: {local_ops} local-memory instructions"
