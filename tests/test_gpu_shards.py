"""Sharded frames: the two hand-outs (interleaved tiles gathered in one framebuffer; interleaved tile rows, every shard copying
its own rows to host memory), the device-side start gate and join, and the C++ host side — all byte-identical to the frame one
GPU renders alone.  Shard contexts are plain csg_upload_shard contexts, so most of this runs on ONE GPU (several shards of a
frame on the same device); the in-process multi-GPU forms run where the box has more GPUs."""
import os
import subprocess

import numpy as np
import pytest

import scenes
from oracle_py import View, oblique_view, orbit_view

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cam_of(csg, v):
    return csg.Camera(pos=v.pos, pitch=v.pitch, yaw=v.yaw, fov=v.fov)


def cases():
    out = [("inline:nested", 257, 129, orbit_view(257, 129, 3, radius=4.0)),      # partial tiles, ragged last tile row
           ("inline:deep_left_chain", 640, 360, View(640, 360, pos=(2.5, 0.5, 4.0), pitch=-0.1, yaw=0.5))]
    if "testCheese512" in scenes.corpus_names():
        # 1920x1080 = 30 x 34 macro tiles = 1020, not a multiple of 8; off-centre cameras move the traced rectangle around
        out += [("corpus:testCheese512", 1920, 1080, View(1920, 1080)),
                ("corpus:testCheese512", 1920, 1080, View(1920, 1080, pos=(14.0, 6.0, 9.0), pitch=-0.05, yaw=0.1)),
                ("corpus:testCheese512", 1920, 1080, oblique_view(1920, 1080))]
    return out


def single(csg, sc, w, h, cam, light, ss=1):
    one = sc.upload(w, h)
    if ss > 1:
        one.set_supersampling(ss)
    img = one.render(cam, light).copy()
    one.close()
    return img


@pytest.mark.parametrize("count", [2, 3, 8])
def test_row_sharded_host_frames_equal_the_single_gpu_frame(count, csg):
    """csg_render(host pointer) on shard contexts: every shard renders its rows of tiles and copies them itself (one 2-D copy per
    band); all shards into the same host frame == one GPU's frame."""
    light = csg.Light()
    for scene_id, w, h, v in cases():
        sc = csg.Scene.parse(scenes.text_of(scene_id))
        cam = cam_of(csg, v)
        want = single(csg, sc, w, h, cam, light)
        got = np.zeros((h, w, 4), np.uint8)
        ctxs = [sc.upload_shard(w, h, 0, r, count) for r in range(count)]
        for c in ctxs:
            c.render(cam, light, got)
        assert np.array_equal(got, want), f"{scene_id} {w}x{h} rows over {count} shards: {int((got != want).sum())} bytes differ"
        for c in ctxs:
            c.close()
        sc.close()


def test_row_sharded_host_frames_pinned_and_supersampled(csg):
    txt = csg.Scene.generate_text(300, seed=3)
    w, h, count = 200, 120, 4
    sc = csg.Scene.parse(txt)
    cam, light = csg.Camera(pos=(0.0, 0.0, 5.0)), csg.Light()
    want = single(csg, sc, w, h, cam, light, ss=2)
    got = np.zeros((h, w, 4), np.uint8)
    csg.pin_host_buffer(got.ctypes.data, got.nbytes)
    try:
        ctxs = [sc.upload_shard(w, h, 0, r, count).set_supersampling(2) for r in range(count)]
        for c in ctxs:
            c.render(cam, light, got)
        assert np.array_equal(got, want)
        for c in ctxs:
            c.close()
    finally:
        csg.unpin_host_buffer(got.ctypes.data)


@pytest.mark.parametrize("count", [2, 8])
def test_tile_sharded_frames_gathered_in_one_framebuffer(count, csg):
    """Interleaved tiles, every shard storing into shard 0's framebuffer (here: all shards on one device, one after the other,
    no gate — csg_set_gather_target with a bare pointer)."""
    light = csg.Light()
    for scene_id, w, h, v in cases():
        sc = csg.Scene.parse(scenes.text_of(scene_id))
        cam = cam_of(csg, v)
        want = single(csg, sc, w, h, cam, light)
        ctxs = [sc.upload_shard(w, h, 0, r, count) for r in range(count)]
        fb = ctxs[0].framebuffer()
        for c in ctxs:
            c.set_gather_target(fb)
        for c in ctxs:
            c.enqueue(cam, light)
            c.sync()
        got = ctxs[0].read_framebuffer()
        assert np.array_equal(got, want), f"{scene_id} {w}x{h} tiles over {count} shards: {int((got != want).sum())} bytes differ"
        for c in ctxs:
            c.close()
        sc.close()


@pytest.mark.parametrize("peer_first", [True, False])
def test_device_side_gate_and_join_between_two_shards(peer_first, csg):
    """Two shards of a frame on one device, joined through the root's sync words: the peer's kernels wait for the root's start word
    (peer enqueued first), the root's frame kernel does not end before the peer's done word (root enqueued first).
    Peer first only on small frames: a waiting frame kernel of 148 CTAs holds every SM's register file, so on ONE device the
    root's kernels could never start (on its own GPU a peer waits alone)."""
    light = csg.Light()
    for scene_id, w, h, v in (cases()[:1] if peer_first else cases()[:3]):   # 257x129: 35 frame CTAs
        sc = csg.Scene.parse(scenes.text_of(scene_id))
        cam = cam_of(csg, v)
        want = single(csg, sc, w, h, cam, light)
        root, peer = sc.upload_shard(w, h, 0, 0, 2), sc.upload_shard(w, h, 0, 1, 2)
        peer.set_gather_root(root)
        for frame in range(3):                      # sequence numbers advance; the words are reused
            for c in ((peer, root) if peer_first else (root, peer)):
                c.enqueue(cam, light)
            root.sync()
            peer.sync()
            got = root.read_framebuffer()
            assert np.array_equal(got, want), f"{scene_id} frame {frame}: {int((got != want).sum())} bytes differ"
            assert root.last_frame_ms() > 0
        peer.close()
        root.close()
        sc.close()


def test_a_missing_peer_is_a_timeout_error_not_a_hang(csg):
    sc = csg.Scene.parse(scenes.INLINE["nested"])
    root = sc.upload_shard(128, 72, 0, 0, 2)
    root.enqueue(csg.Camera(), csg.Light())        # nobody renders shard 1: the join gives up after 2 s
    with pytest.raises(csg.CsgError):
        root.sync()
    root.close()


def _multi_gpu_counts():
    import torch
    n = min(torch.cuda.device_count(), 8)
    return sorted({2, n}) if n >= 2 else []


def test_in_process_multi_gpu_contexts(csg):
    """csg_upload(n_gpus = k): device output = tiles gathered in GPU 0's framebuffer over NVLink (gate + join on the device);
    host output = every GPU copies its own rows.  Needs 2+ GPUs."""
    counts = _multi_gpu_counts()
    if not counts:
        pytest.skip("needs 2 GPUs")
    light = csg.Light()
    for scene_id, w, h, v in cases():
        sc = csg.Scene.parse(scenes.text_of(scene_id))
        cam = cam_of(csg, v)
        want = single(csg, sc, w, h, cam, light)
        for k in counts:
            many = sc.upload(w, h, k)
            for frame in range(2):
                many.enqueue(cam, light)
                got = many.read_framebuffer()
                assert np.array_equal(got, want), f"{scene_id}: {k}-GPU gathered frame differs"
                got = many.render(cam, light)
                assert np.array_equal(got, want), f"{scene_id}: {k}-GPU host frame differs"
            many.close()
        sc.close()


def test_cpp_host_side_renders_the_same_frame(csg, tmp_path):
    """host/csg_render_cli.cpp over host/csg_raycaster.hpp (the C++ mirror of CSGTree::Parse / Raycaster::ChangeSize / Raycast):
    its PPM == the frame through the ctypes binding."""
    exe = os.path.join(ROOT, "cuda-csg-tree-raycasting_b200", "csg_render")
    if not os.path.exists(exe):
        pytest.skip("csg_render not built")
    w, h = 320, 200
    for name, txt, camargs, ss in [("nested", scenes.INLINE["nested"], ["1.5", "1.0", "4.0", "-0.2", "0.3"], 1),
                                   ("chain", scenes.INLINE["deep_left_chain"], ["0.5", "0.2", "6.0", "0.0", "0.1"], 2)]:
        path = tmp_path / (name + ".txt")
        path.write_text(txt)
        out = tmp_path / (name + ".ppm")
        cmd = [exe, str(path), "--w", str(w), "--h", str(h), "--cam", *camargs, "--light", "0.7", "2.0", "--ss", str(ss), "--out", str(out)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "ms on the GPU" in r.stdout
        data = out.read_bytes()
        header = f"P6\n{w} {h}\n255\n".encode()
        assert data.startswith(header)
        ppm = np.frombuffer(data[len(header):], np.uint8).reshape(h, w, 3)
        sc = csg.Scene.parse(txt)
        ctx = sc.upload(w, h)
        if ss > 1:
            ctx.set_supersampling(ss)
        x, y, z, pitch, yaw = (float(a) for a in camargs)
        img = ctx.render(csg.Camera(pos=(x, y, z), pitch=pitch, yaw=yaw), csg.Light(0.7, 2.0))
        assert np.array_equal(ppm, img[::-1, :, :3])      # the PPM is top-down, the framebuffer bottom-up (Q2)
        ctx.close()
    # a malformed scene is reported the way Application::LoadCSGTree reports it (Application.cpp:81), exit code 1
    bad = tmp_path / "bad.txt"
    bad.write_text("Union\n Sphere 0 0 0 FF0000 1\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Cannot load tree" in r.stderr
