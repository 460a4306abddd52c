"""ctypes binding of libcsg_b200.so (the C ABI in include/csg_b200.h).

Host-side mirror of the reference's interface for the raycast path, for Python callers
(tests, bench.py): CSGTree.Parse -> Scene.parse, Raycaster.ChangeSize -> Scene.upload,
Raycaster.Raycast -> Context.render*, Camera / DirectionalLight -> Camera / Light.
There is no fallback: importing works without a GPU (symbols are checked), rendering
raises CsgError(CSG_ERR_NO_DEVICE) when no CUDA device is usable, and a missing
libcsg_b200.so raises ImportError.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSG_B200_LIB") or os.path.join(_HERE, "libcsg_b200.so")   # override = tuning sweeps only

CSG_OK, CSG_ERR_PARSE, CSG_ERR_IO, CSG_ERR_CUDA, CSG_ERR_ARG, CSG_ERR_NO_DEVICE, CSG_ERR_LIMIT = range(7)

# every symbol include/csg_b200.h declares
EXPORTS = [
    "csg_load_scene", "csg_parse_scene", "csg_free_scene", "csg_scene_counts", "csg_scene_dump", "csg_scene_flatten", "csg_scene_write",
    "csg_generate_scene", "csg_camera_default", "csg_camera_set", "csg_camera_set_fov_degrees", "csg_light_default",
    "csg_light_direction", "csg_upload", "csg_upload_shard", "csg_free_context", "csg_scene_set_optimize",
    "csg_render", "csg_render_batch", "csg_render_f32", "csg_render_aov", "csg_render_stats", "csg_set_supersampling", "csg_set_pruning", "csg_set_view_cache", "csg_prune_stats", "csg_render_enqueue", "csg_sync", "csg_last_frame_ms",
    "csg_launch_count", "csg_framebuffer", "csg_framebuffer_ipc_handle", "csg_set_gather_target_ipc",
    "csg_set_gather_target", "csg_read_framebuffer", "csg_device_tan_half_fov", "csg_fp32_peak_tflops", "csg_context_info",
    "csg_last_error", "csg_version", "csg_cube_normal_threshold", "csg_shard_tile",
    "csg_pin_host_buffer", "csg_unpin_host_buffer", "csg_set_gather_root", "csg_stream",
]


class CsgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"csg_b200 error {code}: {message}")
        self.code = code
        self.message = message


class CCamera(C.Structure):
    """csg_camera == Camera (Camera.h:9-16)."""
    _fields_ = [("pos", C.c_float * 3), ("pitch", C.c_float), ("yaw", C.c_float), ("fov", C.c_float),
                ("forward", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3)]


class CLight(C.Structure):
    _fields_ = [("polar", C.c_float), ("azimuth", C.c_float)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "csg_load_scene": (i, [C.c_char_p, C.POINTER(vp)]),
        "csg_parse_scene": (i, [C.c_char_p, C.c_size_t, C.POINTER(vp)]),
        "csg_free_scene": (None, [vp]),
        "csg_scene_counts": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "csg_scene_dump": (i, [vp, vp, vp]),
        "csg_scene_flatten": (i, [vp, vp, vp, C.POINTER(i), C.POINTER(i)]),
        "csg_scene_write": (C.c_size_t, [vp, C.c_char_p, C.c_size_t]),
        "csg_generate_scene": (C.c_size_t, [i, C.c_uint64, C.c_char_p, C.c_size_t]),
        "csg_camera_default": (None, [C.POINTER(CCamera)]),
        "csg_camera_set": (None, [C.POINTER(CCamera), f, f, f, f, f]),
        "csg_camera_set_fov_degrees": (None, [C.POINTER(CCamera), f]),
        "csg_light_default": (None, [C.POINTER(CLight)]),
        "csg_light_direction": (None, [C.POINTER(CLight), C.POINTER(f)]),
        "csg_upload": (i, [vp, i, i, i, C.POINTER(vp)]),
        "csg_upload_shard": (i, [vp, i, i, i, i, i, C.POINTER(vp)]),
        "csg_free_context": (None, [vp]),
        "csg_scene_set_optimize": (i, [vp, i]),
        "csg_render": (i, [vp, C.POINTER(CCamera), C.POINTER(CLight), vp]),
        "csg_render_batch": (i, [vp, vp, i, C.POINTER(CLight), vp]),
        "csg_render_f32": (i, [vp, C.POINTER(CCamera), C.POINTER(CLight), vp]),
        "csg_render_aov": (i, [vp, C.POINTER(CCamera), vp, vp, vp]),
        "csg_render_stats": (i, [vp, C.POINTER(CCamera), vp]),
        "csg_set_supersampling": (i, [vp, i]),
        "csg_set_pruning": (i, [vp, i]),
        "csg_set_view_cache": (i, [vp, i]),
        "csg_prune_stats": (i, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(C.c_longlong)]),
        "csg_render_enqueue": (i, [vp, C.POINTER(CCamera), C.POINTER(CLight), vp]),
        "csg_sync": (i, [vp]),
        "csg_last_frame_ms": (i, [vp, C.POINTER(f)]),
        "csg_launch_count": (C.c_uint64, [vp]),
        "csg_stream": (i, [vp, C.POINTER(vp)]),
        "csg_framebuffer": (i, [vp, C.POINTER(vp)]),
        "csg_framebuffer_ipc_handle": (i, [vp, vp]),
        "csg_set_gather_target_ipc": (i, [vp, vp]),
        "csg_set_gather_target": (i, [vp, vp]),
        "csg_read_framebuffer": (i, [vp, vp]),
        "csg_device_tan_half_fov": (i, [vp, f, C.POINTER(f)]),
        "csg_fp32_peak_tflops": (i, [i, C.POINTER(f)]),
        "csg_context_info": (C.c_char_p, [vp]),
        "csg_last_error": (C.c_char_p, []),
        "csg_version": (C.c_char_p, []),
        "csg_cube_normal_threshold": (f, [f, f]),
        "csg_shard_tile": (i, [i] * 10 + [C.POINTER(i)] * 5),
        "csg_pin_host_buffer": (i, [vp, C.c_size_t]),
        "csg_unpin_host_buffer": (i, [vp]),
        "csg_set_gather_root": (i, [vp, vp]),
    }
    for name in EXPORTS:
        fn = getattr(lib, name)  # AttributeError here = the library does not export what the header declares
        fn.restype, fn.argtypes = sig[name]
    return lib


lib = _load()


def _check(rc):
    if rc != CSG_OK:
        raise CsgError(rc, lib.csg_last_error().decode(errors="replace"))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def version():
    return lib.csg_version().decode()


def cube_normal_threshold(half_size, level):
    return float(lib.csg_cube_normal_threshold(float(half_size), float(level)))


def shard_tile(macro_x, macro_y, rect, mode, rank, count, tile):
    """(mx, my, slot, n_tiles, n_slots) of tile number `tile` of shard `rank`; tile < 0: only the two counts."""
    mx, my, slot, nt, ns = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib.csg_shard_tile(macro_x, macro_y, rect[0], rect[1], rect[2], rect[3], mode, rank, count, tile,
                            C.byref(mx), C.byref(my), C.byref(slot), C.byref(nt), C.byref(ns))
    if tile < 0:
        return None, None, None, nt.value, ns.value
    _check(rc)
    return mx.value, my.value, slot.value, nt.value, ns.value


def pin_host_buffer(ptr, nbytes):
    _check(lib.csg_pin_host_buffer(_ptr(ptr), nbytes))


def unpin_host_buffer(ptr):
    _check(lib.csg_unpin_host_buffer(_ptr(ptr)))


def fp32_peak_tflops(device=0):
    out = C.c_float()
    _check(lib.csg_fp32_peak_tflops(device, C.byref(out)))
    return out.value


class Camera:
    """Mirror of the reference's Camera (RenderManager/Camera/Camera.h)."""

    def __init__(self, pos=None, pitch=0.0, yaw=0.0, fov=None, fov_degrees=None):
        self.c = CCamera()
        lib.csg_camera_default(C.byref(self.c))
        if pos is not None or pitch or yaw:
            p = pos if pos is not None else (0.0, 0.0, 5.0)
            self.set(p[0], p[1], p[2], pitch, yaw)
        if fov_degrees is not None:
            lib.csg_camera_set_fov_degrees(C.byref(self.c), fov_degrees)
        if fov is not None and fov > 0:
            self.c.fov = fov

    def set(self, x, y, z, pitch, yaw):
        lib.csg_camera_set(C.byref(self.c), x, y, z, pitch, yaw)
        return self

    def as_array(self):
        return np.frombuffer(bytes(self.c), dtype=np.float32).copy()


class Light:
    """Mirror of DirectionalLight (RenderManager/DirectionalLight.h)."""

    def __init__(self, polar=None, azimuth=None):
        self.c = CLight()
        lib.csg_light_default(C.byref(self.c))
        if polar is not None:
            self.c.polar = polar
            self.c.azimuth = azimuth if azimuth is not None else 0.0

    def direction(self):
        out = (C.c_float * 3)()
        lib.csg_light_direction(C.byref(self.c), out)
        return np.array(out, dtype=np.float32)


class Scene:
    """Parsed CSG tree (CSGTree::Parse)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def parse(cls, text, optimize=None):
        if isinstance(text, str):
            text = text.encode()
        h = C.c_void_p()
        _check(lib.csg_parse_scene(text, len(text), C.byref(h)))
        s = cls(h)
        if optimize is not None:
            s.set_optimize(optimize)
        return s

    @classmethod
    def load(cls, path, optimize=None):
        h = C.c_void_p()
        _check(lib.csg_load_scene(os.fsencode(path), C.byref(h)))
        s = cls(h)
        if optimize is not None:
            s.set_optimize(optimize)
        return s

    @staticmethod
    def generate_text(n_primitives, seed=1234):
        n = lib.csg_generate_scene(n_primitives, seed, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib.csg_generate_scene(n_primitives, seed, buf, n + 1)
        return buf.value

    def set_optimize(self, level):
        _check(lib.csg_scene_set_optimize(self.h, int(level)))
        return self

    def counts(self):
        a, b, d = C.c_int(), C.c_int(), C.c_int()
        _check(lib.csg_scene_counts(self.h, C.byref(a), C.byref(b), C.byref(d)))
        return a.value, b.value, d.value

    def dump(self):
        nn, npr, _ = self.counts()
        nodes = np.zeros((nn, 44), np.uint8)
        prims = np.zeros((npr, 48), np.uint8)
        _check(lib.csg_scene_dump(self.h, _ptr(nodes), _ptr(prims)))
        return nodes, prims

    def flatten(self):
        """The GPU layout: (records as uint32[n, 8] — view as float32 for the first 6 words — , parents int32[n], depth)."""
        n, d = C.c_int(), C.c_int()
        _check(lib.csg_scene_flatten(self.h, None, None, C.byref(n), C.byref(d)))
        rec = np.zeros((n.value, 8), np.uint32)
        par = np.zeros(n.value, np.int32)
        _check(lib.csg_scene_flatten(self.h, _ptr(rec), _ptr(par), C.byref(n), C.byref(d)))
        return rec, par, d.value

    def write(self):
        n = lib.csg_scene_write(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib.csg_scene_write(self.h, buf, n + 1)
        return buf.value

    def upload(self, width, height, n_gpus=1):
        h = C.c_void_p()
        _check(lib.csg_upload(self.h, width, height, n_gpus, C.byref(h)))
        return Context(h, width, height)

    def upload_shard(self, width, height, device, rank, count):
        h = C.c_void_p()
        _check(lib.csg_upload_shard(self.h, width, height, device, rank, count, C.byref(h)))
        return Context(h, width, height)

    def close(self):
        if self.h:
            lib.csg_free_scene(self.h)
            self.h = None

    __del__ = close


class Context:
    """Uploaded scene + framebuffers on the device(s) (Raycaster after ChangeSize)."""

    def __init__(self, handle, width, height):
        self.h, self.width, self.height = handle, width, height

    def info(self):
        return json.loads(lib.csg_context_info(self.h).decode())

    def render(self, cam, light, out=None):
        """RGBA8 frame.  out: numpy uint8 array (host) or an int device address; returns the array."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        _check(lib.csg_render(self.h, C.byref(cam.c), C.byref(light.c), _ptr(out)))
        return out

    def render_f32(self, cam, light, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        _check(lib.csg_render_f32(self.h, C.byref(cam.c), C.byref(light.c), _ptr(out)))
        return out

    def render_aov(self, cam):
        n = self.width * self.height
        hit = np.empty(n, np.uint8)
        prim = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        _check(lib.csg_render_aov(self.h, C.byref(cam.c), _ptr(hit), _ptr(prim), _ptr(t)))
        return hit, prim, t

    def set_supersampling(self, samples_per_axis):
        _check(lib.csg_set_supersampling(self.h, int(samples_per_axis)))
        return self

    def render_batch(self, cams, light, out=None):
        """n frames, one per camera, pipelined over two frame slots; out = host array / pinned or device pointer, or None."""
        n = len(cams)
        arr = (CCamera * n)(*[c.c for c in cams])
        ret = None
        if out is None:
            ret = np.empty((n, self.height, self.width, 4), np.uint8)
            out = ret
        _check(lib.csg_render_batch(self.h, arr, n, C.byref(light.c), _ptr(out)))
        return ret

    def set_pruning(self, enabled):
        """False/0: every tile reads the whole tree; True/1: per-tile trees (default kernel); 2: per-tile trees by the tree-walking kernel."""
        _check(lib.csg_set_pruning(self.h, int(enabled)))
        return self

    def set_view_cache(self, enabled):
        _check(lib.csg_set_view_cache(self.h, int(bool(enabled))))
        return self

    def prune_stats(self):
        a, b, c_, n = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        _check(lib.csg_prune_stats(self.h, C.byref(a), C.byref(b), C.byref(c_), C.byref(n)))
        return {"traced_tiles": a.value, "empty_tiles": b.value, "fallback_tiles": c_.value, "pruned_nodes": n.value}

    def render_stats(self, cam):
        it = np.empty(self.width * self.height, np.int32)
        _check(lib.csg_render_stats(self.h, C.byref(cam.c), _ptr(it)))
        return it

    def enqueue(self, cam, light, out_dev=None):
        _check(lib.csg_render_enqueue(self.h, C.byref(cam.c), C.byref(light.c), _ptr(out_dev)))

    def sync(self):
        _check(lib.csg_sync(self.h))

    def last_frame_ms(self):
        ms = C.c_float()
        _check(lib.csg_last_frame_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib.csg_launch_count(self.h))

    def stream(self):
        """cudaStream_t (as an integer) the root GPU's share of a frame is enqueued on (csg_stream)."""
        p = C.c_void_p()
        _check(lib.csg_stream(self.h, C.byref(p)))
        return p.value or 0

    def framebuffer(self):
        p = C.c_void_p()
        _check(lib.csg_framebuffer(self.h, C.byref(p)))
        return p.value

    def read_framebuffer(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        _check(lib.csg_read_framebuffer(self.h, _ptr(out)))
        return out

    def ipc_handle(self):
        buf = C.create_string_buffer(64)
        _check(lib.csg_framebuffer_ipc_handle(self.h, buf))
        return buf.raw

    def set_gather_target_ipc(self, handle_bytes):
        buf = C.create_string_buffer(handle_bytes, 64)
        _check(lib.csg_set_gather_target_ipc(self.h, buf))

    def set_gather_root(self, root_ctx):
        _check(lib.csg_set_gather_root(self.h, root_ctx.h))

    def set_gather_target(self, dev_ptr):
        _check(lib.csg_set_gather_target(self.h, _ptr(dev_ptr)))

    def device_tan_half_fov(self, fov):
        out = C.c_float()
        _check(lib.csg_device_tan_half_fov(self.h, fov, C.byref(out)))
        return out.value

    def close(self):
        if self.h:
            lib.csg_free_context(self.h)
            self.h = None

    __del__ = close
