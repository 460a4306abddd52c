"""TEST INFRASTRUCTURE — ctypes access to the checkers (never imported by the product).

  Oracle   : oracle/libcsg_oracle.so      our plain-C restatement (csg_oracle.c)
  RefCPU   : oracle/_ref/libref_cpu.so    the reference's own sources, host build
  RefGPU   : oracle/_ref/libref_gpu.so    the reference's own CUDA kernels, sm_100 build

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
"""
import ctypes as C
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES_DIR = os.path.join(HERE, "_ref", "scenes")

DEFAULT_LIGHT = 1e10


def scene_text(name):
    """Scene text by corpus name ('testCheese512') or path."""
    path = name if os.path.exists(name) else os.path.join(SCENES_DIR, name + ".txt")
    with open(path, "rb") as f:
        return f.read()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OrcNode(C.Structure):
    _fields_ = [("type", C.c_int32), ("prim", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("parent", C.c_int32), ("bmin", C.c_float * 3), ("bmax", C.c_float * 3)]


class OrcPrim(C.Structure):
    _fields_ = [("id", C.c_int32), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("r", C.c_float), ("g", C.c_float), ("b", C.c_float), ("p", C.c_float * 5)]


class OrcScene(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_prims", C.c_int32), ("nodes", C.POINTER(OrcNode)),
                ("prims", C.POINTER(OrcPrim))]


class OrcCamera(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("rotX", C.c_float), ("rotY", C.c_float),
                ("fov", C.c_float), ("forward", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3)]

    def as_array(self):
        return np.frombuffer(bytes(self), dtype=np.float32).copy()


class OrcCounters(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("hits", C.c_uint64), ("iters", C.c_uint64), ("goto_calls", C.c_uint64),
                ("compute_calls", C.c_uint64), ("aabb", C.c_uint64), ("sphere", C.c_uint64),
                ("cylinder", C.c_uint64), ("cube", C.c_uint64), ("max_iters", C.c_uint64),
                ("max_action_depth", C.c_int32), ("max_hit_depth", C.c_int32), ("max_time_depth", C.c_int32),
                ("stack_overflows", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RefView(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("pos", C.c_float * 3), ("pitch", C.c_float),
                ("yaw", C.c_float), ("fov", C.c_float), ("polar", C.c_float), ("azimuth", C.c_float)]


class View:
    """Camera + light parameters of one frame (fields of Camera.h / DirectionalLight.h)."""

    def __init__(self, width, height, pos=(0.0, 0.0, 5.0), pitch=0.0, yaw=0.0, fov=-1.0,
                 polar=DEFAULT_LIGHT, azimuth=0.0):
        self.width, self.height = int(width), int(height)
        self.pos = tuple(float(x) for x in pos)
        self.pitch, self.yaw, self.fov = float(pitch), float(yaw), float(fov)
        self.polar, self.azimuth = float(polar), float(azimuth)

    def ref(self):
        return RefView(self.width, self.height, (C.c_float * 3)(*self.pos), self.pitch, self.yaw, self.fov,
                       self.polar, self.azimuth)


class Frame:
    """AOVs of one rendered frame (row 0 = bottom scanline)."""

    def __init__(self, w, h, want_flags=False):
        n = w * h
        self.w, self.h = w, h
        self.hit = np.zeros(n, np.uint8)
        self.prim = np.full(n, -1, np.int32)
        self.t = np.full(n, -1, np.float32)
        self.rgba = np.zeros(n * 4, np.float32)
        self.flags = np.zeros(n, np.uint8) if want_flags else None

    def rgba8(self):
        """Q12: u8 = (int)(clamp(c,0,1)*255 + 0.5)."""
        return (np.clip(self.rgba, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)


class ParseError(Exception):
    pass


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "libcsg_oracle.so")
        self.lib = L = C.CDLL(path)
        L.orc_parse.argtypes = [C.c_char_p, C.POINTER(C.POINTER(OrcScene)), C.c_char_p, C.c_int]
        L.orc_free.argtypes = [C.POINTER(OrcScene)]
        L.orc_camera_init.argtypes = [C.POINTER(OrcCamera)] + [C.c_float] * 6
        L.orc_light_dir.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.orc_render.argtypes = [C.POINTER(OrcScene), C.c_int, C.c_int, C.POINTER(OrcCamera), C.POINTER(C.c_float),
                                 C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(OrcCounters)]
        L.orc_hit_primitive.argtypes = [C.POINTER(OrcPrim), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.orc_aabb_hit.argtypes = [C.POINTER(C.c_float)] * 4 + [C.c_float]
        L.orc_raygen.argtypes = [C.POINTER(OrcCamera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                 C.POINTER(C.c_float)]

    def parse(self, text):
        if isinstance(text, str):
            text = text.encode()
        out = C.POINTER(OrcScene)()
        err = C.create_string_buffer(512)
        rc = self.lib.orc_parse(text, C.byref(out), err, 512)
        if rc:
            raise ParseError(err.value.decode())
        return out

    def free(self, scene):
        self.lib.orc_free(scene)

    def tree_arrays(self, scene):
        s = scene.contents
        nodes = np.frombuffer(C.string_at(s.nodes, s.n_nodes * 44), dtype=np.uint8).reshape(s.n_nodes, 44).copy()
        prims = np.frombuffer(C.string_at(s.prims, s.n_prims * 48), dtype=np.uint8).reshape(s.n_prims, 48).copy()
        return nodes, prims

    def camera(self, view):
        cam = OrcCamera()
        self.lib.orc_camera_init(C.byref(cam), view.pos[0], view.pos[1], view.pos[2], view.pitch, view.yaw, view.fov)
        return cam

    def light_dir(self, view):
        out = (C.c_float * 3)()
        self.lib.orc_light_dir(view.polar, view.azimuth, out)
        return out

    def render(self, text, view, rows=None, nthreads=0, tan_half_fov=float("nan"), want_flags=False,
               want_rgba=True):
        scene = self.parse(text)
        try:
            w, h = view.width, view.height
            fr = Frame(w, h, want_flags)
            cam = self.camera(view)
            ld = self.light_dir(view)
            y0, y1 = rows if rows else (0, h)
            cnt = OrcCounters()
            rc = self.lib.orc_render(scene, w, h, C.byref(cam), ld, tan_half_fov, y0, y1, nthreads, _p(fr.hit),
                                     _p(fr.prim), _p(fr.t), _p(fr.flags), _p(fr.rgba) if want_rgba else None,
                                     C.byref(cnt))
            if rc:
                raise RuntimeError("orc_render failed")
            fr.counters = cnt.as_dict()
            return fr
        finally:
            self.free(scene)


class _Ref:
    prefix = None

    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        f = lambda n: getattr(L, self.prefix + n)
        f("tree_info").argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        f("tree_dump").argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
        f("camera").argtypes = [C.POINTER(RefView), C.POINTER(C.c_float)]
        f("light_dir").argtypes = [C.POINTER(RefView), C.POINTER(C.c_float)]
        self._f = f

    def tree_info(self, text):
        if isinstance(text, str):
            text = text.encode()
        nn, npr = C.c_int(), C.c_int()
        err = C.create_string_buffer(512)
        rc = self._f("tree_info")(text, C.byref(nn), C.byref(npr), err, 512)
        if rc:
            raise ParseError(err.value.decode())
        return nn.value, npr.value

    def tree_arrays(self, text):
        if isinstance(text, str):
            text = text.encode()
        nn, npr = self.tree_info(text)
        nodes = np.zeros((nn, 44), np.uint8)
        prims = np.zeros((npr, 48), np.uint8)
        self._f("tree_dump")(text, _p(nodes), _p(prims))
        return nodes, prims

    def camera(self, view):
        out = (C.c_float * 15)()
        rv = view.ref()
        self._f("camera")(C.byref(rv), out)
        return np.array(out, dtype=np.float32)

    def light_dir(self, view):
        out = (C.c_float * 3)()
        rv = view.ref()
        self._f("light_dir")(C.byref(rv), out)
        return np.array(out, dtype=np.float32)


class RefCPU(_Ref):
    prefix = "refcpu_"

    def __init__(self, path=None):
        super().__init__(path or os.path.join(HERE, "_ref", "libref_cpu.so"))
        self.lib.refcpu_render.argtypes = [C.c_char_p, C.POINTER(RefView), C.c_int, C.c_int, C.c_int, C.c_int] + \
            [C.c_void_p] * 4 + [C.POINTER(C.c_double), C.c_char_p, C.c_int]
        self.lib.refcpu_max_threads.restype = C.c_int

    def max_threads(self):
        return self.lib.refcpu_max_threads()

    def render(self, text, view, rows=None, nthreads=0, outputs=True, row_step=1):
        if isinstance(text, str):
            text = text.encode()
        w, h = view.width, view.height
        fr = Frame(w, h) if outputs else None
        y0, y1 = rows if rows else (0, h)
        sec = C.c_double()
        err = C.create_string_buffer(512)
        rv = view.ref()
        rc = self.lib.refcpu_render(text, C.byref(rv), y0, y1, row_step, nthreads,
                                    _p(fr.hit) if fr else None, _p(fr.prim) if fr else None,
                                    _p(fr.t) if fr else None, _p(fr.rgba) if fr else None,
                                    C.byref(sec), err, 512)
        if rc:
            raise ParseError(err.value.decode())
        if fr is None:
            return sec.value
        fr.seconds = sec.value
        return fr


class RefGPU(_Ref):
    prefix = "refgpu_"

    def __init__(self, path=None):
        super().__init__(path or os.path.join(HERE, "_ref", "libref_gpu.so"))
        self.lib.refgpu_render.argtypes = [C.c_char_p, C.POINTER(RefView)] + [C.c_void_p] * 4 + \
            [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]

    def render_window(self, text, view, x0, y0, ww, wh):
        """The reference kernels at the full size of `view`; only the window [x0, x0+ww) x [y0, y0+wh) comes back (a Frame of
        ww x wh pixels, .ms_kernels = the time of the two kernels over the whole grid)."""
        if isinstance(text, str):
            text = text.encode()
        self.lib.refgpu_render_window.argtypes = [C.c_char_p, C.POINTER(RefView)] + [C.c_int] * 4 + [C.c_void_p] * 4 + \
            [C.POINTER(C.c_float), C.c_char_p, C.c_int]
        fr = Frame(ww, wh)
        ms = C.c_float()
        err = C.create_string_buffer(512)
        rv = view.ref()
        rc = self.lib.refgpu_render_window(text, C.byref(rv), x0, y0, ww, wh, _p(fr.hit), _p(fr.prim), _p(fr.t), _p(fr.rgba),
                                           C.byref(ms), err, 512)
        if rc == 1:
            raise ParseError(err.value.decode())
        if rc:
            raise RuntimeError("refgpu_render_window: " + err.value.decode())
        fr.ms_kernels = ms.value
        return fr

    def render(self, text, view, warmup=0, iters=1, shipped=False, outputs=True):
        if isinstance(text, str):
            text = text.encode()
        w, h = view.width, view.height
        fr = Frame(w, h) if outputs else Frame(1, 1)
        msk = np.zeros(max(iters, 1), np.float32)
        mss = np.zeros(max(iters, 1), np.float32) if shipped else None
        err = C.create_string_buffer(512)
        rv = view.ref()
        rc = self.lib.refgpu_render(text, C.byref(rv),
                                    _p(fr.hit) if outputs else None, _p(fr.prim) if outputs else None,
                                    _p(fr.t) if outputs else None, _p(fr.rgba) if outputs else None,
                                    warmup, iters, _p(msk), _p(mss), err, 512)
        if rc == 1:
            raise ParseError(err.value.decode())
        if rc:
            raise RuntimeError("refgpu_render: " + err.value.decode())
        fr.ms_kernels = msk
        fr.ms_shipped = mss
        return fr


def have_ref_cpu():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_cpu.so"))


def have_ref_gpu():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_gpu.so"))


def orbit_view(width, height, k, n=64, radius=5.0, pitch_deg=-20.0, target=(0.0, 0.0, 0.0)):
    """orbit camera k of n looking at `target` from `radius`.
    forward from Camera.cpp:10-12; pos = target - radius*forward."""
    pitch = np.float32(pitch_deg * math.pi / 180.0)
    yaw = np.float32(2.0 * math.pi * k / n)
    fwd = (-math.sin(yaw) * math.cos(pitch), math.sin(pitch), -math.cos(yaw) * math.cos(pitch))
    pos = tuple(target[i] - radius * fwd[i] for i in range(3))
    return View(width, height, pos=pos, pitch=float(pitch), yaw=float(yaw))


def oblique_view(width, height):
    """SURVEY.md §8(d) configs 3/4 second camera."""
    return View(width, height, pos=(11.96, 9.30, -1.39), pitch=-0.3981, yaw=0.5713)
