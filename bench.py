#!/usr/bin/env python
"""bench.py — primary rays/s of the CSG raycast path on testCheese512 @ 3840x2160 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our arm (libcsg_b200.so through its C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's own implementation on the host cores

A step = one frame = one pass of the hot path over all width*height primary rays.

N > 1: launched by torchrun, one process per GPU.
  value (device time): the frame is sharded in interleaved 64x32-pixel tiles, every rank's kernel stores its pixels straight
      into rank 0's framebuffer over NVLink (CUDA-IPC peer pointer); no collective on the data path.  The frame is started and
      joined ON THE DEVICE through sync words behind rank 0's framebuffer: a peer's kernels do nothing before rank 0's GPU has
      started the frame, and rank 0's last kernel does not end before every peer has reported its pixels landed.  ms_per_step
      is therefore rank 0's CUDA-event span "frame started on the root GPU -> framebuffer complete on the root GPU" (SURVEY.md
      8d) — peers' work and any lateness of their launches are inside it.
  e2e: csg_render() into ONE host framebuffer (shared memory, page-locked in every rank): the frame is dealt out in rows of
      tiles, every GPU renders its rows and copies them over its own PCIe link.  Time = first rank's call -> last rank's
      return (CLOCK_MONOTONIC is system-wide).
  parity_n: outside the timed region the gathered N-GPU framebuffer (both paths) is compared byte for byte with the frame
      rank 0 renders alone.
Prints ONE JSON line on rank 0; `configs` (every BASELINE.json config at this N) and `parity_n` are its last keys.
"""
import argparse
import hashlib
import json
import math
import mmap
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = "testCheese512"
WIDTH, HEIGHT = 3840, 2160
WORKLOAD = f"Test/{SCENE}.txt @ {WIDTH}x{HEIGHT}, camera (0,0,5) pitch 0 yaw 0 fov 90*3.14159/180, default light, 1 frame per step"
FLOP_KEY = f"{SCENE}@{WIDTH}x{HEIGHT}/default"
SM_COUNT, FP32_LANES = 148, 128


def config_of(n_gpus):
    """`config` of the JSON line — the same for both arms (the reference arm renders the same frame on the host cores)."""
    return {"workload": WORKLOAD,
            "parallelism": f"screen tiles 64x32 interleaved over {n_gpus} GPU(s), NVLink peer stores into rank 0" if n_gpus > 1
                           else "1 GPU",
            "l2": "256 MiB device memset in front of every timed frame, on the frame's stream (L2 flush); inputs are 33 KB of tree + 68 B of camera/light"}


def corpus_scene(name):
    path = os.path.join(ROOT, "oracle", "_ref", "scenes", name + ".txt")
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    return None


def scene_bytes():
    """The reference's scene file (staged by `make -C oracle ref`); falls back to a seeded cheese of our own making."""
    text = corpus_scene(SCENE)
    if text is not None:
        return text, f"reference scene file Test/{SCENE}.txt (no randomness; no dataset or checkpoint involved)"
    import random
    rnd = random.Random(512)
    leaves = [f"Sphere {rnd.uniform(-10, 10):.5f} {rnd.uniform(-10, 10):.5f} {rnd.uniform(-30, -10):.5f} F5F500 {rnd.uniform(0.01, 2):.5f}"
              for _ in range(512)]

    def union(xs):
        if len(xs) == 1:
            return xs[0]
        h = len(xs) // 2
        return "Union\n" + union(xs[:h]) + "\n" + union(xs[h:])
    return ("Difference\nCube 0 0 -20 FFFF00 20\n" + union(leaves)).encode(), "synthetic cheese (512 seeded spheres), reference corpus absent"


def cpu_model():
    """Model name of the host CPU the baseline ran on (printed next to the CPU baseline)."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_threads():
    """Host threads the CPU arms use: every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def source_hash():
    """Hash of the kernel sources; profiles/ncu_summary.json records the one it was captured from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "cuda-csg-tree-raycasting_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h", ".cpp")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def bind_to_gpu_numa_node(local):
    """Run this rank on the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device), so that the host pages it
    first-touches — its rows of the shared host frame — and its launch thread sit on the GPU's own socket."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus of {path}"
    except Exception as e:   # no sysfs / no permission: stay where we are
        return f"unbound ({type(e).__name__})"
    return "unbound"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        # median of the upper half = clocks while the kernels were running (idle gaps between frames sample low)
        return {"sm_mhz": sm[(len(sm) * 3) // 4], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class SharedHost:
    """One host buffer all ranks of the node see: a /dev/shm file mapped by every rank and page-locked (cudaHostRegister) in
    every rank, plus a few words for a spin barrier.  world == 1: plain pinned memory."""

    def __init__(self, g, nbytes, rank, world, dist, tag, pin=True, first_touch=None):
        import numpy as np
        import torch
        self.g, self.pin = g, pin
        self.rank, self.world, self.dist = rank, world, dist
        self.nbytes = nbytes
        self.gen = 0
        self.path = None
        if world == 1:
            self.t = torch.empty(nbytes, dtype=torch.uint8)
            if pin:
                self.t = self.t.pin_memory()
            self.arr = self.t.numpy()
            self.ptr = self.t.data_ptr()
            self.slots = None
            return
        self.path = f"/dev/shm/csg_b200_{os.environ.get('MASTER_PORT', '0')}_{tag}"
        total = nbytes + 4096
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(total)
        dist.barrier()
        self.f = open(self.path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), total)
        whole = np.frombuffer(self.mm, dtype=np.uint8)
        self.arr = whole[:nbytes]
        self.slots = whole[nbytes:nbytes + 8 * world].view(np.int64)
        # touch every page before it is page-locked: first_touch(rank, array) lets every rank touch the part it will write
        # (its rows of the frame), so that those pages land on its own NUMA node; default: rank 0 touches everything
        if first_touch is not None:
            first_touch(rank, self.arr)
            if rank == 0:
                whole[nbytes:] = 0
        elif rank == 0:
            whole[:] = 0
        dist.barrier()
        self.ptr = self.arr.ctypes.data
        if pin:
            g.pin_host_buffer(self.ptr, nbytes)
        dist.barrier()

    def spin_barrier(self):
        """All ranks leave within about a microsecond of each other (NCCL/gloo barriers release ranks tens of us apart)."""
        if self.slots is None:
            return
        self.gen += 1
        self.slots[self.rank] = self.gen
        t0 = time.monotonic()
        while int(self.slots.min()) < self.gen:
            if time.monotonic() - t0 > 60:
                raise RuntimeError("spin barrier timed out")

    def close(self):
        if self.path is None:
            return
        if self.pin:
            self.g.unpin_host_buffer(self.ptr)
        self.dist.barrier()
        self.slots = None
        self.arr = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.f.close()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def dist_setup():
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


class Job:
    """One workload on this rank's GPU: sharded context (+ the NVLink gather target), cameras, light."""

    def __init__(self, g, env, text, width, height, ss=1, optimize=1, cams=None):
        self.g, self.env = g, env
        self.text, self.width, self.height, self.ss = text, width, height, ss
        self.nrays = width * height * ss * ss
        self.scene = g.Scene.parse(text, optimize=optimize)
        self.ctx = self.scene.upload_shard(width, height, env.local, env.rank, env.world)
        if ss > 1:
            self.ctx.set_supersampling(ss)
        self.cams = cams or [g.Camera()]
        self.light = g.Light()
        import torch
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=env.dev)   # the stream the root GPU's share of a frame goes on
        if env.world > 1:   # gather target: rank 0's framebuffer (and its sync words), opened on the other ranks through CUDA IPC
            box = [self.ctx.ipc_handle() if env.rank == 0 else None]
            env.dist.broadcast_object_list(box, src=0)
            if env.rank != 0:
                self.ctx.set_gather_target_ipc(box[0])
            env.dist.barrier()

    def frame_device(self, k=0, prequeued=False, idle_start=False):
        """One frame, device-resident; returns this rank's CUDA-event span (rank 0: the whole sharded frame).
        Default: on the root the L2 flush of the step (a 256 MiB memset) is queued on the context's own stream (csg_stream) and the
        frame right behind it, so the frame's launches are already waiting when the GPU gets to them — the steady state of a render
        loop, and the usual way of timing queued kernels with CUDA events.  The other ranks have flushed before the barrier and
        enqueue their share at once: their kernels wait on the device for the root's start word (two memsets on two GPUs do not end
        together, and a peer whose flush ended after the root's would be late for a reason that has nothing to do with the frame).  idle_start (a diagnostic): the flush has been waited
        for and the frame is enqueued on an idle GPU; whatever the host then takes between the start event and the first launch
        shows up as device time.  prequeued (a diagnostic): the peers enqueue first — their kernels wait on the device for the
        root's start word — and the root enqueues once they have."""
        import torch
        env = self.env
        env.spin()
        if not idle_start and env.rank == 0:
            with torch.cuda.stream(self.stream):
                env.flush.zero_()
        if prequeued and env.rank == 0:
            env.spin()
        self.ctx.enqueue(self.cams[k % len(self.cams)], self.light)
        if prequeued and env.rank != 0:
            env.spin()
        self.ctx.sync()
        return self.ctx.last_frame_ms()

    def time_device(self, steps, warmup, prequeued=False, idle_start=False):
        env = self.env
        ms = []
        t0 = 0.0
        for k in range(warmup + steps):
            if k == warmup:
                env.barrier()
                t0 = time.perf_counter()
            if idle_start or env.rank != 0:
                env.flush.zero_()
            env.barrier()
            t = self.frame_device(k, prequeued, idle_start)
            if k >= warmup:
                ms.append(t)
        env.barrier()
        wall = (time.perf_counter() - t0) * 1e3 / max(steps, 1)
        return ms, wall

    def single_gpu_frame(self, k=0):
        """The same frame rendered by this rank's GPU alone (a separate single-GPU context): what the N-GPU frame must equal."""
        import numpy as np
        one = self.scene.upload_shard(self.width, self.height, self.env.local, 0, 1)
        if self.ss > 1:
            one.set_supersampling(self.ss)
        one.enqueue(self.cams[k % len(self.cams)], self.light)
        img = one.read_framebuffer(np.empty((self.height, self.width, 4), np.uint8))
        one.close()
        return img

    def gathered_frame(self, k=0):
        """The N-GPU frame as it sits in rank 0's framebuffer after one sharded frame."""
        import numpy as np
        self.env.barrier()
        self.frame_device(k)
        self.env.barrier()
        if self.env.rank != 0:
            return None
        return self.ctx.read_framebuffer(np.empty((self.height, self.width, 4), np.uint8))

    def close(self):
        self.ctx.close()
        self.scene.close()


class Env:
    pass


def orbit_cameras(g, n=64, radius=5.0, pitch_deg=-20.0):
    """camera k of n looks at the origin from `radius`; forward as Camera.cpp:10-12 computes it."""
    import numpy as np
    cams = []
    for k in range(n):
        pitch = float(np.float32(pitch_deg * math.pi / 180.0))
        yaw = float(np.float32(2.0 * math.pi * k / n))
        fwd = (-math.sin(yaw) * math.cos(pitch), math.sin(pitch), -math.cos(yaw) * math.cos(pitch))
        cams.append(g.Camera(pos=tuple(-radius * f for f in fwd), pitch=pitch, yaw=yaw))
    return cams


def run_configs(g, env, args):
    """Every BASELINE.json config at this N: a few steps each, device time as for the headline (root-side span), and the
    gathered frame compared with the single-GPU frame."""
    import numpy as np
    import torch
    out = {}
    steps, warmup = 5, 3

    def sharded(key, text, w, h, ss, what):
        if text is None:
            out[key] = {"skipped": "scene corpus not staged"}
            return
        job = Job(g, env, text, w, h, ss=ss)
        ms, _ = job.time_device(steps, warmup)
        got = job.gathered_frame()
        rec = None
        if env.rank == 0:
            ref = job.single_gpu_frame()
            t = float(np.mean(ms))
            rec = {"what": what, "n_gpus": env.world, "ms_per_step": t, "rays_per_s": job.nrays / (t * 1e-3), "steps": steps,
                   "mismatching_bytes_vs_1gpu": int((got != ref).sum())}
        job.close()
        if env.rank == 0:
            out[key] = rec

    sharded("configs[0]", corpus_scene("testWikipedia"), 1920, 1080, 1, "Test/testWikipedia.txt @ 1920x1080, default camera, 1 frame per step")
    if env.world == 1:
        # configs[1]: the 64-frame orbit on one GPU, through csg_render_batch (two pipelined frame slots), frames stay on the device
        text = corpus_scene("testSphereCutByCubesAndCylinder")
        if text is None:
            out["configs[1]"] = {"skipped": "scene corpus not staged"}
        else:
            sc = g.Scene.parse(text)
            ctx = sc.upload(WIDTH, HEIGHT)
            cams = orbit_cameras(g)
            dev = torch.empty(len(cams) * WIDTH * HEIGHT * 4, dtype=torch.uint8, device=env.dev)
            light = g.Light()
            ts = []
            for k in range(2 + 3):
                env.flush.zero_()
                torch.cuda.synchronize()
                ctx.render_batch(cams, light, dev.data_ptr())
                if k >= 2:
                    ts.append(ctx.last_frame_ms())
            t = float(np.mean(ts))
            out["configs[1]"] = {"what": "Test/testSphereCutByCubesAndCylinder.txt @ 3840x2160, 64-frame orbit (R=5, pitch -20 deg) per step, csg_render_batch, frames device-resident",
                                 "n_gpus": 1, "ms_per_step": t, "ms_per_frame": t / len(cams), "rays_per_s": len(cams) * WIDTH * HEIGHT / (t * 1e-3), "steps": 3}
            del dev
            ctx.close()
            sc.close()
        sharded("configs[2]", corpus_scene("testCheese256"), WIDTH, HEIGHT, 1, "Test/testCheese256.txt @ 3840x2160, default camera, 1 frame per step")
    sharded("configs[4]", g.Scene.generate_text(4096, seed=1234), 7680, 4320, 4,
            "synthetic balanced tree, 4096 primitives (csg_generate_scene seed 1234) @ 7680x4320 x 16 rays/pixel, default camera")
    return out


def bench_ours(args):
    import numpy as np
    import torch
    import csg_b200 as g

    rank, world, local = dist_setup()
    if world != args.gpus:
        if rank == 0:
            print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})"}))
        return 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libcsg_b200 has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
    torch.cuda.set_device(local)

    numa = bind_to_gpu_numa_node(local) if world > 1 else "not bound (1 GPU)"
    env = Env()
    env.rank, env.world, env.local, env.dist = rank, world, local, dist
    env.dev = torch.device("cuda", local)
    env.flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.dev)   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    env.barrier = barrier
    # N > 1: a few shared words for a spin barrier — every rank enqueues its share of a frame within about a microsecond of the
    # others (NCCL / gloo barriers release the ranks tens of microseconds apart, and that skew would sit inside the root's span)
    env.sync = SharedHost(g, 4096, rank, world, dist, "sync", pin=False) if world > 1 else None
    env.spin = env.sync.spin_barrier if env.sync else (lambda: None)

    text, data_note = scene_bytes()
    job = Job(g, env, text, WIDTH, HEIGHT, optimize=args.optimize)
    ctx, cam, light = job.ctx, job.cams[0], job.light
    nrays = WIDTH * HEIGHT
    warmup = max(3, args.warmup)

    # ---- warm-up (the clock sampler starts here: nvidia-smi needs a few hundred ms before its first sample), then the timed
    # region: exactly K steps; per-step device time from CUDA events on the launching stream, L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    job.time_device(0, warmup)
    launches0 = ctx.launch_count()
    step_ms, wall_ms = job.time_device(args.steps, 0)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    pre_ms = None
    if world > 1:   # diagnostic: the same frames with the peers' launches queued before the root starts
        pm, _ = job.time_device(min(args.steps, 20), 2, prequeued=True)
        pre_ms = float(np.mean(pm))
    # diagnostic: the same frames enqueued on an idle GPU (flush waited for first): the span then includes the host's launch calls
    im, _ = job.time_device(min(args.steps, 20), 2, idle_start=True)
    idle_ms = float(np.mean(im))
    ms = torch.tensor(step_ms, dtype=torch.float64, device=env.dev)
    lc = torch.tensor([launches], dtype=torch.int64, device=env.dev)
    ms_max = ms.clone()
    if dist is not None:
        dist.broadcast(ms, src=0)                        # the root GPU's span covers every rank's work (device-side gate + join)
        dist.all_reduce(ms_max, op=dist.ReduceOp.MAX)    # per-rank spans, for the record (a peer's includes its wait for the root's start)
        dist.all_reduce(lc, op=dist.ReduceOp.SUM)
    ms = ms.cpu().numpy()
    ms_per_step = float(ms.mean())
    value = nrays / (ms_per_step * 1e-3)

    # ---- e2e: the synchronous public call with a HOST buffer: parameters H2D, kernels, framebuffer D2H into page-locked memory
    def touch_own_rows(r, arr):
        # csg_render deals the frame out in rows of 64x32 tiles, row m to rank m % world: 32 scanlines = 491 520 bytes = 120 pages
        row = 32 * WIDTH * 4
        for m in range(r, (HEIGHT + 31) // 32, world):
            arr[m * row:min((m + 1) * row, arr.size)] = 0
    host = SharedHost(g, nrays * 4, rank, world, dist, "fb", first_touch=touch_own_rows)
    e2e_steps = max(5, min(args.steps, 20))

    def frame_e2e():
        host.spin_barrier()
        t0 = time.monotonic_ns()
        ctx.render(cam, light, host.ptr)
        return t0, time.monotonic_ns()
    for _ in range(3):
        barrier()
        frame_e2e()
    span = []
    for _ in range(e2e_steps):
        env.flush.zero_()
        barrier()
        span.append(frame_e2e())
    t = torch.tensor(span, dtype=torch.int64, device=env.dev)            # [steps, 2]
    t0s, t1s, own = t[:, 0].clone(), t[:, 1].clone(), (t[:, 1] - t[:, 0]).clone()
    if dist is not None:
        dist.all_reduce(t0s, op=dist.ReduceOp.MIN)
        dist.all_reduce(t1s, op=dist.ReduceOp.MAX)
        dist.all_reduce(own, op=dist.ReduceOp.MAX)
    e2e_s = float((t1s - t0s).double().mean().item()) * 1e-9              # first rank's call -> last rank's return
    e2e_own_s = float(own.double().mean().item()) * 1e-9                  # slowest rank's own call
    barrier()
    e2e_frame = host.arr.copy().reshape(HEIGHT, WIDTH, 4) if rank == 0 else None

    # ---- diagnostic next to e2e: the same device->host copies with no rendering at all (every rank moves its own rows of the
    # frame from its framebuffer into the shared host frame, all ranks at once) — what the PCIe links and the host memory take
    d2h_only_s = None
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaMemcpy2D.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
        rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        fb = ctx.framebuffer()
        row_bytes, n_macro = 32 * WIDTH * 4, (HEIGHT + 31) // 32
        mine = list(range(rank, n_macro, world))
        full = [m for m in mine if (m + 1) * 32 <= HEIGHT]
        ragged = [m for m in mine if (m + 1) * 32 > HEIGHT]
        spans = []
        for k in range(8):
            barrier()
            host.spin_barrier()
            t0 = time.monotonic_ns()
            rc = 0
            if full:
                off = full[0] * row_bytes
                rc |= rt.cudaMemcpy2D(host.ptr + off, world * row_bytes, fb + off, world * row_bytes, row_bytes, len(full), 2)
            for m in ragged:
                off = m * row_bytes
                rc |= rt.cudaMemcpy(host.ptr + off, fb + off, (HEIGHT - m * 32) * WIDTH * 4, 2)
            t1 = time.monotonic_ns()
            if rc:
                raise RuntimeError(f"cudaMemcpy2D failed ({rc})")
            if k >= 3:
                spans.append((t0, t1))
        tt = torch.tensor(spans, dtype=torch.int64, device=env.dev)
        c0, c1 = tt[:, 0].clone(), tt[:, 1].clone()
        if dist is not None:
            dist.all_reduce(c0, op=dist.ReduceOp.MIN)
            dist.all_reduce(c1, op=dist.ReduceOp.MAX)
        d2h_only_s = float((c1 - c0).double().mean().item()) * 1e-9
    except Exception as e:   # diagnostic only
        d2h_only_s = None
        if rank == 0:
            print(f"bench.py: d2h-only diagnostic skipped ({type(e).__name__}: {e})", file=sys.stderr)

    # ---- parity of the N-GPU frames with the frame one GPU renders alone (outside the timed regions)
    gathered = job.gathered_frame()
    parity_n = None
    if rank == 0:
        ref = job.single_gpu_frame()
        parity_n = {"n": world, "mismatching_bytes": int((gathered != ref).sum()), "e2e_mismatching_bytes": int((e2e_frame != ref).sum()),
                    "bytes": int(ref.size), "hit_pixels": int((ref.reshape(-1, 4)[:, :3] != np.array([20, 20, 28], np.uint8)).any(axis=1).sum())}
    host.close()

    # ---- extra (not the headline): the same frames with the opt-in view cache — an unchanged camera keeps its per-tile
    # trees, so the pruning kernel is skipped (the reference application's static camera with a moving light)
    static_ms = None
    if world == 1:
        ctx.set_view_cache(True)
        sv = []
        for k in range(13):
            barrier()
            tt = job.frame_device()
            if k >= 3:
                sv.append(tt)
        ctx.set_view_cache(False)
        static_ms = float(np.mean(sv))
    info = ctx.info()
    job.close()

    configs = run_configs(g, env, args) if not args.no_configs else None
    if env.sync:
        env.sync.close()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline (rank 0).  The path is FP32 / instruction-issue bound, not HBM: `frac` is the FP32 work the
    # frame kernel really executed (ncu count of the committed profile of this same command, at this run's frame time)
    # against the FFMA-only probe measured now on this GPU; `issue_frac` is the kernel's issue-slot utilisation from the same
    # profile.  The reference algorithm's flop count is a separate yard-stick (vs_reference_algorithm): the kernels skip ~98 %
    # of that work, so it says how fast the frame is, not how busy the pipes are.
    fpr = None
    try:
        with open(os.path.join(ROOT, "profiles", "flop_per_ray.json")) as f:
            fpr = json.load(f)[FLOP_KEY]["flop_per_ray"]
    except Exception:
        pass
    peak_measured = g.fp32_peak_tflops(local)
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    peak_nominal = SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12
    traffic, executed, achieved, frac, issue_frac = None, None, None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            prof = json.load(f)
        traffic = prof.get("dram_bytes_per_launch")
        fk = prof.get("frame_kernel")
        if fk:
            matches = prof.get("source_hash") == source_hash()
            # the frame kernel's own launch duration, measured live: with the view cache on a step is that one launch (static_ms)
            kernel_ms = static_ms if world == 1 else None
            share = kernel_ms / ms_per_step if kernel_ms else None
            exec_tflops = fk["executed_fp32_flop"] / (kernel_ms * 1e-3) / 1e12 if kernel_ms else None
            achieved = exec_tflops
            frac = exec_tflops / peak_measured if (exec_tflops and peak_measured) else None
            issue_frac = fk["issue_active_pct"] / 100.0
            executed = {"fp32_flop_per_frame": fk["executed_fp32_flop"], "fp32_flop_per_ray": fk["executed_fp32_flop"] / nrays,
                        "warp_instructions_per_frame": fk["warp_instructions"], "frame_kernel_us_in_profile": fk["us"],
                        "frame_kernel_ms_live": kernel_ms, "frame_kernel_share_of_step": share, "fma_pipe_pct": fk["fma_pipe_pct"], "alu_pipe_pct": fk["alu_pipe_pct"],
                        "issue_active_pct": fk["issue_active_pct"], "local_memory_instructions": (fk.get("local_load_instructions") or 0) + (fk.get("local_store_instructions") or 0),
                        "sm_instruction_cache_hit_pct": fk.get("sm_instruction_cache_hit_pct"),
                        "gpc_cache_instruction_requests_pct_of_peak": fk.get("gpc_cache_instruction_requests_pct_of_peak"),
                        "source": "profiles/ncu_summary.json (" + prof.get("tag", "?") + ")",
                        "profile_matches_source": matches, "source_hash": source_hash(), "profile_source_hash": prof.get("source_hash")}
    except Exception:
        pass
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    hbm_achieved = nrays * 4 / (ms_per_step * 1e-3) / 1e9
    ref_alg = value * fpr / 1e12 if fpr else None
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_measured, "unit": "TFLOP/s", "frac": frac, "issue_frac": issue_frac,
                "traffic": traffic,
                "peak_source": "FFMA-only probe kernel measured in this run (csg_fp32_peak_tflops)", "peak_nominal": peak_nominal,
                "executed": executed,
                "vs_reference_algorithm": {"flop_per_ray": fpr, "tflops_equivalent": ref_alg, "ratio_to_peak": (ref_alg / peak_measured) if ref_alg else None,
                                           "note": "rays/s x the REFERENCE algorithm's algorithmic flop/ray (instrumented oracle, profiles/flop_per_ray.json): "
                                                   "a speed-up-over-the-reference's-arithmetic figure, NOT a roofline fraction"},
                "hbm": {"algorithmic_bytes_per_frame": nrays * 4, "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_achieved / hbm_peak, "peak_source": hbm_src,
                        "note": "4 B/ray of RGBA8 out is all the path has to move: a few per cent of HBM bandwidth, hence the FP32/issue roofline"},
                "note": "frac = FP32 flop the frame kernel executed (fadd + fmul + 2 ffma thread instructions, ncu) / its duration / measured FFMA peak; "
                        "the kernel is instruction-issue / latency bound (issue_frac), not FP32-pipe bound; N > 1: per-GPU figures are not derived"}

    # ---- baselines measured beside it (rank 0, N=1 only): the reference on the host cores and the reference CUDA kernel
    cpu_baseline = None
    ref_cuda = None
    if world == 1 and not args.no_baselines:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py
        v = oracle_py.View(WIDTH, HEIGHT)
        nt = host_threads()
        if oracle_py.have_ref_cpu():
            rc = oracle_py.RefCPU()
            step = 4
            sec = rc.render(text, v, row_step=step, nthreads=nt, outputs=False)
            rows = (HEIGHT + step - 1) // step
            cpu_baseline = {"value": rows * WIDTH / sec, "unit": "rays/s", "cores": nt, "cpu_model": cpu_model(), "kind": "reference",
                            "sample": f"every {step}th scanline of the same frame ({rows} rows, {rows * WIDTH} rays, {sec:.2f} s); "
                                      "reference RaycastKernel+LightningKernel source compiled for the host, OpenMP dynamic over rows"}
        else:
            orc = oracle_py.Oracle()
            t0 = time.perf_counter()
            orc.render(text, v, rows=(HEIGHT // 2 - 64, HEIGHT // 2 + 64), nthreads=nt)
            sec = time.perf_counter() - t0
            cpu_baseline = {"value": 128 * WIDTH / sec, "unit": "rays/s", "cores": nt, "cpu_model": cpu_model(), "kind": "port",
                            "sample": "128 centre scanlines, C oracle with OpenMP"}
        if oracle_py.have_ref_gpu():
            rg = oracle_py.RefGPU()
            fr = rg.render(text, v, warmup=2, iters=10, shipped=True, outputs=False)
            mk, msh = float(np.median(fr.ms_kernels)), float(np.median(fr.ms_shipped))
            ref_cuda = {"ms_per_frame_kernels": mk, "ms_per_frame_as_shipped": msh, "rays_per_s": nrays / (mk * 1e-3),
                        "speedup_of_value": (value / (nrays / (mk * 1e-3))),
                        "what": "the reference's own RaycastKernel+LightningKernel rebuilt with nvcc -O3 -arch=sm_100, same frame, this GPU"}

    out = {
        "metric": "primary rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": data_note,
        "config": config_of(world),
        "timing": {"what": "rank 0's CUDA events: frame started on the root GPU -> framebuffer complete on the root GPU; peers' kernels are gated on "
                           "the root's start word and the root's last kernel joins every peer's done word (device-side, over NVLink); every "
                           "step = [L2 flush, start event, the frame's kernels, done event] queued on the root's stream (csg_stream), then waited "
                           "for; the other ranks flush before the step's barrier and enqueue their share while the root's flush runs",
                   "ms_per_step_min": float(ms.min()), "ms_per_step_max": float(ms.max()),
                   "ms_per_step_idle_start": idle_ms,
                   "idle_start_note": "diagnostic: the same frames enqueued on an idle GPU (the flush waited for first): the root's span then also "
                                      "holds the host-side cost of the launch calls between the start event and the first kernel",
                   "ms_per_step_max_over_ranks_own_spans": float(ms_max.cpu().numpy().mean()),
                   "ms_per_step_peers_prequeued": pre_ms,
                   "prequeued_note": "diagnostic only: peers' launches queued (waiting on the device) before the root starts, i.e. the root's "
                                     "span without host-side launch skew between the processes",
                   "host_binding": numa,
                   "wall_ms_per_step_incl_flush_and_barriers": wall_ms},
        "e2e": {"value": nrays / e2e_s, "unit": "rays/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": 256 * world,
                "d2h_bytes_per_step": nrays * 4, "steps": e2e_steps, "ms_per_step_slowest_rank_own_call": e2e_own_s * 1e3,
                "d2h_only_ms": None if d2h_only_s is None else d2h_only_s * 1e3,
                "d2h_only_note": "diagnostic: the same device->host copies (every rank its own rows, all ranks at once) with no rendering — the PCIe / "
                                 "host-memory floor of e2e on this box",

                "what": "csg_render() into one page-locked host RGBA8 frame (N > 1: shared memory registered in every rank): camera/light as kernel "
                        "parameters, the frame dealt out in rows of 64x32 tiles, every GPU renders its rows in bands and copies each band D2H over its "
                        "own PCIe link while the next band renders; time = first rank's call -> last rank's return (CLOCK_MONOTONIC)"},
        "gpu_launches": int(lc.item()),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "ref_cuda_baseline": ref_cuda,
        "static_view": None if static_ms is None else {
            "ms_per_step": static_ms, "value": nrays / (static_ms * 1e-3), "unit": "rays/s",
            "what": "NOT the headline: csg_set_view_cache(1), camera unchanged between frames -> per-tile trees reused, 1 launch per frame"},
        "details": {"launches_per_frame": "2 per GPU: csg_prune_flat_kernel (per-tile pruned trees, rebuilt every frame) + csg_frame_kernel, "
                                          "chained by programmatic dependent launch; both inside the timed region",
                    "optimize": args.optimize, "launch": info},
        "configs": configs,
        "parity_n": parity_n,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref/libref_cpu.so: its RaycastKernel + LightningKernel
    source compiled for the host, OpenMP over rows) on every host core, same frame, same metric: W warm-up frames, then
    exactly K full frames.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    text, data_note = scene_bytes()
    v = oracle_py.View(WIDTH, HEIGHT)
    nt = host_threads()
    if oracle_py.have_ref_cpu():
        rc = oracle_py.RefCPU()
        kind = "reference"
        run = lambda: rc.render(text, v, nthreads=nt, outputs=False)  # noqa: E731
    else:
        orc = oracle_py.Oracle()
        kind = "port"

        def run():
            t0 = time.perf_counter()
            orc.render(text, v, nthreads=nt, want_rgba=True)
            return time.perf_counter() - t0
    for _ in range(args.warmup):
        run()
    secs = [run() for _ in range(args.steps)]
    sec = sum(secs) / len(secs)
    value = WIDTH * HEIGHT / sec
    out = {"impl": "reference", "metric": "primary rays/s", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_note,
           "config": config_of(args.gpus),
           "cpu_baseline": {"value": value, "unit": "rays/s", "cores": nt, "cpu_model": cpu_model(), "kind": kind,
                            "sample": f"the whole frame, every step ({HEIGHT} rows, {WIDTH * HEIGHT} rays); OMP_NUM_THREADS ignored: num_threads({nt})"},
           "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--optimize", type=int, default=1, help="load-time tree optimisation level (csg_scene_set_optimize)")
    ap.add_argument("--no-baselines", action="store_true", help="skip the CPU / reference-CUDA baselines")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configs")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 5:   # the defaults are sized for the GPU arm: a full CPU frame takes seconds
            args.steps, args.warmup = 5, 1
        return bench_reference(args)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
