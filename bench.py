#!/usr/bin/env python
"""bench.py — primary rays/s of the CSG raycast path on testCheese512 @ 3840x2160 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our arm (libcsg_b200.so through its C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference's own implementation on the host cores

A step = one frame = one pass of the hot path over all width*height primary rays.
N > 1: launched by torchrun, one process per GPU; the frame is sharded in interleaved 64x32-pixel tiles, every rank's
kernel stores its pixels straight into rank 0's framebuffer over NVLink (CUDA-IPC peer pointer); no collective on the
data path.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
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


def scene_bytes():
    """The reference's scene file (staged by `make -C oracle ref`); falls back to a seeded cheese of our own making."""
    path = os.path.join(ROOT, "oracle", "_ref", "scenes", SCENE + ".txt")
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read(), f"reference scene file Test/{SCENE}.txt (no randomness; no dataset or checkpoint involved)"
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


def dist_setup(n):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def bench_ours(args):
    import numpy as np
    import torch
    import csg_b200 as g

    rank, world, local = dist_setup(args.gpus)
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
    dev = torch.device("cuda", local)

    text, data_note = scene_bytes()
    scene = g.Scene.parse(text, optimize=args.optimize)
    ctx = scene.upload_shard(WIDTH, HEIGHT, local, rank, world)
    cam, light = g.Camera(), g.Light()
    nrays = WIDTH * HEIGHT

    # gather target: rank 0's framebuffer, opened on the other ranks through CUDA IPC (NVLink peer stores)
    if world > 1:
        box = [ctx.ipc_handle() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            ctx.set_gather_target_ipc(box[0])

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def frame_device():
        ctx.enqueue(cam, light)
        ctx.sync()
        return ctx.last_frame_ms()

    # ---- warm-up (the clock sampler starts here: nvidia-smi needs a few hundred ms before its first sample)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(max(3, args.warmup)):
        flush.zero_()
        barrier()
        frame_device()
    barrier()

    # ---- timed: exactly K steps; per-step device time from CUDA events on the launching stream, L2 flushed between steps
    launches0 = ctx.launch_count()
    step_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        step_ms.append(frame_device())
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    ms = torch.tensor(step_ms, dtype=torch.float64, device=dev)
    lc = torch.tensor([launches], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)    # a frame is done when the slowest shard is done
        dist.all_reduce(lc, op=dist.ReduceOp.SUM)
    ms = ms.cpu().numpy()
    ms_per_step = float(ms.mean())
    value = nrays / (ms_per_step * 1e-3)

    # ---- e2e: the synchronous public call with a HOST buffer: parameters H2D, kernel(s), framebuffer D2H into pinned memory
    host_fb = torch.empty(nrays * 4, dtype=torch.uint8).pin_memory() if rank == 0 else None
    e2e_steps = max(5, min(args.steps, 20))

    def frame_e2e():
        if world == 1:
            ctx.render(cam, light, host_fb.data_ptr())
        else:
            ctx.enqueue(cam, light)
            ctx.sync()
            dist.barrier()                      # all shards have landed in rank 0's framebuffer
            if rank == 0:
                ctx.read_framebuffer(host_fb.data_ptr())
    for _ in range(3):
        barrier()
        frame_e2e()
    e2e_t = []
    for _ in range(e2e_steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        frame_e2e()
        torch.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    e2e = torch.tensor(e2e_t, dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e.cpu().numpy().mean())

    # ---- extra (not the headline): the same frames with the opt-in view cache — an unchanged camera keeps its per-tile
    # trees, so the pruning kernel is skipped (the reference application's static camera with a moving light)
    static_ms = None
    if world == 1:
        ctx.set_view_cache(True)
        sv = []
        for k in range(13):
            flush.zero_()
            barrier()
            t = frame_device()
            if k >= 3:
                sv.append(t)
        ctx.set_view_cache(False)
        static_ms = float(np.mean(sv))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline (rank 0): FP32-bound path.  achieved = rays/s x algorithmic flop/ray of the REFERENCE
    # algorithm (instrumented oracle, profiles/flop_per_ray.json); peak = FFMA-only probe measured now on this GPU.
    fpr = None
    try:
        with open(os.path.join(ROOT, "profiles", "flop_per_ray.json")) as f:
            fpr = json.load(f)[FLOP_KEY]["flop_per_ray"]
    except Exception:
        pass
    peak_measured = g.fp32_peak_tflops(local)
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    peak_nominal = SM_COUNT * FP32_LANES * 2 * sm_max * 1e6 / 1e12
    traffic, executed = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            prof = json.load(f)
        traffic = prof.get("dram_bytes_per_launch")
        fk = prof.get("frame_kernel")
        if fk:
            # what the frame kernel actually issued (ncu capture of this same command, committed under profiles/): FP32 flop per
            # frame, pipe and issue-slot utilisation.  The frame rate is this run's; the per-frame counts are the profile's.
            exec_tflops = fk["executed_fp32_flop"] / (ms_per_step * 1e-3) / 1e12
            executed = {"fp32_flop_per_frame": fk["executed_fp32_flop"], "fp32_flop_per_ray": fk["executed_fp32_flop"] / nrays,
                        "achieved": exec_tflops, "frac": exec_tflops / peak_measured if peak_measured else None,
                        "fma_pipe_pct": fk["fma_pipe_pct"], "alu_pipe_pct": fk["alu_pipe_pct"], "issue_active_pct": fk["issue_active_pct"],
                        "warp_instructions_per_frame": fk["warp_instructions"], "source": "profiles/ncu_summary.json (" + prof.get("tag", "?") + ")"}
    except Exception:
        pass
    achieved = value * fpr / 1e12 if fpr else None
    # the HBM view of the same frame, to show why it is not the bound: algorithmic bytes = the RGBA8 framebuffer (4 B/ray)
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    hbm_achieved = nrays * 4 / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_measured, "unit": "TFLOP/s",
                "frac": (achieved / peak_measured) if achieved else None, "traffic": traffic,
                "peak_source": "FFMA-only probe kernel measured in this run (csg_fp32_peak_tflops)",
                "peak_nominal": peak_nominal, "flop_per_ray": fpr, "executed": executed,
                "hbm": {"algorithmic_bytes_per_frame": nrays * 4, "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_achieved / hbm_peak, "peak_source": hbm_src,
                        "note": "4 B/ray of RGBA8 out is all the path has to move: a few per cent of HBM bandwidth, hence the FP32/issue roofline"},
                "note": "achieved = rays/s x the REFERENCE algorithm's algorithmic flop/ray (fixed yard-stick, SURVEY.md 8d).  Our kernels skip "
                        "almost all of that work (per-tile pruned trees, nearest-hit search, tighter boxes), so frac exceeds 1: it measures the "
                        "frame against doing the reference's arithmetic at FP32 peak.  'executed' is what the frame kernel really issued: it is "
                        "issue/latency-bound, not FP32-bound."}

    # ---- baselines measured beside it (rank 0, N=1 only): the reference on the host cores and the reference CUDA kernel
    cpu_baseline = None
    ref_cuda = None
    if world == 1 and not args.no_baselines:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py
        v = oracle_py.View(WIDTH, HEIGHT)
        if oracle_py.have_ref_cpu():
            rc = oracle_py.RefCPU()
            step = 4
            sec = rc.render(text, v, row_step=step, outputs=False)
            rows = (HEIGHT + step - 1) // step
            cpu_baseline = {"value": rows * WIDTH / sec, "unit": "rays/s", "cores": rc.max_threads(), "cpu_model": cpu_model(), "kind": "reference",
                            "sample": f"every {step}th scanline of the same frame ({rows} rows, {rows * WIDTH} rays, {sec:.2f} s); "
                                      "reference RaycastKernel+LightningKernel source compiled for the host, OpenMP dynamic over rows"}
        else:
            orc = oracle_py.Oracle()
            t0 = time.perf_counter()
            fr = orc.render(text, v, rows=(HEIGHT // 2 - 64, HEIGHT // 2 + 64))
            sec = time.perf_counter() - t0
            cpu_baseline = {"value": 128 * WIDTH / sec, "unit": "rays/s", "cores": os.cpu_count(), "cpu_model": cpu_model(), "kind": "port",
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
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": data_note,
        "config": {"workload": WORKLOAD, "parallelism": f"screen tiles 64x32 interleaved over {world} GPU(s), NVLink peer stores into rank 0",
                   "l2": "256 MiB device memset between timed frames (L2 flush); inputs are 33 KB of tree + 68 B of camera/light",
                   "launches_per_frame": "2 per GPU: csg_prune_flat_kernel (per-tile pruned trees, rebuilt every frame) + csg_frame_kernel, "
                                         "chained by programmatic dependent launch; both inside the timed region",
                   "optimize": args.optimize, "launch": ctx.info()},
        "e2e": {"value": nrays / e2e_s, "unit": "rays/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": 256,
                "d2h_bytes_per_step": nrays * 4, "steps": e2e_steps,
                "what": "csg_render() with a pinned host RGBA8 buffer: camera/light as kernel parameters, the frame rendered in bands of tile rows (6 at 4K), each band copied D2H (33 MB in all, PCIe-bound) while the next one renders"},
        "gpu_launches": int(lc.item()),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "ref_cuda_baseline": ref_cuda,
        "static_view": None if static_ms is None else {
            "ms_per_step": static_ms, "value": nrays / (static_ms * 1e-3), "unit": "rays/s",
            "what": "NOT the headline: csg_set_view_cache(1), camera unchanged between frames -> per-tile trees reused, 1 launch per frame"},
        "ms_per_step_min": float(ms.min()), "ms_per_step_max": float(ms.max()),
        "wall_ms_per_step_incl_flush": (t_wall1 - t_wall0) * 1e3 / args.steps,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref/libref_cpu.so: its RaycastKernel + LightningKernel
    source compiled for the host, OpenMP over rows), all host threads, same config and metric.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    text, data_note = scene_bytes()
    v = oracle_py.View(WIDTH, HEIGHT)
    step = 8   # bounded sample: every 8th scanline of the frame per step
    rows = (HEIGHT + step - 1) // step
    if oracle_py.have_ref_cpu():
        rc = oracle_py.RefCPU()
        kind, cores = "reference", rc.max_threads()
        run = lambda: rc.render(text, v, row_step=step, outputs=False)  # noqa: E731
    else:
        orc = oracle_py.Oracle()
        kind, cores = "port", os.cpu_count()
        rows = 128

        def run():
            t0 = time.perf_counter()
            orc.render(text, v, rows=(HEIGHT // 2 - 64, HEIGHT // 2 + 64), want_rgba=True)
            return time.perf_counter() - t0
    for _ in range(min(args.warmup, 1)):
        run()
    steps = max(1, min(args.steps, 5))
    secs = [run() for _ in range(steps)]
    sec = sum(secs) / len(secs)
    value = rows * WIDTH / sec
    sample = f"every {step}th scanline of the frame per step ({rows} rows, {rows * WIDTH} rays)"
    out = {"impl": "reference", "metric": "primary rays/s", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3 * (HEIGHT / rows) if kind == "reference" else None,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_note,
           "config": {"workload": WORKLOAD},
           "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "cpu_model": cpu_model(), "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "ms_per_step is the sample time scaled to a full frame"}
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
    args = ap.parse_args()
    if args.impl == "reference":
        return bench_reference(args)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
