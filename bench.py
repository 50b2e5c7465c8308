#!/usr/bin/env python3
"""bench.py -- headline benchmark of the cuda OSPRay device (BASELINE.json: Mrays/s and frame ms, 1024x1024
pathtracer cornell_box, 1 spp per frame, at 1/2/4/8 B200; CPU baseline beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W      (N > 1)

One step = one ospRenderFrame of workload C1 (progressive accumulation, frameID advances every step).
N > 1: one process per GPU; the scene is replicated, image tiles are partitioned `tile % N == rank`
(modules/multiDevice semantics), each rank converts its tiles to sRGBA8 and stores them straight into GPU 0's image
over NVLink peer memory (ospray_b200.multigpu.PeerTileExchange: one push kernel per rank and a flag handshake per
frame; `--gather nccl` = pack + one NCCL gather + unpack instead) -- total work is fixed, so this is STRONG scaling.

`value`     device-timed Mrays/s: every ray handed to traversal (primary + continuation + shadow segments) on
            all ranks / (max over ranks of the summed per-step CUDA-event time: renderer kernel + accumulation +
            colour conversion [+ the tile exchange]); scene and framebuffer resident in HBM.
`e2e`       the same metric through the public API with host buffers, wall clock per step: ospCommit(camera)
            (the step's host-side inputs; they reach the device as the kernel parameter block), ospRenderFrame,
            ospWait, ospMapFrameBuffer(OSP_FB_COLOR) = D2H of the finished sRGBA8 image into pinned memory.
`roofline`  dominant kernel = the renderer megakernel; algorithmic bytes per DESIGN.md section 4.
`cpu_baseline` / `--impl reference`: the reference's own PathTracer kernels, compiled from /root/reference in their scalar
            C++ mode into oracle/_ref/libospref.so (oracle/build_ref.py; kind "reference"), on all host threads; falls
            back to the scalar oracle port (oracle/oracle.cpp, kind "port") when that library did not travel. The
            reference's ISPC/Embree CPU device itself cannot be built here (no ISPC / Embree / Open VKL / rkcommon)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C1: ospray_testing cornell_box, pathtracer, 1024x1024, 1 spp/frame, OSP_FB_SRGBA COLOR|ACCUM|DEPTH"
# Algorithmic HBM bytes per unit of each kernel class: the path / ray state the launch has to read and write
# once. Hierarchy nodes and leaf records are NOT in these figures: on C1 (34 quads) they live in L1 / L2, and counting
# them as HBM traffic would credit cache hits to the memory roofline.
ALGO_BYTES = {
    "GEN": 132,     # per path: flags 4 + thr/con/lastP/lastN/sZ 80 + smp 8 + ray 32 + queue entry 4, all written
    "TRACE": 84,    # per ray: queue entry 4 + ray 32 read; continuation: hit 36 B written; shadow: folded into its path (shC 16 +
                    # flags 4 + con 16 read, con 16 + flags 4 written = 56 B); C1 mix ~ 60 % continuation: mean 48
    "SHADE": 320,   # per path vertex: state read 148 + state written 112 + shadow ray emitted 52 + frame 8 (the fold moved to TRACE)
    "TAIL": 420,    # per path vertex carried by the tail kernel: SHADE's state + the ray / hit round trip it keeps in L1
    "POST": 20,     # per pixel: float4 accumulation buffer read 16 + sRGBA8 written 4
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C1", choices=sorted(WORKLOADS),
                    help="C1 (default) is the configuration BASELINE.json's metric is quoted on; the others are BASELINE.json's "
                         "larger configurations, timed the same way (their CPU arm is not run: the scalar checker needs minutes per frame)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = every rank stores its tiles into GPU 0's image over NVLink peer memory (CUDA IPC); "
                         "'nccl' = pack + one NCCL gather + unpack")
    return ap.parse_args()


WORKLOADS = {
    # name: (description, renderer, width, height, spp per step)
    "C1": (WORKLOAD, "pathtracer", 1024, 1024, 1),
    "C2": ("C2: ospray_testing gravity_spheres_volume 128^3, scivis, 1920x1080, 4 spp", "scivis", 1920, 1080, 4),
    "C3": ("C3: 9.97 M-triangle procedural mesh (94^3 boxes), ao, 3840x2160, 16 spp", "ao", 3840, 2160, 16),
    "C4": ("C4: ospray_testing random_spheres (thinGlass) + streamlines, pathtracer, 1920x1080, 1 spp/frame of the 64 accumulated", "pathtracer", 1920, 1080, 1),
    "C5": ("C5: cornell_box + 9.97 M-triangle mesh, pathtracer, 3840x2160, 1 spp/frame of the 64 accumulated", "pathtracer", 3840, 2160, 1),
}


def workload_spec(name):
    from ospray_b200 import testing as T
    merged = lambda a, b: T.SceneSpec(instances=a.instances + b.instances, lights=a.lights)
    if name == "C1":
        return T.cornell_box()
    if name == "C2":
        return T.gravity_spheres_volume(128)
    if name == "C3":
        return T.box_mesh(94)
    if name == "C4":
        return merged(T.random_spheres(renderer_type="pathtracer"), T.streamlines())
    return merged(T.cornell_box(), T.box_mesh(94))


def kernel_class_profile():
    """profiles/r2_c1_kernel_classes.json: measured DRAM bytes, lanes per instruction and issue-slot utilisation per kernel
    class of one C1 frame (ncu, tools/ncu_classes.py), committed with the code it was taken from"""
    p = os.path.join(ROOT, "profiles", "r2_c1_kernel_classes.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def fb_ops_4k(peak):
    """the framebuffer pixel ops alone at 3840x2160 (the clean HBM-bound kernels): ospcuda_fb_accumulate over
    colour + depth + variance and the fused tonemap + sRGBA8 conversion, CUDA events on the library's stream, L2 flushed"""
    import ctypes as C
    import torch
    from ospray_b200 import lib_path

    class OcFrameBuffer(C.Structure):  # include/ospcuda.h
        _fields_ = [("sizeX", C.c_int32), ("sizeY", C.c_int32), ("frameID", C.c_int32), ("channels", C.c_uint32),
                    ("accumulateVariance", C.c_uint32), ("projectedDepth", C.c_uint32), ("targetFrames", C.c_int32),
                    ("colorFormat", C.c_uint32), ("color", C.c_void_p), ("variance", C.c_void_p), ("depth", C.c_void_p),
                    ("normal", C.c_void_p), ("firstNormal", C.c_void_p), ("albedo", C.c_void_p), ("position", C.c_void_p),
                    ("primID", C.c_void_p), ("objID", C.c_void_p), ("instID", C.c_void_p), ("numRanks", C.c_uint32),
                    ("rank", C.c_uint32), ("taskError", C.c_void_p), ("errorThreshold", C.c_float),
                    ("rayCounters", C.c_void_p), ("control", C.c_void_p)]

    class OcToneMapper(C.Structure):
        _fields_ = [("exposure", C.c_float), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float), ("d", C.c_float), ("acesColor", C.c_uint32)]
    L = C.CDLL(lib_path("libospcuda.so"))
    w, h = 3840, 2160
    n = w * h
    dev = torch.device("cuda")
    color = torch.rand(n, 4, device=dev)
    var = torch.rand(n, 4, device=dev)
    depth = torch.full((n,), 3.0, device=dev)
    sample = torch.rand(n, 4, device=dev)
    sdepth = torch.rand(n, device=dev)
    out8 = torch.empty(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fb = OcFrameBuffer()
    fb.sizeX, fb.sizeY, fb.frameID, fb.accumulateVariance = w, h, 3, 1
    fb.color, fb.variance, fb.depth = color.data_ptr(), var.data_ptr(), depth.data_ptr()
    fb.numRanks, fb.rank = 1, 0
    tm = OcToneMapper(1.0, 1.6773, 1.1174272, 0.2446756, 0.9714, 1)  # the tonemapper image operation's default curve
    L.ospcuda_fb_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ospcuda_fb_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.ospcuda_stream.restype = C.c_void_p
    stream = torch.cuda.ExternalStream(L.ospcuda_stream())
    res = {}
    for name, call, nbytes in (
            ("accumulate", lambda: L.ospcuda_fb_accumulate(C.byref(fb), sample.data_ptr(), sdepth.data_ptr()), n * (16 + 32 + 32 + 4 + 8)),
            ("tonemap+convert", lambda: L.ospcuda_fb_postprocess(color.data_ptr(), None, out8.data_ptr(), n, 2, C.byref(tm)), n * (16 + 4))):
        times = []
        for i in range(8):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            if call() != 0:
                raise RuntimeError(name)
            b.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                times.append(a.elapsed_time(b))
        ms = sum(times) / len(times)
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": round(ms, 4), "algorithmic_bytes": nbytes, "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
    res["size"] = "3840x2160"
    return res


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"], out["samples"] = sorted(reasons), len(sm)
        return out


def oracle_setup(width, height):
    from oracle import pyoracle as O
    from ospray_b200 import testing as T
    import numpy as np
    scene = O.Scene(T.cornell_box())
    pose = T.arcball_camera(*scene.bounds)
    cam = O.make_camera(pose["position"], pose["direction"], pose["up"], aspect=np.float32(width) / np.float32(height))
    ren = O.RendererParams("pathtracer", pixelSamples=1)
    return O, scene, cam, ren


def cpu_backend():
    """("ref", "reference") when oracle/_ref/libospref.so (the reference's own kernels, oracle/build_ref.py) travelled
    with the snapshot and the frame is a multiple of its 8x8 render task, else ("oracle", "port")"""
    from oracle import pyoracle as O
    return ("ref", "reference") if O.ref_available() else ("oracle", "port")


CPU_NOTE = {"reference": "the reference's own PathTracer kernels (modules/cpu/**/*.ispc compiled in their scalar C++ mode, oracle/build_ref.py; "
                         "brute-force scene queries stand in for Embree), one pixel per work item on all host threads",
            "port": "scalar CPU port of the reference algorithm (oracle/oracle.cpp, brute-force intersection)"}


def oracle_frames(width, height, frames, warm, stride=1, backend="oracle"):
    """time `frames` CPU frames on all host threads; stride k renders the tiles with tileID % k == 0 only
    (a bounded sample of the same workload). Returns (rays, seconds, threads)."""
    O, scene, cam, ren = oracle_setup(width, height)
    fb = O.HostFrameBuffer(width, height, ("color", "depth"), num_ranks=stride, rank=0)
    threads = O.lib().oracle_hardware_threads()
    for _ in range(warm):
        O.render_frame(scene, cam, ren, fb, threads, backend=backend)
    rays, secs = 0, 0.0
    for _ in range(frames):
        t0 = time.perf_counter()
        O.render_frame(scene, cam, ren, fb, threads, backend=backend)
        secs += time.perf_counter() - t0
        rays += int(fb.counters.sum())
    return rays, secs, threads


def calibrate_stride(width, height, budget_s, backend="oracle"):
    """pick the tile stride so that one CPU step stays under budget_s on this host"""
    rays, secs, _ = oracle_frames(width, height, 1, 0, stride=16, backend=backend)
    full = secs * 16
    stride = 1
    while full / stride > budget_s and stride < 64:
        stride *= 2
    return stride


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h = args.width, args.height
    backend, kind = cpu_backend()
    if w % 8 or h % 8:
        backend, kind = "oracle", "port"
    stride = calibrate_stride(w, h, 4.0, backend)
    _, _, threads = oracle_frames(w, h, 0, 0)
    O, scene, cam, ren = oracle_setup(w, h)
    fb = O.HostFrameBuffer(w, h, ("color", "depth"), num_ranks=stride, rank=0)
    for _ in range(args.warmup):
        O.render_frame(scene, cam, ren, fb, threads, backend=backend)
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.render_frame(scene, cam, ren, fb, threads, backend=backend)
        secs += time.perf_counter() - t0
        rays += int(fb.counters.sum())
    v = rays / secs / 1e6
    sample = f"each step = tiles with tileID % {stride} == 0 of the {w}x{h} frame ({100.0 / stride:.1f}% of the pixels), all {threads} host threads"
    ms_full = secs / args.steps * stride * 1e3
    line = {"impl": "reference", "metric": "Mrays/s", "value": round(v, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_full, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": CPU_NOTE[kind] + "; the reference's ISPC/Embree CPU device itself is unbuildable "
                       "here; ms_per_step is scaled to the full frame"},
            "cpu_baseline": {"value": round(v, 3), "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(v, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from ospray_b200 import capi as osp, multigpu, testing as T

    if args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        # convenience: relaunch under torchrun (the driver launches us that way itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the cuda OSPRay device has no CPU fallback (use --impl reference for the CPU arm)")
    # stdout carries exactly ONE JSON line: libraries that print there on their own (NCCL's version banner goes to stdout
    # whatever NCCL_DEBUG_FILE says on some boxes) are pointed at stderr for the lifetime of the process, the line itself
    # is written to a duplicate of the original descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank, world, local = multigpu.init_process_group_from_env()
    torch.cuda.set_device(local)
    wl_name, ren_kind, wl_w, wl_h, spp = WORKLOADS[args.workload]
    w, h = (args.width, args.height) if args.workload == "C1" else (wl_w, wl_h)
    K, W = args.steps, max(args.warmup, 3)
    osp.init(cudaDevice=local, numSubdevices=world, subdeviceRank=rank)
    stream = torch.cuda.Stream()
    osp.set_render_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        launches0 = osp.kernel_launch_count()
        t_setup = time.perf_counter()
        world_obj, keep = T.to_ospray(workload_spec(args.workload))
        setup_ms = (time.perf_counter() - t_setup) * 1e3
        cam, pose = T.make_camera(world_obj, w, h)
        ren = T.make_renderer(ren_kind, pixelSamples=spp)
        fb = osp.FrameBuffer(w, h, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH)
        conv_ptr, bpp = fb.device_pointer(osp.OSP_FB_COLOR)
        assert bpp == 4
        if world == 1:
            gather = None
        else:
            gather, gather_note = None, None
            if args.gather == "peer":  # tiles stored straight into GPU 0's image over NVLink peer memory (one kernel per rank)
                try:
                    gather = multigpu.PeerTileExchange(w, h, 4, conv_ptr)
                except multigpu.PeerSetupError as e:  # raised on every rank together: all ranks take the NCCL gather
                    gather_note = f"peer mapping unavailable ({e}); "
            if gather is None:         # pack + one NCCL gather + unpack
                gather = multigpu.TileGather(w, h, 4, torch.device("cuda", local), ops=multigpu.CudaTileOps(shared_stream=True))
        begin = getattr(gather, "begin_frame", lambda: None)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

        def step(timed_events=None):
            if timed_events is not None:
                timed_events[0].record(stream)
            begin()
            f = fb.render_frame(ren, cam, world_obj)
            if gather is not None:
                gather.gather(conv_ptr)
            if timed_events is not None:
                timed_events[1].record(stream)
            return f

        # ---- device-timed leg ------------------------------------------------------------------------
        for _ in range(W):
            step().wait()
            flush.zero_()
        clocks = ClockSampler(local) if rank == 0 else None
        time.sleep(0.3 if rank == 0 else 0.0)  # let nvidia-smi start sampling (before the barrier: ranks start the timed steps together)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        rays = np.zeros(3, np.int64)
        kernel_ms, post_ms = 0.0, 0.0
        launches1 = osp.kernel_launch_count()
        for i in range(K):
            f = step(ev[i])
            f.wait()
            rays += np.array(fb.ray_counts(), np.int64)
            a, b = f.stage_ms()
            kernel_ms += a
            post_ms += b
            flush.zero_()  # L2 flush between timed iterations (outside the event brackets)
        torch.cuda.synchronize()
        launches = osp.kernel_launch_count() - launches1

        # ---- roofline leg: per-launch-class CUDA-event times of the wavefront kernels (same workload, 3 frames) ----
        import ctypes as C
        from ospray_b200 import lib_path
        LC = C.CDLL(lib_path("libospcuda.so"))
        LC.ospcuda_profile_enable(1)
        prof_rays = np.zeros(3, np.int64)
        prof_fused_rays = 0
        PF = 3
        for i in range(PF):
            f = step()
            f.wait()
            prof_rays += np.array(fb.ray_counts(), np.int64)
            prof_fused_rays += fb.fused_ray_count()
            flush.zero_()
        cls_ms, cls_n = (C.c_float * 4)(), (C.c_uint32 * 4)()
        LC.ospcuda_profile_read(cls_ms, cls_n)
        LC.ospcuda_profile_enable(0)
        cls_ms, cls_n = [float(x) / PF for x in cls_ms], [int(x) // PF for x in cls_n]
        if world > 1:
            dist.barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        clk = clocks.stop() if clocks else None

        # ---- end-to-end leg (public API, host buffers, wall clock) -------------------------------------
        e2e_rays = np.zeros(3, np.int64)
        for i in range(W):                                       # warm the host path as well: first ospMapFrameBuffer pins its
            cam.set("position", pose["position"]).commit()       # staging buffer, the allocator and the page cache settle
            begin()
            f = fb.render_frame(ren, cam, world_obj)
            if gather is not None:
                gather.gather(conv_ptr)
            f.wait()
            if rank == 0:
                with fb.mapped(osp.OSP_FB_COLOR) as img:
                    e2e_probe = int(img[h // 2, w // 2, 0])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            cam.set("position", pose["position"]).commit()       # the step's host inputs
            begin()
            f = fb.render_frame(ren, cam, world_obj)
            if gather is not None:
                gather.gather(conv_ptr)
            f.wait()
            if rank == 0:
                with fb.mapped(osp.OSP_FB_COLOR) as img:           # ospMapFrameBuffer: D2H of the finished frame into pinned
                    e2e_probe = int(img[h // 2, w // 2, 0])        # host memory, read in place like a C application would
            e2e_rays += np.array(fb.ray_counts(), np.int64)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

        # N > 1: the image composed on GPU 0 must hold, tile for tile, what every rank rendered (checksums of each rank's
        # owned tiles, taken from its own display buffer, against the same tiles of GPU 0's image)
        composed_ok = multigpu.verify_composition(fb.map(osp.OSP_FB_COLOR)) if world > 1 else None

        tot = torch.tensor([float(rays.sum()), float(e2e_rays.sum()), float(launches)], dtype=torch.float64, device="cuda")
        mx = torch.tensor([dev_ms, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        tot, mx = tot.cpu().numpy(), mx.cpu().numpy()
    if rank == 0:
        peak, peak_src = peaks()
        value = tot[0] / (mx[0] * 1e-3) / 1e6
        e2e = tot[1] / mx[1] / 1e6
        # ---- roofline, per kernel class ------------------------------------------------------------------
        # live: CUDA-event time per class (profile leg above) and the units each class processed; algorithmic bytes =
        # units x ALGO_BYTES; measured DRAM bytes / lanes / issue utilisation per class come from the committed ncu capture
        # of the same command (profiles/r2_c1_kernel_classes.json, tools/ncu_classes.py) -- bench.py never runs under ncu.
        trace_ms, shade_ms, gen_ms, tail_ms = cls_ms
        rays_frame = prof_rays / PF
        tail_rays = prof_fused_rays / PF
        vertices = rays_frame[0] + rays_frame[1]                   # every primary / continuation ray ends in one shaded vertex
        tail_vertices = vertices * (tail_rays / max(rays_frame.sum(), 1.0))
        units = {"GEN": float(rays_frame[0]), "TRACE": float(rays_frame.sum() - tail_rays), "SHADE": float(vertices - tail_vertices),
                 "TAIL": float(tail_vertices), "POST": float(w * h)}
        ms = {"GEN": gen_ms, "TRACE": trace_ms, "SHADE": shade_ms, "TAIL": tail_ms, "POST": post_ms / K}
        nl = {"GEN": cls_n[2], "TRACE": cls_n[0], "SHADE": cls_n[1], "TAIL": cls_n[3], "POST": 1}
        what = {"GEN": "ptGenKernel (sampler + camera ray per path)", "TRACE": "ptTraceKernel (ray queue: hierarchy traversal + intersection)",
                "SHADE": "ptShadeKernel (shadow fold, postIntersect, BSDF, light sampling, roulette)",
                "TAIL": "ptTailKernel (per-lane completion of the thinned-out frame)", "POST": "postprocessKernel (float4 -> sRGBA8)"}
        measured = kernel_class_profile() if args.workload == "C1" and (w, h) == (1024, 1024) and world == 1 else None
        classes = []
        for c in ("GEN", "TRACE", "SHADE", "TAIL", "POST"):
            if ms[c] <= 0:
                continue
            algo = units[c] * ALGO_BYTES[c]
            gbs = algo / (ms[c] * 1e-3) / 1e9
            e = {"class": c, "kernel": what[c], "ms_per_frame": round(ms[c], 4), "launches_per_frame": nl[c], "units_per_frame": int(units[c]),
                 "algorithmic_bytes_per_unit": ALGO_BYTES[c], "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)}
            mc = (measured or {}).get("classes", {}).get(c)
            if mc:
                e["measured_dram_bytes_per_frame"] = int(mc["dram_bytes"])
                e["wasted_traffic"] = round(mc["dram_bytes"] / algo, 2) if algo else None   # measured / algorithmic
                e["lanes_per_instruction"] = mc["lanes_per_instruction"]
                e["issue_slot_utilisation_pct"] = mc["issue_slot_utilisation_pct"]
            classes.append(e)
        dom = max(classes, key=lambda e: e["ms_per_frame"])
        dom_launch_ms = dom["ms_per_frame"] / max(dom["launches_per_frame"], 1)
        line = {"metric": "Mrays/s", "value": round(float(value), 2), "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(float(mx[0] / K), 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl_name if args.workload != "C1" else WORKLOAD, "partition": (f"image tiles 64x64, tile % {world} == rank, " + ("sRGBA8 tiles stored into GPU 0's image over NVLink peer memory (one push kernel per rank, flag handshake)"
                                                                                              if isinstance(gather, multigpu.PeerTileExchange) else (gather_note or "") + "NCCL gather of sRGBA8 to GPU 0")) if world > 1 else "single device",
                           "l2": "256 MiB memset between timed steps (L2 flush)", **({"composition_check": "GPU 0's image == every rank's own tiles (checksums)"} if composed_ok else {}), "frame_ms": round(float(mx[0] / K), 4),
                           "rays_per_frame": int(tot[0] / K), "mpixel_samples_per_s": round(w * h * spp * K / (mx[0] * 1e-3) / 1e6, 1),
                           "scene_setup_ms_incl_bvh": round(setup_ms, 2)},
                "e2e": {"value": round(float(e2e), 2), "unit": "Mrays/s", "h2d_bytes_per_step": 472, "d2h_bytes_per_step": w * h * 4 + 24,
                        "ms_per_step": round(float(mx[1] / K * 1e3), 4), "warmup_steps": W},
                "gpu_launches": int(tot[2]),
                "clocks": clk,
                "roofline": {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac_of_hbm_peak"],
                             "traffic": (dom.get("measured_dram_bytes_per_frame") / max(dom["launches_per_frame"], 1)) if dom.get("measured_dram_bytes_per_frame") else None,
                             "peak_source": peak_src, "kernel": dom["kernel"], "kernel_class": dom["class"],
                             "kernel_ms_per_frame": dom["ms_per_frame"], "launches_per_frame": dom["launches_per_frame"],
                             "avg_launch_ms": round(dom_launch_ms, 4),
                             "algorithmic_bytes_per_launch": int(dom["units_per_frame"] * dom["algorithmic_bytes_per_unit"] / max(dom["launches_per_frame"], 1)),
                             "classes": classes,
                             "traffic_source": (f"profiles/r2_c1_kernel_classes.json (ncu, commit {measured.get('commit')})" if measured else None),
                             "note": "dominant class by time; `achieved` = algorithmic path-state bytes / time. No class of this frame is HBM-bound: "
                                     "the scene (34 quads) lives in L1/L2 and the kernels are limited by issue slots, divergence (lanes_per_instruction) "
                                     "and dependent-load latency -- the memory roofline is reported because the contract asks for it, the limiter is in "
                                     "the per-class ncu figures (DESIGN.md 4, profiles/)"}}
        if world == 1 and args.workload == "C1":
            try:
                line["roofline"]["fb_ops_4k"] = fb_ops_4k(peak)
            except Exception as e:
                line["roofline"]["fb_ops_4k"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline and args.workload == "C1":
            try:
                backend, kind = cpu_backend() if (w % 8 == 0 and h % 8 == 0) else ("oracle", "port")
                stride = calibrate_stride(w, h, 6.0, backend)
                r, s, threads = oracle_frames(w, h, 2, 1, stride, backend)
                line["cpu_baseline"] = {"value": round(r / s / 1e6, 3), "unit": "Mrays/s", "cores": threads, "kind": kind,
                                        "sample": f"2 frames, tiles with tileID % {stride} == 0 of the {w}x{h} frame; " + CPU_NOTE[kind],
                                        "caveat": "scalar build of the reference kernels, no ISPC SIMD, brute-force scene queries instead of Embree: "
                                                  "the reference's real CPU device would be several times faster; a reported baseline, not a target",
                                        "frame_ms_full": round(s / 2 * stride * 1e3, 1)}
            except Exception as e:  # the baseline must never take the bench line down
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        if hasattr(gather, "check"):
            gather.check()   # a device-side wait that timed out would have produced a wrong frame: fail loudly
            gather.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
