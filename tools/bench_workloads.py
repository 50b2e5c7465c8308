#!/usr/bin/env python3
"""Device-timed numbers for the BASELINE.json configs other than the headline one (bench.py covers C1):

  C1  cornell_box            pathtracer 1024x1024  1 spp
  C2  gravity_spheres_volume scivis     1920x1080  4 spp   (structured volume 128^3; --dim for 32/512)
  C2pt gravity_spheres_volume pathtracer 1920x1080 1 spp/frame (ospBenchmark's gravity_spheres_volume/dim_*/pathtracer)
  C3  ~10 M-triangle mesh    ao         3840x2160 16 spp   (94^3 boxes x 12 triangles; --boxes K for smaller)
  C4  random_spheres (thinGlass) + streamlines (round Catmull-Rom curves), pathtracer 1920x1080, 1 spp/frame
      accumulated progressively (--frames 64 gives the config's 64 spp)
  C5s cornell_box + mesh     pathtracer 3840x2160  1 spp/frame (the single-GPU leg of C5)

  python tools/bench_workloads.py C2 C3 [--frames 5] [--out gpurun_out/x/workloads.json]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_workloads.py C5s ...   (N GPUs: config C5 --
      replicated scene, tile % N == rank, tiles stored into GPU 0's image over NVLink peer memory; frame time = max over ranks)

Prints one JSON object per workload: frame ms (CUDA events on the render stream around ospRenderFrame's kernels,
L2 flushed between frames), rays per frame, Mrays/s, scene commit / BVH build ms. These are parity-test-sized
configs made into measurements; they are NOT the bench.py contract line."""
import argparse
import gc
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def merged(a, b):
    from ospray_b200 import testing as T
    return T.SceneSpec(instances=a.instances + b.instances, lights=a.lights)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["C1", "C2", "C3"])
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--boxes", type=int, default=94)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ospray_b200 import capi as osp, multigpu, testing as T
    rank, nranks, local = multigpu.init_process_group_from_env()
    torch.cuda.set_device(local)
    osp.init(cudaDevice=local, numSubdevices=nranks, subdeviceRank=rank)
    stream = torch.cuda.Stream()
    osp.set_render_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    results = []
    with torch.cuda.stream(stream):
        for wl in args.workloads:
            t0 = time.perf_counter()
            if wl == "C1":
                spec, ren_kind, w, h, spp, rp = T.cornell_box(), "pathtracer", 1024, 1024, 1, {}
            elif wl == "C2":
                spec, ren_kind, w, h, spp, rp = T.gravity_spheres_volume(args.dim), "scivis", 1920, 1080, 4, {}
            elif wl == "C2pt":  # the same volume under the pathtracer (delta tracking + Henyey-Greenstein), 1 spp/frame
                spec, ren_kind, w, h, spp, rp = T.gravity_spheres_volume(args.dim), "pathtracer", 1920, 1080, 1, {}
            elif wl == "C3":
                spec, ren_kind, w, h, spp, rp = T.box_mesh(args.boxes), "ao", 3840, 2160, 16, {}
            elif wl == "C4":
                spec, ren_kind, w, h, spp, rp = merged(T.random_spheres(renderer_type="pathtracer"), T.streamlines()), "pathtracer", 1920, 1080, 1, {}
            elif wl == "C5s":
                spec, ren_kind, w, h, spp, rp = merged(T.cornell_box(), T.box_mesh(args.boxes)), "pathtracer", 3840, 2160, 1, {}
            else:
                raise SystemExit(f"unknown workload {wl}")
            gen_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            world, keep = T.to_ospray(spec)
            torch.cuda.synchronize()
            commit_ms = (time.perf_counter() - t0) * 1e3
            build_ms = [g.build_ms() for g in keep if isinstance(g, osp.Group)]
            import ctypes as C
            from ospray_b200 import lib_path
            LC = C.CDLL(lib_path("libospcuda.so"))
            LC.ospcuda_bvh_peak_stack_need.restype = C.c_uint32
            stack_need = int(LC.ospcuda_bvh_peak_stack_need(1))  # deepest of the hierarchies built for this scene (device limit: 64)
            cam, pose = T.make_camera(world, w, h)
            ren = T.make_renderer(ren_kind, pixelSamples=spp, **rp)
            fb = osp.FrameBuffer(w, h, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH)
            conv_ptr, bpp = fb.device_pointer(osp.OSP_FB_COLOR)
            exch = multigpu.PeerTileExchange(w, h, bpp, conv_ptr) if nranks > 1 else None

            def frame():
                if exch:
                    exch.begin_frame()
                f = fb.render_frame(ren, cam, world)
                if exch:
                    exch.gather(conv_ptr)
                return f

            for _ in range(args.warmup):
                frame().wait()
                flush.zero_()
            torch.cuda.synchronize()
            if nranks > 1:
                dist.barrier()
            ms, rays = [], np.zeros(3, np.int64)
            for _ in range(args.frames):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                f = frame()
                e1.record(stream)
                f.wait()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
                rays += np.array(fb.ray_counts(), np.int64)
                flush.zero_()
            img = fb.map(osp.OSP_FB_COLOR)
            if nranks > 1:
                multigpu.verify_composition(img)  # GPU 0's image holds every rank's tiles (raises otherwise)
            if nranks > 1:  # frame time = max over ranks per frame, rays = sum over ranks
                t = torch.tensor(ms, dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.tolist()
                rt = torch.tensor(rays, dtype=torch.float64, device="cuda")
                dist.all_reduce(rt, op=dist.ReduceOp.SUM)
                rays = rt.cpu().numpy().astype(np.int64)
                exch.check()
                exch.close()
            tot_ms = float(np.sum(ms))
            r = {"workload": wl, "renderer": ren_kind, "size": [w, h], "spp": spp, "frames": args.frames,
                 "frame_ms": round(tot_ms / args.frames, 4), "frame_ms_min": round(float(np.min(ms)), 4), "frame_ms_all": [round(float(m), 3) for m in ms],
                 "rays_per_frame": int(rays.sum() // args.frames), "rays_split_primary_secondary_shadow": [int(x) // args.frames for x in rays],
                 "mrays_per_s": round(float(rays.sum()) / (tot_ms * 1e-3) / 1e6, 1),
                 "mpixel_samples_per_s": round(w * h * spp * args.frames / (tot_ms * 1e-3) / 1e6, 1),
                 "scene_generate_s": round(gen_s, 2), "scene_commit_ms": round(commit_ms, 1), "bvh_build_ms": [round(b, 2) for b in build_ms], "bvh_stack_need_last_built": stack_need,
                 "mean_rgb": [round(float(x), 2) for x in img[..., :3].reshape(-1, 3).mean(axis=0)]}
            if wl in ("C3", "C5s"):
                r["triangles"] = int(sum(m.geometry.index.shape[0] for i in spec.instances for m in i.group.models if m.geometry.index is not None))
            if wl == "C2":
                r["volume_dim"] = args.dim
            if nranks > 1:
                r["n_gpus"] = nranks
                r["partition"] = f"tile % {nranks} == rank, peer-memory tile exchange to GPU 0; frame_ms = max over ranks"
            if rank == 0:
                print(json.dumps(r), flush=True)
            results.append(r)
            del fb, ren, cam, world, keep, frame, exch
            gc.collect()                  # the previous workload's device memory is freed (cudaFree synchronises) before the
            torch.cuda.synchronize()      # next one is timed, not in the middle of its frames
    if nranks > 1:
        dist.barrier()
        dist.destroy_process_group()
    if args.out and rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
