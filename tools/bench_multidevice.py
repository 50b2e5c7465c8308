#!/usr/bin/env python3
"""Frame time of the in-process multidevice (device type "multidevice") against the number of sub-devices, one process:
   python tools/bench_multidevice.py 1 2 4 8 [--workload C1|C5s] [--frames K]
Wall clock around ospRenderFrame + ospWait (the tile pull is inside the frame), after warm-up; frames accumulate."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("subdevices", nargs="+", type=int)
    ap.add_argument("--workload", default="C1")
    ap.add_argument("--frames", type=int, default=30)
    args = ap.parse_args()
    import bench
    from ospray_b200 import capi as osp, testing as T
    osp.init()
    name, ren_kind, w, h, spp = bench.WORKLOADS["C5" if args.workload == "C5s" else args.workload]
    for n in args.subdevices:
        dev = osp.new_device("multidevice", numSubdevices=n)
        osp.set_current(dev)
        world, keep = T.to_ospray(bench.workload_spec("C5" if args.workload == "C5s" else args.workload))
        cam, _ = T.make_camera(world, w, h)
        ren = T.make_renderer(ren_kind, pixelSamples=spp)
        fb = osp.FrameBuffer(w, h, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH)
        for _ in range(5):
            fb.render_frame(ren, cam, world).wait()
        t0 = time.perf_counter()
        rays = 0
        for _ in range(args.frames):
            fb.render_frame(ren, cam, world).wait()
            rays += sum(fb.ray_counts())
        dt = time.perf_counter() - t0
        img = fb.map(osp.OSP_FB_COLOR)
        print(json.dumps({"workload": name, "device": "multidevice", "subdevices": n, "frame_ms": round(dt / args.frames * 1e3, 4),
                          "mrays_per_s": round(rays / dt / 1e6, 1), "mean_rgb": [round(float(x), 3) for x in img[..., :3].mean(axis=(0, 1))]}), flush=True)
        del fb, world, keep, cam, ren


if __name__ == "__main__":
    main()
