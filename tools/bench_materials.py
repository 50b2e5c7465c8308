#!/usr/bin/env python3
"""Frame time of the cornell box at 1024x1024, 1 spp, with the boxes' materials swapped: the plain kernels (obj) against the
extended-material instantiation (metal, plastic, glass + alloy, principled variants, carPaint, textures).  Device-timed with CUDA
events around ospRenderFrame on the render stream, min of the frames.
   gpurun -- 'python tools/bench_materials.py'"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from ospray_b200 import capi as osp, testing as T  # noqa: E402

osp.init()
stream = torch.cuda.Stream()
osp.set_render_stream(stream.cuda_stream)
names = sys.argv[1:] or ["cornell_pt", "metal_rough_pt", "plastic_smooth_pt", "glass_alloy_pt", "principled_plastic_pt", "principled_glass_pt",
                         "principled_coat_pt", "principled_sheen_pt", "principled_tex_pt", "carpaint_pt", "textured_pt"]
for name in names:
    build, kind, params = scenes.CASES[name][:3]
    with torch.cuda.stream(stream):
        world, keep = T.to_ospray(build())
        cam, _ = T.make_camera(world, 1024, 1024)
        ren = T.make_renderer(kind, pixelSamples=1)
        fb = osp.FrameBuffer(1024, 1024, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH)
        ms = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fb.render_frame(ren, cam, world).wait()
            e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
    print(json.dumps({"case": name, "frame_ms_min": round(min(ms[2:]), 3), "frame_ms_all": [round(m, 3) for m in ms],
                      "rays": [int(x) for x in fb.ray_counts()]}), flush=True)
