#!/usr/bin/env python3
"""Wall-clock breakdown of one bench.py e2e step (C1) through the public API: commit / ospRenderFrame / ospWait / map."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ospray_b200 import capi as osp, testing as T
osp.init(cudaDevice=0)
stream = torch.cuda.Stream()
osp.set_render_stream(stream.cuda_stream)
w = h = 1024
with torch.cuda.stream(stream):
    world, keep = T.to_ospray(T.cornell_box())
    cam, pose = T.make_camera(world, w, h)
    ren = T.make_renderer("pathtracer", pixelSamples=1)
    fb = osp.FrameBuffer(w, h, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH)
    for _ in range(5):
        fb.render_frame(ren, cam, world).wait()
    torch.cuda.synchronize()
    acc = np.zeros(5)
    N = 30
    for i in range(N):
        t0 = time.perf_counter()
        cam.set("position", pose["position"]).commit()
        t1 = time.perf_counter()
        f = fb.render_frame(ren, cam, world)
        t2 = time.perf_counter()
        f.wait()
        t3 = time.perf_counter()
        img = fb.map(osp.OSP_FB_COLOR)
        t4 = time.perf_counter()
        rc = fb.ray_counts()
        t5 = time.perf_counter()
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
        dev = f.stage_ms()
    import ctypes as C
    L = osp.lib()
    ta = tb = tc = 0.0
    for i in range(N):
        f = fb.render_frame(ren, cam, world); f.wait()
        t0 = time.perf_counter()
        p = L.ospMapFrameBuffer(fb.h, osp.OSP_FB_COLOR)
        t1 = time.perf_counter()
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (h, w, 4))
        t2 = time.perf_counter()
        b = a.copy()
        t3 = time.perf_counter()
        L.ospUnmapFrameBuffer(p, fb.h)
        ta += t1 - t0; tb += t2 - t1; tc += t3 - t2
    print("map split ms: ospMapFrameBuffer %.3f  as_array %.3f  copy %.3f" % (ta / N * 1e3, tb / N * 1e3, tc / N * 1e3))
    print("ms per step: commit %.3f  renderFrame(call) %.3f  wait %.3f  map %.3f  ray_counts %.3f  total %.3f | device stages of last frame: %s"
          % (*(acc / N * 1e3), acc.sum() / N * 1e3, dev))
