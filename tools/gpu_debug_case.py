#!/usr/bin/env python3
"""Render test cases on the device and with the checker; dump both to gpurun_out/<tag>/<case>.npz for inspection.
   gpurun -- 'python tools/gpu_debug_case.py r2dbg textured_pt cutout_pt'"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from ospray_b200 import capi as osp  # noqa: E402

tag, names = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "gpurun_out", tag)
os.makedirs(out, exist_ok=True)
osp.init()
import json  # noqa: E402
extra = json.loads(os.environ.get("DBG_EXTRA", "{}"))  # renderer parameters laid over the case's own
for name in names:
    if extra:
        c = scenes.CASES[name]
        scenes.CASES[name] = (c[0], c[1], {**c[2], **extra}) + tuple(c[3:])
    backend = "ref" if name in scenes.ORACLE_UNSUPPORTED else "oracle"
    pf = "point" if os.environ.get("DBG_POINT") else None
    w, h = (int(os.environ.get("DBG_W", scenes.W)), int(os.environ.get("DBG_H", scenes.H)))
    want = scenes.oracle_render(name, w, h, pixel_filter=pf, backend=backend)
    fb, keep = scenes.device_render(osp, name, w, h, pixel_filter=pf)
    d = {"dev_color": fb.map(osp.OSP_FB_COLOR), "ref_color": want.buf["color"], "dev_prim": fb.map(osp.OSP_FB_ID_PRIMITIVE),
         "ref_prim": want.buf["primID"], "dev_obj": fb.map(osp.OSP_FB_ID_OBJECT), "ref_obj": want.buf["objID"],
         "dev_depth": fb.map(osp.OSP_FB_DEPTH), "ref_depth": want.buf["depth"], "dev_albedo": fb.map(osp.OSP_FB_ALBEDO),
         "ref_albedo": want.buf["albedo"], "dev_normal": fb.map(osp.OSP_FB_NORMAL), "ref_normal": want.buf["normal"],
         "dev_rays": np.array(fb.ray_counts()), "ref_rays": want.counters}
    np.savez_compressed(os.path.join(out, name + ".npz"), **d)
    err = np.abs(d["dev_color"] - d["ref_color"]).max(axis=-1)
    print(name, "bad frac", (err > 1e-4 + 2e-3 * np.abs(d["ref_color"]).max(axis=-1)).mean(), d["dev_rays"], d["ref_rays"])
