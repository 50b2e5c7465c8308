#!/usr/bin/env python3
"""Generates tests/golden/oracle_renders.npz: float colour / depth / id buffers of the small test cases in
tests/scenes.py rendered by the CPU oracle. They pin the oracle against regressions (CPU suite) and are a
second, committed reference for the GPU parity tests. The reference ships no images (its baselines are md5
stubs), so these are NOT upstream baselines -- they are outputs of oracle/oracle.cpp.
Run:  make -C oracle && python tests/golden/gen_render_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import scenes  # noqa: E402

out = {}
for name in scenes.CASES:
    if name in scenes.ORACLE_UNSUPPORTED:
        continue
    fb = scenes.oracle_render(name, threads=1)
    out[name + "/color"] = fb.buf["color"].astype(np.float32)
    out[name + "/depth"] = fb.buf["depth"].astype(np.float32)
    out[name + "/primID"] = fb.buf["primID"]
    out[name + "/rays"] = fb.counters.copy()
    print(name, fb.buf["color"].reshape(-1, 4).mean(axis=0), fb.counters)
np.savez_compressed(os.path.join(HERE, "oracle_renders.npz"), **out)
