#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE's own kernels (oracle/_ref/libospref.so, built by oracle/build_ref.py from
/root/reference/modules/cpu/**/*.ispc in their scalar C++ mode).  Run where /root/reference is present:

    python oracle/build_ref.py && python tests/golden/gen_ref_golden.py

Writes tests/golden/ref_renders.npz (every case of tests/scenes.py the scalar reference kernels define, 48x40: colour,
depth, ids) and tests/golden/ref_kats.npz (sampler streams, Gaussian pixel-filter LUT, frame ops and accumulation on
seeded inputs) and tests/golden/ref_albedo_tables.npy (the microfacet albedo tables).  The committed files let the oracle (and through it the CUDA path) be checked against the reference on
machines where neither /root/reference nor oracle/_ref exists."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import scenes  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

KAT_SEEDS = [0x49c5f37e ^ 12345, 7, 0x49c5f37e ^ (1024 * 511 + 77)]


def kat_inputs():
    rng = np.random.default_rng(2024)
    img = (rng.random((40, 48, 4), np.float32) * 1.3).astype(np.float32)
    var = rng.random((40, 48, 4), np.float32).astype(np.float32)
    samples = rng.random((5, 40, 48, 4), np.float32).astype(np.float32)
    depths = (rng.random((5, 40, 48), np.float32) * 10).astype(np.float32)
    return img, var, samples, depths


def kats(backend):
    out = {}
    for kind in range(5):
        for si, seed in enumerate(KAT_SEEDS):
            out[f"sampler/{kind}/{si}"] = O.sampler_kat(kind, seed, 0, 512, backend).view(np.uint32)
    cx, cy = O.gauss_pixelfilter_lut(3.0, 15, backend)
    out["lut/cdf_x"], out["lut/cdf_y"] = cx, cy
    img, var, samples, depths = kat_inputs()
    for fmt in (1, 2, 3):
        out[f"color_convert/{fmt}"] = O.color_convert(img, fmt, backend)
    out["tonemap"] = O.tonemap(img * 3, O.default_tonemapper(), backend)
    out["variance"] = O.variance(img, var, 8, backend)
    fb = O.HostFrameBuffer(48, 40, ("color", "depth"), variance=True)
    for i in range(5):
        O.accumulate(fb, samples[i], depths[i], backend)
    out["accumulate/color"], out["accumulate/depth"], out["accumulate/variance"] = fb.buf["color"], fb.buf["depth"], fb.buf["variance"]
    return out


def main():
    renders = {}
    for name in scenes.CASES:
        if name in scenes.REF_UNSUPPORTED:
            continue
        fb = scenes.oracle_render(name, backend="ref")
        for ch in ("color", "depth", "primID", "objID", "instID", "normal", "albedo"):
            renders[f"{name}/{ch}"] = fb.buf[ch]
    np.savez_compressed(os.path.join(HERE, "ref_renders.npz"), **renders)
    np.savez_compressed(os.path.join(HERE, "ref_kats.npz"), **kats("ref"))
    np.save(os.path.join(HERE, "ref_albedo_tables.npy"), O.ref_albedo_tables())  # precomputeMicrofacetAlbedoTables
    print("wrote ref_renders.npz (%d arrays), ref_kats.npz" % len(renders))


if __name__ == "__main__":
    main()
