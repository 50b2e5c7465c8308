"""CPU suite for the oracle itself: known answers for rays, framebuffer ops and pixel filter, the committed
golden renders, and size-independent properties (tile partition union, accumulation law)."""
import ctypes as C
import os

import numpy as np
import pytest

import scenes
from oracle import pyoracle as O
from ospray_b200 import testing as T

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_renders.npz"))


def _rays(orgs, dirs, t0=0.0, t1=np.inf):
    r = np.zeros(len(orgs), O.RAYHIT_DTYPE)
    r["org"], r["dir"], r["t0"], r["t"] = orgs, dirs, t0, t1
    r["primID"] = r["geomID"] = r["instID"] = O.INVALID_ID
    return r


def test_cornell_known_hits():
    # a camera-height ray above the boxes hits the back wall (quad 2, z = +1) at t = 1 + 2.4641016
    sc = O.Scene(T.cornell_box())
    r = O.trace(sc, _rays([[-0.5, 0.5, -2.4641016]], [[0, 0, 1]]))
    assert r["primID"][0] == 2 and r["geomID"][0] == 0 and r["instID"][0] == 0
    assert abs(r["t"][0] - 3.4641016) < 1e-6
    # straight down from the middle: floor (quad 0) at y = -1; Ng points along +-y
    r = O.trace(sc, _rays([[0.5, 0.5, -0.8]], [[0, -1, 0]]))
    assert r["primID"][0] == 0 and abs(r["t"][0] - 1.5) < 1e-6 and abs(abs(r["Ng"][0][1]) - np.linalg.norm(r["Ng"][0])) < 1e-6
    # occlusion query: reference convention ray.t = -inf when occluded (World.ih:334)
    r = O.trace(sc, _rays([[0, 0, -2.4641016], [0, 0, -5]], [[0, 0, 1], [0, 0, -1]]), occlusion=True)
    assert r["t"][0] == -np.inf and r["t"][1] == np.inf


def test_sphere_and_box_known_hits():
    spec = T.SceneSpec(instances=[T.InstanceSpec(T.GroupSpec(models=[
        T.ModelSpec(T.GeometrySpec("sphere", sphere_position=np.array([[0, 0, 5]], np.float32), sphere_radius=np.array([1.5], np.float32))),
        T.ModelSpec(T.GeometrySpec("box", boxes=np.array([[2, -1, 3, 4, 1, 6]], np.float32)))]))])
    sc = O.Scene(spec)
    r = O.trace(sc, _rays([[0, 0, 0], [3, 0, 0], [3, 0, 4], [0, 0, 5]], [[0, 0, 1], [0, 0, 1], [0, 0, 1], [1, 0, 0]]))
    assert r["geomID"].tolist() == [0, 1, 1, 0]
    np.testing.assert_allclose(r["t"], [3.5, 3.0, 2.0, 1.5], rtol=1e-6)
    assert r["Ng"][1].tolist() == [0, 0, -1]      # entry face (Intersect.ih:206-211)
    assert r["Ng"][2].tolist() == [0, 0, 1]       # origin inside: exit hit (FilterIntersect.ih:108-116)
    assert r["Ng"][3][0] > 0                      # inside the sphere: back hit, Ng = P - centre


def test_quad_second_triangle_uv_remap():
    # quad (v0,v1,v2,v3): hits in triangle (v2,v3,v1) report (1-u, 1-v) (Mesh.ispc:325-345)
    q = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    spec = T.SceneSpec(instances=[T.InstanceSpec(T.GroupSpec(models=[T.ModelSpec(T.GeometrySpec("mesh", vertex=q, quad_soup=True))]))])
    sc = O.Scene(spec)
    r = O.trace(sc, _rays([[0.25, 0.25, -1], [0.75, 0.75, -1]], [[0, 0, 1], [0, 0, 1]]))
    np.testing.assert_allclose(np.stack([r["u"], r["v"]], 1), [[0.25, 0.25], [0.75, 0.75]], atol=1e-6)


def test_color_conversion_known_answers():
    # ColorConversion.ispc:11-36 with rkcommon linear_to_srgba8 / cvt_uint32 (IEC 61966-2-1)
    v = np.array([[0.0, 0.5, 1.0, 0.5], [0.0031308, 0.2, 2.0, 1.5], [-1.0, 0.01, 0.735, 0.25]], np.float32)
    srgb = O.color_convert(v, 2).view(np.uint8).reshape(-1, 4)
    lin = np.clip(v[:, :3].astype(np.float64), 0, None)
    want = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055)
    np.testing.assert_array_equal(srgb[:, :3], np.rint(255 * np.clip(want, 0, 1)).astype(np.uint8))
    np.testing.assert_array_equal(srgb[:, 3], [128, 255, 64])
    rgba8 = O.color_convert(v, 1).view(np.uint8).reshape(-1, 4)
    np.testing.assert_array_equal(rgba8[0], [0, 128, 255, 128])
    f = O.color_convert(v, 3)
    np.testing.assert_array_equal(f[2], np.array([0.0, 0.01, 0.735, 0.25], np.float32))
    assert f[1, 3] == 1.0 and f[1, 2] == 2.0        # colour only clamped below, alpha to [0,1]


def test_tonemapper_matches_float64_formula():
    tm = O.default_tonemapper()
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 4, (256, 4)).astype(np.float32)
    got = O.tonemap(x, tm)
    Min = np.array([[0.5972782409, 0.3545713181, 0.0482176639], [0.0760130499, 0.9083220973, 0.0156579968], [0.0284085382, 0.1338243154, 0.8375684636]])
    Mout = np.array([[1.6047539945, -0.5310794927, -0.0736720338], [-0.1020831870, 1.1081322801, -0.0060518756], [-0.0032670420, -0.0727552477, 1.0760219533]])
    c = x[:, :3].astype(np.float64) @ Min.T
    c = c ** tm.a / (c ** (tm.a * tm.d) * tm.b + tm.c)
    want = np.clip(c @ Mout.T, 0, 1)
    np.testing.assert_allclose(got[:, :3], want, atol=2e-5)
    np.testing.assert_array_equal(got[:, 3], x[:, 3])
    # defaults reproduce the published ACES-fit: mid grey stays mid grey, hdrMax maps to 1
    g = O.tonemap(np.array([[0.18, 0.18, 0.18, 1]], np.float32), O.default_tonemapper(acesColor=False))
    np.testing.assert_allclose(g[0, :3], 0.18, atol=1e-4)
    g = O.tonemap(np.array([[11.0785] * 3 + [1]], np.float32), O.default_tonemapper(acesColor=False))
    np.testing.assert_allclose(g[0, :3], 1.0, atol=1e-4)


def test_accumulation_law():
    # FRAMEBUFFER_ACCUMULATE_VALUE (FrameBuffer.ih:17-22): first frame overwrites, then running mean
    rng = np.random.default_rng(0)
    fb = O.HostFrameBuffer(8, 4, ("color", "depth"), variance=True)
    fb.buf["color"][:] = 123.0  # stale data must not leak into frame 0
    samples = rng.uniform(0, 1, (5, 4, 8, 4)).astype(np.float32)
    depths = rng.uniform(1, 9, (5, 4, 8)).astype(np.float32)
    for i in range(5):
        fb.begin_frame()
        O.lib().oracle_fb_accumulate(C.byref(fb.sh), samples[i].ctypes.data, depths[i].ctypes.data)
    np.testing.assert_allclose(fb.buf["color"], samples.mean(axis=0), rtol=2e-6)
    np.testing.assert_array_equal(fb.buf["depth"], depths.min(axis=0))   # min-|z| sample wins (LocalFB.ispc:40-46)
    # the variance buffer accumulates on odd frames only, with scale rcp(frameID/2+1)
    np.testing.assert_allclose(fb.buf["variance"], samples[[1, 3]].mean(axis=0), rtol=2e-6)


def test_variance_known_answer():
    c = np.zeros((8, 16, 4), np.float32)
    v = np.zeros((8, 16, 4), np.float32)
    c[..., :3], c[..., 3] = 0.5, 1.0
    v[..., :3], v[..., 3] = 0.25, 1.0
    err = O.variance(c, v)
    # per pixel: sum|c-v| = 0.75, den2 = 1.5 -> 0.75/sqrt(1.5); task mean * 64 (Variance.ispc:40-58)
    np.testing.assert_allclose(err, [0.75 / np.sqrt(1.5) * 64] * 2, rtol=1e-6)


def test_gauss_pixel_filter_lut():
    r = O.RendererParams("pathtracer")
    cx, cy = r.cdfx.reshape(15, 15), r.cdfy
    assert (np.diff(cy) > 0).all() and (np.diff(cx, axis=1) >= 0).all()
    assert cy[-1] > 1.0 and (cx[:, -1] > 1.0).all()           # nextAfter1 terminator (Distribution1D.ispc:27-37)
    pdf_y = np.diff(np.concatenate([[0], np.minimum(cy, 1)]))
    np.testing.assert_allclose(pdf_y, pdf_y[::-1], rtol=1e-4)  # symmetric Gaussian
    assert pdf_y.argmax() == 7


@pytest.mark.parametrize("name", [n for n in scenes.CASES if n not in scenes.ORACLE_UNSUPPORTED])
def test_golden_renders(name):
    fb = scenes.oracle_render(name)
    np.testing.assert_allclose(fb.buf["color"], GOLD[name + "/color"], rtol=1e-4, atol=1e-5)
    d, g = fb.buf["depth"], GOLD[name + "/depth"]
    assert (np.isinf(d) == np.isinf(g)).all()
    np.testing.assert_allclose(d[np.isfinite(d)], g[np.isfinite(g)], rtol=1e-5)
    assert (fb.buf["primID"] != GOLD[name + "/primID"]).mean() < 0.002
    np.testing.assert_array_equal(fb.counters[0], GOLD[name + "/rays"][0])


@pytest.mark.parametrize("name", ["cornell_pt", "spheres_ao", "gravity_scivis"])
def test_tile_partition_union_equals_single_device(name):
    # MultiDevice.cpp:403-416: the ranks' tile sets are disjoint and their union is the frame, bit for bit
    w, h = 150, 70  # 3 x 2 tiles
    full = scenes.oracle_render(name, w, h)
    parts = [scenes.oracle_render(name, w, h, num_ranks=3, rank=r) for r in range(3)]
    comp = np.zeros_like(full.buf["color"])
    for r, p in enumerate(parts):
        O.lib().oracle_fb_gather_tiles(comp.ctypes.data, p.buf["color"].ctypes.data, w, h, 16, 3, r)
    np.testing.assert_array_equal(comp, full.buf["color"])
    assert sum(int(p.counters.sum()) for p in parts) == int(full.counters.sum())


def test_converged_cornell_is_stable():
    # the 1-spp estimator is unbiased: the mean radiance of 16 vs 64 accumulated frames agrees within noise
    from oracle import pyoracle as O2
    sc = O2.Scene(T.cornell_box())
    pose = T.arcball_camera(*sc.bounds)
    cam = O2.make_camera(pose["position"], aspect=1.0)
    ren = O2.RendererParams("pathtracer")
    fb = O2.HostFrameBuffer(32, 32, ("color",))
    means = {}
    for i in range(64):
        O2.render_frame(sc, cam, ren, fb)
        if i + 1 in (16, 64):
            means[i + 1] = fb.buf["color"][..., :3].mean()
    assert abs(means[16] - means[64]) / means[64] < 0.05
    assert 0.2 < means[64] < 1.0
