"""Edge cases the reference's own scenes cover (ospray_testing `empty`, single-primitive groups, degenerate primitives):
an empty world renders the background under every renderer, a one-triangle group takes the hierarchy-free path, a
zero-area triangle is never hit.  CPU: oracle against the reference's own kernels; GPU: the device against the oracle."""
import os

import numpy as np
import pytest

import scenes
from ospray_b200 import testing as T

BG = (0.1, 0.2, 0.3, 1.0)
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libospref.so")),
                               reason="oracle/_ref not built (needs /root/reference)")


def _one_triangle(degenerate=False):
    v = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32)
    if degenerate:
        v[2] = v[1]  # zero area
    big = np.array([[-2, -2, -1], [2, -2, -1], [2, 2, -1], [-2, 2, -1]], np.float32)  # something behind it to see through to
    tri = T.GeometrySpec("mesh", vertex=v, index=np.array([[0, 1, 2]], np.uint32))
    quad = T.GeometrySpec("mesh", vertex=big, index=np.array([[0, 1, 2, 3]], np.uint32))
    g = T.GroupSpec(models=[T.ModelSpec(tri, T.MaterialSpec()), T.ModelSpec(quad, T.MaterialSpec())])
    return T.SceneSpec(instances=[T.InstanceSpec(g)], lights=[T.LightSpec("ambient", {})])


POSE = {"position": (0.0, 0.0, 3.0), "direction": (0.0, 0.0, -1.0), "up": (0.0, 1.0, 0.0)}


def _oracle(spec, kind, backend="oracle", **params):
    from oracle import pyoracle as O
    sc = O.Scene(spec)
    cam = O.make_camera(POSE["position"], POSE["direction"], POSE["up"], aspect=np.float32(48) / np.float32(40))
    fb = O.HostFrameBuffer(48, 40, ("color", "depth", "primID", "objID", "instID"))
    O.render_frame(sc, cam, O.RendererParams(kind, pixelFilter="point", backgroundColor=BG, **params), fb, 1, backend=backend)
    return fb


@needs_ref
@pytest.mark.parametrize("kind", ["pathtracer", "scivis", "ao", "debug"])
def test_empty_world_is_background_in_checker(kind):
    for backend in ("oracle", "ref"):
        fb = _oracle(T.SceneSpec(instances=[], lights=[]), kind, backend)
        np.testing.assert_array_equal(fb.buf["color"].reshape(-1, 4), np.tile(np.float32(BG), (48 * 40, 1)), err_msg=backend)
        assert np.isinf(fb.buf["depth"]).all()


@needs_ref
@pytest.mark.parametrize("degenerate", [False, True])
def test_single_and_degenerate_triangle_in_checker(degenerate):
    a = _oracle(_one_triangle(degenerate), "scivis"), _oracle(_one_triangle(degenerate), "scivis", "ref")
    np.testing.assert_array_equal(a[0].buf["objID"], a[1].buf["objID"])
    np.testing.assert_array_equal(a[0].buf["depth"], a[1].buf["depth"])
    hit_front = (a[0].buf["objID"] == 0).mean()
    assert (hit_front == 0.0) if degenerate else (0.05 < hit_front < 0.6)   # a zero-area triangle is never hit
    assert (a[0].buf["objID"] == 1).mean() > 0.3                            # the quad behind it is


def _device(osp, spec, kind, **params):
    world, keep = T.to_ospray(spec)
    cam = osp.Camera("perspective")
    cam.set("aspect", float(np.float32(48) / np.float32(40)))
    cam.set("position", POSE["position"]).set("direction", POSE["direction"]).set("up", POSE["up"]).commit()
    ren = T.make_renderer(kind, pixelFilter=osp.Typed(osp.OSP_UINT, 0), backgroundColor=BG, **params)
    fb = osp.FrameBuffer(48, 40, osp.OSP_FB_RGBA32F, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH | osp.OSP_FB_ID_OBJECT)
    fb.render_frame(ren, cam, world).wait()
    return fb, (world, keep, cam, ren)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["pathtracer", "scivis", "ao", "debug"])
def test_empty_world_is_background_on_device(osp, kind):
    fb, keep = _device(osp, T.SceneSpec(instances=[], lights=[]), kind)
    np.testing.assert_array_equal(fb.map(osp.OSP_FB_COLOR).reshape(-1, 4), np.tile(np.float32(BG), (48 * 40, 1)))
    assert np.isinf(fb.map(osp.OSP_FB_DEPTH)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("degenerate", [False, True])
def test_single_and_degenerate_triangle_on_device(osp, degenerate):
    want = _oracle(_one_triangle(degenerate), "scivis")
    fb, keep = _device(osp, _one_triangle(degenerate), "scivis")
    got = fb.map(osp.OSP_FB_ID_OBJECT)
    assert (got != want.buf["objID"]).mean() <= 0.005     # edge pixels only (ray set-up rounds differently in the last bit)
    d, dw = fb.map(osp.OSP_FB_DEPTH), want.buf["depth"]
    same = got == want.buf["objID"]
    assert (np.isinf(d[same]) == np.isinf(dw[same])).all()
    fin = same & np.isfinite(dw)
    np.testing.assert_allclose(d[fin], dw[fin], rtol=2e-5)
