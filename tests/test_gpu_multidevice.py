"""The in-process multidevice (device type "multidevice", csrc/device/MultiDevice.cpp; reference: modules/multiDevice): N
cuda sub-devices in one process behind the unchanged OSPRay API. Every object of the application is replicated on every
sub-device, each renders the tiles with tile % N == its index, GPU 0 pulls the other GPUs' tiles inside ospRenderFrame and
ospMapFrameBuffer returns the composed image. On a box with fewer GPUs than sub-devices several sub-devices share a GPU
(same code path apart from the peer mapping), so this runs on the one-GPU test box too; `gpurun --gpus 2` exercises NVLink."""
import os
import subprocess

import numpy as np
import pytest

from ospray_b200 import testing as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "native", "_bin", "ospTutorial")


def _render(osp, spec, kind, w, h, frames, channels, fmt, **params):
    world, keep = T.to_ospray(spec)
    cam, _ = T.make_camera(world, w, h)
    ren = T.make_renderer(kind, **params)
    fb = osp.FrameBuffer(w, h, fmt, channels)
    for _ in range(frames):
        fb.render_frame(ren, cam, world).wait()
    return fb, (world, keep, cam, ren)


def _materials_and_hdri():
    """principled boxes with texture maps (device-side map tables per material) under an hdri light (device-side distribution per
    light) plus a group light in a rotated instance: everything a sub-device has to build for itself"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    spec = scenes.principled_tex_cornell(False)
    spec.lights += scenes.hdri_lit().lights
    spec.instances.append(scenes.hdri_lit(True).instances[-1])
    return spec


@pytest.mark.parametrize("case", ["cornell_pt", "spheres_ao", "gravity_scivis", "materials_hdri_pt"])
@pytest.mark.parametrize("n", [2, 3])
def test_multidevice_frame_equals_the_single_device_frame(osp, case, n):
    spec, kind, params = {"cornell_pt": (T.cornell_box(), "pathtracer", {"pixelSamples": 2}),
                          "materials_hdri_pt": (_materials_and_hdri(), "pathtracer", {"pixelSamples": 1}),
                          "spheres_ao": (T.random_spheres(), "ao", {"pixelSamples": 2}),
                          "gravity_scivis": (T.gravity_spheres_volume(32), "scivis", {"pixelSamples": 1})}[case]
    w, h = 320, 200   # 5 x 4 tiles, the last row and column clipped
    ch = osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH | osp.OSP_FB_ID_PRIMITIVE | osp.OSP_FB_NORMAL
    single, keep = _render(osp, spec, kind, w, h, 2, ch, osp.OSP_FB_SRGBA, **params)
    want = {c: single.map(c) for c in (osp.OSP_FB_COLOR, osp.OSP_FB_DEPTH, osp.OSP_FB_ID_PRIMITIVE, osp.OSP_FB_NORMAL)}
    want_rays = single.ray_counts()
    want_bounds = keep[0].bounds()
    main = osp._device
    dev = osp.new_device("multidevice", numSubdevices=n)
    osp.set_current(dev)
    try:
        multi, keep2 = _render(osp, spec, kind, w, h, 2, ch, osp.OSP_FB_SRGBA, **params)
        for c, wv in want.items():
            np.testing.assert_array_equal(multi.map(c), wv)      # bit-identical: same kernels, same streams, tiles only moved
        assert multi.ray_counts() == want_rays                   # summed over the sub-devices
        lo, hi = keep2[0].bounds()
        np.testing.assert_array_equal(lo, want_bounds[0])
        np.testing.assert_array_equal(hi, want_bounds[1])
        del multi, keep2
    finally:
        osp.set_current(main)


def test_multidevice_pick_and_float_colour(osp):
    """OSP_FB_RGBA32F display format through the tile pull (16 B pixels) and ospPick handing back the application's handles"""
    spec = T.cornell_box()
    w, h = 192, 128
    single, keep = _render(osp, spec, "pathtracer", w, h, 1, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM, osp.OSP_FB_RGBA32F)
    want = single.map(osp.OSP_FB_COLOR)
    main = osp._device
    dev = osp.new_device("multidevice", numSubdevices=4)
    osp.set_current(dev)
    try:
        multi, keep2 = _render(osp, spec, "pathtracer", w, h, 1, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM, osp.OSP_FB_RGBA32F)
        np.testing.assert_array_equal(multi.map(osp.OSP_FB_COLOR), want)
        world, _, cam, ren = keep2
        single_pick = None
        hit, pos, inst, model, prim = osp.pick(multi, ren, cam, world, 0.5, 0.5)
        assert hit and inst and model
        osp.set_current(main)
        single_pick = osp.pick(single, keep[3], keep[2], keep[0], 0.5, 0.5)
        osp.set_current(dev)
        assert single_pick[0] and prim == single_pick[4]
        np.testing.assert_array_equal(pos, single_pick[1])
        del multi, keep2
    finally:
        osp.set_current(main)


def _read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = map(int, f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h * 3), np.uint8).reshape(h, w, 3)


@pytest.mark.skipif(not os.path.exists(BIN), reason="tests/native/_bin/ospTutorial is built by __graft_entry__.build() where /root/reference is mounted")
def test_unchanged_reference_tutorial_on_the_multidevice(tmp_path):
    """the reference's ospTutorial.c binary, unchanged, on 4 sub-devices: only the command line names another device"""
    outs = {}
    for name, args, env in (("cuda", ["--osp:load-modules=cuda", "--osp:device=cuda"], {}),
                            ("multi", ["--osp:load-modules=cuda", "--osp:device=multidevice"], {"OSPRAY_NUM_SUBDEVICES": "4"})):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([BIN] + args, cwd=d, env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
        outs[name] = (_read_ppm(d / "firstFrame.ppm"), _read_ppm(d / "accumulatedFrame.ppm"), r.stdout)
    np.testing.assert_array_equal(outs["multi"][0], outs["cuda"][0])
    np.testing.assert_array_equal(outs["multi"][1], outs["cuda"][1])
    assert "prim: " in outs["multi"][2] and "(nil)" not in outs["multi"][2].split("ospPick")[1].splitlines()[0]
