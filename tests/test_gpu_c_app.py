"""A compiled C application on the cuda device: the reference's apps/ospTutorial/ospTutorial.c, built UNCHANGED against this
repo's include/ and libospray.so by __graft_entry__.build() (the source stays in /root/reference; the binary travels to the
GPU box), started the way an OSPRay user would: --osp:load-modules=cuda --osp:device=cuda. It writes firstFrame.ppm and
accumulatedFrame.ppm (1024x768, pathtracer, two triangles with vertex colours under an ambient light); the images are
compared with the same scene rendered through the Python binding of the same C API."""
import os
import subprocess

import numpy as np
import pytest

from ospray_b200 import testing as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "native", "_bin", "ospTutorial")


def _read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = map(int, f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(w * h * 3), np.uint8).reshape(h, w, 3)


def _tutorial_scene(osp):
    """ospTutorial.c:54-196 through the Python binding"""
    vertex = np.array([[-1, -1, 3], [-1, 1, 3], [1, -1, 3], [0.1, 0.1, 0.3]], np.float32)
    color = np.array([[0.9, 0.5, 0.5, 1], [0.8, 0.8, 0.8, 1], [0.8, 0.8, 0.8, 1], [0.5, 0.9, 0.5, 1]], np.float32)
    index = np.array([[0, 1, 2], [1, 2, 3]], np.uint32)
    geom = T.GeometrySpec("mesh", vertex=vertex, vertex_color=color, index=index)
    spec = T.SceneSpec(instances=[T.InstanceSpec(T.GroupSpec(models=[T.ModelSpec(geom, T.MaterialSpec())]))],
                       lights=[T.LightSpec("ambient", {})])
    world, keep = T.to_ospray(spec)
    cam = osp.Camera("perspective")
    cam.set("aspect", float(np.float32(1024) / np.float32(768)))
    cam.set("position", (0.0, 0.0, 0.0)).set("direction", (0.1, 0.0, 1.0)).set("up", (0.0, 1.0, 0.0)).commit()
    ren = osp.Renderer("pathtracer")
    ren.set("backgroundColor", 1.0).commit()
    return world, keep, cam, ren


@pytest.mark.skipif(not os.path.exists(BIN), reason="tests/native/_bin/ospTutorial is built by __graft_entry__.build() where /root/reference is mounted")
def test_unchanged_reference_tutorial_runs_on_the_cuda_device(osp, tmp_path):
    env = dict(os.environ)
    r = subprocess.run([BIN, "--osp:load-modules=cuda", "--osp:device=cuda"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "done" in r.stdout.lower()
    first = _read_ppm(tmp_path / "firstFrame.ppm")
    accum = _read_ppm(tmp_path / "accumulatedFrame.ppm")
    assert first.shape == accum.shape == (768, 1024, 3)
    # the two triangles cover the middle of the image, the background is white (backgroundColor 1.0)
    assert (first[0, 0] == 255).all() and first[384, 400].max() < 250
    # the same frames through the Python binding of the same API: SRGBA bytes of frame 1 and of 11 accumulated frames
    world, keep, cam, ren = _tutorial_scene(osp)
    fb = osp.FrameBuffer(1024, 768, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM)
    fb.render_frame(ren, cam, world).wait()
    got_first = fb.map(osp.OSP_FB_COLOR)[::-1, :, :3]      # writePPM flips: row 0 of the file is the top row
    for _ in range(10):
        fb.render_frame(ren, cam, world).wait()
    got_accum = fb.map(osp.OSP_FB_COLOR)[::-1, :, :3]
    for a, b in ((first, got_first), (accum, got_accum)):
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert (d > 1).mean() < 1e-3, (d > 1).mean()        # same device, same streams: identical up to the odd rounding


def test_cancel_and_progress_of_a_running_frame(osp):
    """ospCancel / ospGetProgress (ospray.h; fb/FrameBuffer.cpp:180-187, fb/FrameBuffer.ih:27-50): a long frame (2048x2048,
    pathtracer, 32 spp) reports progress in [0, 1) while it runs and 1 when it is done; cancelled right after it was queued
    it ends in a fraction of the time, and the next frame on the same framebuffer renders normally again"""
    import time
    w = h = 2048
    world, keep = T.to_ospray(T.cornell_box())
    cam, _ = T.make_camera(world, w, h)
    ren = T.make_renderer("pathtracer", pixelSamples=32)
    fb = osp.FrameBuffer(w, h, osp.OSP_FB_RGBA32F, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM)
    fb.render_frame(ren, cam, world).wait()                # warm-up: scratch allocation, module load
    t0 = time.perf_counter()
    f = fb.render_frame(ren, cam, world)
    seen = []
    while not f.is_ready():
        seen.append(f.progress())
    f.wait()
    full = time.perf_counter() - t0
    assert f.progress() == 1.0
    assert all(0.0 <= p <= 1.0 for p in seen) and seen == sorted(seen)   # (the frame may end between is_ready() and progress())
    running = [p for p in seen if p < 1.0]
    assert len(running) > 3 and max(running) > 0.3, (len(running), max(running) if running else None)
    t0 = time.perf_counter()
    f = fb.render_frame(ren, cam, world)
    f.cancel()
    f.wait()
    cancelled = time.perf_counter() - t0
    assert cancelled < 0.5 * full, (cancelled, full)
    assert f.progress() == 1.0                             # a finished (here: ended early) task reports 1
    f = fb.render_frame(ren, cam, world)
    f.wait()
    assert fb.ray_counts()[0] == w * h * 32                # the flag is cleared by the next frame's beginFrame
