"""GPU parity tests: the cuda device (through the OSPRay C API and the libospcuda C-ABI) against the CPU oracle
on the same seeded inputs. Integer outputs (ids, hit/miss, 8-bit colour) are compared exactly up to the edge
band north_star allows; floating point within the tolerances written beside each assert."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest
import torch

import scenes
from ospray_b200 import imgdiff, lib_path
from ospray_b200 import testing as T

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_renders.npz"))
REF_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_renders.npz"))


@pytest.fixture(scope="module")
def abi(osp):
    L = C.CDLL(lib_path("libospcuda.so"), mode=C.RTLD_GLOBAL)
    L.ospcuda_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
    L.ospcuda_last_error.restype = C.c_char_p
    return L


def _random_rays(lo, hi, n, seed):
    rng = np.random.default_rng(seed)
    c, r = 0.5 * (lo + hi), 0.75 * np.linalg.norm(hi - lo)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = c + d * r
    tgt = c + rng.uniform(-0.5, 0.5, (n, 3)) * (hi - lo)
    dirs = tgt - org
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    from oracle import pyoracle as O
    rays = np.zeros(n, O.RAYHIT_DTYPE)
    rays["org"], rays["dir"], rays["t0"], rays["t"] = org.astype(np.float32), dirs.astype(np.float32), 0.0, np.inf
    rays["primID"] = rays["geomID"] = rays["instID"] = O.INVALID_ID
    # a third of the rays start inside the scene (secondary-ray like)
    k = n // 3
    rays["org"][:k] = (c + rng.uniform(-0.45, 0.45, (k, 3)) * (hi - lo)).astype(np.float32)
    return rays


def _gpu_trace(abi, world, rays, occlusion):
    t = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
    torch.cuda.synchronize()
    rc = abi.ospcuda_trace(world.device_pointer(), t.data_ptr(), len(rays), int(occlusion))
    assert rc == 0, abi.ospcuda_last_error()
    assert abi.ospcuda_synchronize() == 0
    from oracle import pyoracle as O
    return t.cpu().numpy().view(O.RAYHIT_DTYPE)


TRACE_SCENES = {
    "cornell_box": T.cornell_box, "random_spheres": T.random_spheres, "boxes": T.boxes, "box_mesh_8": lambda: T.box_mesh(8),
    "instanced": scenes.instanced_boxes_and_spheres, "many_instances": lambda: _many_instances(),
    "streamlines": scenes.few_streamlines, "linear_curves": scenes.linear_curves,
}


def _many_instances():
    """> 4 instances: exercises the top-level BVH (World.cpp:16-35 instID = array index)"""
    g = T.boxes((2, 2, 2)).instances[0].group
    inst = []
    for i in range(9):
        s = 0.05 + 0.01 * i
        inst.append(T.InstanceSpec(g, xfm=np.array([s, 0, 0, 0, s, 0, 0, 0, s, (i % 3) - 1.0, (i // 3) - 1.0, 0.1 * i], np.float32)))
    return T.SceneSpec(instances=inst, lights=[T.LightSpec("ambient", {})])


@pytest.mark.parametrize("name", list(TRACE_SCENES))
def test_traversal_matches_bruteforce(osp, abi, name):
    """ospcuda_trace (LBVH + wide-BVH traversal) vs the oracle's brute-force loop over every primitive:
    identical hit/miss and bit-identical t (same fused arithmetic); ids may differ only on exact-t ties."""
    from oracle import pyoracle as O
    spec = TRACE_SCENES[name]()
    sc = O.Scene(spec)
    world, keep = T.to_ospray(spec)
    rays = _random_rays(*sc.bounds, 20000, seed=zlib.crc32(name.encode()) % 1000)  # (str hash is salted per process)
    want = O.trace(sc, rays.copy())
    got = _gpu_trace(abi, world, rays, False)
    hit_w, hit_g = want["geomID"] != O.INVALID_ID, got["geomID"] != O.INVALID_ID
    assert hit_w.mean() > (0.01 if name in ("streamlines", "linear_curves") else 0.2)  # thin curves: few random rays hit
    np.testing.assert_array_equal(hit_g, hit_w)
    # t is bit-identical (same fused arithmetic) except on near-ties between two primitives (coplanar box
    # bottoms / floor, shared edges) where the visiting order decides which of two hits <= 2 ulp apart is kept
    t_bits_g, t_bits_w = got["t"].view(np.int32).astype(np.int64), want["t"].view(np.int32).astype(np.int64)
    assert np.abs(t_bits_g - t_bits_w).max() <= 2
    assert (t_bits_g != t_bits_w).mean() < 0.01
    ids_differ = (got["primID"] != want["primID"]) | (got["geomID"] != want["geomID"]) | (got["instID"] != want["instID"])
    assert ids_differ.mean() < 0.02, ids_differ.mean()   # (near-)equal-t ties only (cornell: box bottoms coplanar with the floor)
    same = ~ids_differ & hit_w & (t_bits_g == t_bits_w)
    np.testing.assert_array_equal(got["u"][same].view(np.uint32), want["u"][same].view(np.uint32))
    np.testing.assert_array_equal(got["Ng"][same].view(np.uint32), want["Ng"][same].view(np.uint32))
    # occlusion rays with a finite far end: rtcOccluded convention t = -inf (World.ih:311-335)
    rays["t"] = np.float32(0.6 * np.linalg.norm(sc.bounds[1] - sc.bounds[0]))
    want_o = O.trace(sc, rays.copy(), occlusion=True)
    got_o = _gpu_trace(abi, world, rays, True)
    np.testing.assert_array_equal(got_o["t"] == -np.inf, want_o["t"] == -np.inf)


ID_CASES = ["cornell_pt", "cornell_scivis", "cornell_ao", "boxes_scivis", "spheres_ao", "instanced_scivis", "instanced_pt", "mesh_ao",
            "translucent_scivis", "volume_geometry", "streamlines_scivis", "linear_curves_ao"]


@pytest.mark.parametrize("name", ID_CASES)
def test_primary_ids_and_depth(osp, name):
    """OSP_FB_ID_PRIMITIVE/OBJECT/INSTANCE + depth with OSP_PIXELFILTER_POINT: bit-exact ids except for rays
    within an epsilon of primitive edges (north_star); here <= 0.3 % of the pixels of a 96x80 image."""
    w, h = 96, 80
    want = scenes.oracle_render(name, w, h, pixel_filter="point")
    fb, keep = scenes.device_render(osp, name, w, h, pixel_filter="point")
    bad = np.zeros((h, w), bool)
    for ch, key in ((osp.OSP_FB_ID_PRIMITIVE, "primID"), (osp.OSP_FB_ID_OBJECT, "objID"), (osp.OSP_FB_ID_INSTANCE, "instID")):
        bad |= fb.map(ch) != want.buf[key]
    assert bad.mean() <= 0.003, bad.mean()
    d, dw = fb.map(osp.OSP_FB_DEPTH), want.buf["depth"]
    ok = ~bad
    assert (np.isinf(d[ok]) == np.isinf(dw[ok])).all()
    fin = ok & np.isfinite(dw)
    # curves: the two sides' camera rays differ in the last bit (fma contraction in the ray set-up) and a grazing hit on a
    # capsule of radius ~0.01 amplifies that; with identical rays t is bit-exact (test_traversal_matches_bruteforce)
    np.testing.assert_allclose(d[fin], dw[fin], rtol=1e-3 if "curves" in name or "streamlines" in name else 2e-5)
    assert (want.buf["primID"] != 0xFFFFFFFF).mean() > (0.02 if "curves" in name or "streamlines" in name else 0.1)  # thin curves cover little


@pytest.mark.parametrize("name", [n for n in scenes.CASES if n not in scenes.REF_UNSUPPORTED | scenes.DEVICE_UNSUPPORTED])
def test_images_match_reference_golden(osp, name):
    """the device against vectors produced by the REFERENCE's own kernels (tests/golden/ref_renders.npz, made by
    gen_ref_golden.py from oracle/_ref): same bars as against the oracle - >= 98 % of the pixels within 2e-3 relative
    (+1e-4), mean colour within 1 %, ospTestSuite metric on the sRGB images; ids exact on >= 99.7 % of the pixels for
    the geometry cases (the scalar reference form does not write volume first-hit ids)."""
    fb, keep = scenes.device_render(osp, name)
    got = fb.map(osp.OSP_FB_COLOR)
    wc = REF_GOLD[name + "/color"]
    err = np.abs(got - wc).max(axis=-1)
    ptvol = name in scenes.PT_VOLUME_CASES  # statistical bars, see scenes.PT_VOLUME_CASES
    assert (err > 1e-4 + 2e-3 * np.abs(wc).max(axis=-1)).mean() <= (0.12 if ptvol else 0.08 if name in scenes.REF_UNDEFINED_STATE_CASES else 0.02)
    np.testing.assert_allclose(got.reshape(-1, 4).mean(axis=0), wc.reshape(-1, 4).mean(axis=0), rtol=0.03 if ptvol else 0.01, atol=1e-4)
    from oracle import pyoracle as O
    a = O.color_convert(np.ascontiguousarray(got), 2).view(np.uint8).reshape(got.shape)
    b = O.color_convert(wc, 2).view(np.uint8).reshape(got.shape)
    assert imgdiff.compare(a, b)["ok"]
    if not name.startswith(("gravity", "volume")):
        bad = np.zeros(got.shape[:2], bool)
        for ch, key in ((osp.OSP_FB_ID_PRIMITIVE, "primID"), (osp.OSP_FB_ID_OBJECT, "objID"), (osp.OSP_FB_ID_INSTANCE, "instID")):
            bad |= fb.map(ch) != REF_GOLD[f"{name}/{key}"]
        # the default Gaussian pixel filter jitters the primary ray: edge pixels may pick the neighbouring primitive
        assert bad.mean() <= 0.01, bad.mean()


@pytest.mark.parametrize("name", [n for n in scenes.CASES if n not in scenes.DEVICE_UNSUPPORTED])
def test_images_match_oracle(osp, name):
    """float accumulation buffers after the case's frames: same sampler streams on both sides, so pixels agree
    to rounding (fma contraction, libm) except where a last-bit difference flips a discrete decision. Bars:
    >= 98 % of the pixels within 2e-3 relative (+1e-4 absolute); mean colour within 1 %; the 8-bit sRGB images
    pass the reference's own ospTestSuite metric (test_tools.h:26-29); primary ray counts are identical."""
    from oracle import pyoracle as O
    ref_only = name in scenes.ORACLE_UNSUPPORTED  # textures: the checker is the reference's own kernels (oracle/_ref)
    if ref_only and not O.ref_available():
        pytest.skip("oracle/_ref/libospref.so not built; the committed reference goldens cover this case")
    want = scenes.oracle_render(name, backend="ref" if ref_only else "oracle")
    fb, keep = scenes.device_render(osp, name)
    got = fb.map(osp.OSP_FB_COLOR)
    wc = want.buf["color"]
    err = np.abs(got - wc).max(axis=-1)
    tol = 1e-4 + 2e-3 * np.abs(wc).max(axis=-1)
    frac_bad = (err > tol).mean()
    ptvol = name in scenes.PT_VOLUME_CASES  # statistical bars, see scenes.PT_VOLUME_CASES
    assert frac_bad <= (0.12 if ptvol else 0.08 if name in scenes.REF_UNDEFINED_STATE_CASES else 0.02), frac_bad
    np.testing.assert_allclose(got.reshape(-1, 4).mean(axis=0), wc.reshape(-1, 4).mean(axis=0), rtol=0.03 if ptvol else 0.01, atol=1e-4)
    from oracle import pyoracle as O
    a = O.color_convert(np.ascontiguousarray(got), 2).view(np.uint8).reshape(got.shape)
    b = O.color_convert(wc, 2).view(np.uint8).reshape(got.shape)
    res = imgdiff.compare(a, b)
    assert res["ok"], res
    rc = fb.ray_counts()
    assert rc[0] == int(want.counters[0])
    if ref_only:  # the reference traces every shadow segment as a closest-hit query: only the total is comparable
        assert 0.85 * int(want.counters.sum()) - 4 <= sum(rc) <= 1.01 * int(want.counters.sum()) + 4
        return
    assert abs(rc[0] + rc[1] - int(want.counters[:2].sum())) <= (0.05 if ptvol else 0.01) * int(want.counters[:2].sum()) + 4
    # shadow segments: the device asks ONE any-hit question per shadow ray and walks the reference's closest-hit loop only
    # when the occluder it found is not opaque; which occluder that is depends on the hierarchy, so behind transparent
    # surfaces it may count fewer segments than the reference (never noticeably more)
    assert rc[2] <= 1.01 * int(want.counters[2]) + 4 and rc[2] >= (0.80 if ptvol else 0.85) * int(want.counters[2]) - 4
    # the committed golden fixture (oracle output at generation time) tells the same story
    gold = (REF_GOLD if ref_only else GOLD)[name + "/color"]
    np.testing.assert_allclose(got.reshape(-1, 4).mean(axis=0), gold.reshape(-1, 4).mean(axis=0), rtol=0.03 if ptvol else 0.01, atol=1e-4)
    if ptvol:  # the first vertex is sampled before any rounding can tell the two sides apart: depth of most pixels is identical
        d, dw = fb.map(osp.OSP_FB_DEPTH), want.buf["depth"]
        assert (np.isinf(d) == np.isinf(dw)).mean() >= 0.97
        both = np.isfinite(d) & np.isfinite(dw)
        assert (np.abs(d[both] - dw[both]) <= 1e-4 * np.abs(dw[both])).mean() >= 0.9


def test_aov_channels(osp):
    """normal / albedo AOVs of the scivis renderer (SciVis.ispc:150-165) against the oracle"""
    want = scenes.oracle_render("boxes_scivis", pixel_filter="point")
    fb, keep = scenes.device_render(osp, "boxes_scivis", pixel_filter="point")
    for ch, key in ((osp.OSP_FB_NORMAL, "normal"), (osp.OSP_FB_ALBEDO, "albedo")):
        g, w_ = fb.map(ch), want.buf[key]
        bad = np.abs(g - w_).max(axis=-1) > 1e-4
        assert bad.mean() < 0.005, (key, bad.mean())


@pytest.mark.parametrize("fmt", ["srgba", "rgba8", "rgba32f"])
def test_color_formats_through_api(osp, fmt):
    """ospMapFrameBuffer(OSP_FB_COLOR) for each OSPFrameBufferFormat == oracle conversion of the float buffer"""
    from oracle import pyoracle as O
    code = {"rgba8": (osp.OSP_FB_RGBA8, 1), "srgba": (osp.OSP_FB_SRGBA, 2), "rgba32f": (osp.OSP_FB_RGBA32F, 3)}[fmt]
    fb, keep = scenes.device_render(osp, "cornell_scivis", fmt=code[0], channels=osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM)
    got = fb.map(osp.OSP_FB_COLOR)
    ptr, bpp = fb.device_pointer(osp.OSP_FB_ACCUM)
    assert bpp == 16
    acc = np.zeros((scenes.H * scenes.W, 4), np.float32)
    L = C.CDLL(lib_path("libospcuda.so"))
    L.ospcuda_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    assert L.ospcuda_memcpy_d2h(acc.ctypes.data, ptr, acc.nbytes) == 0
    want = O.color_convert(acc, code[1])
    if fmt == "rgba32f":
        np.testing.assert_array_equal(got.reshape(-1, 4), want)
    else:
        diff = np.abs(got.reshape(-1, 4).astype(int) - want.view(np.uint8).reshape(-1, 4).astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01   # powf last-bit differences may move a code by 1


def test_rerender_is_deterministic_and_resets(osp):
    """ospResetAccumulation restarts the sampler sequence (FrameBuffer.cpp:95-103): identical images"""
    fb, keep = scenes.device_render(osp, "cornell_pt")
    a = fb.map(osp.OSP_FB_COLOR)
    world, _, cam, ren = keep
    fb.reset_accumulation()
    for _ in range(scenes.CASES["cornell_pt"][3]):
        fb.render_frame(ren, cam, world).wait()
    np.testing.assert_array_equal(fb.map(osp.OSP_FB_COLOR), a)


def test_big_mesh_traversal_vs_bruteforce(osp, abi):
    """BVH build + traversal on a 1.2 M-triangle mesh (46^3 boxes) against brute force on a ray sample"""
    from oracle import pyoracle as O
    spec = T.box_mesh(46)
    sc = O.Scene(spec)
    world, keep = T.to_ospray(spec)
    rays = _random_rays(*sc.bounds, 400, seed=11)
    want = O.trace(sc, rays.copy())
    got = _gpu_trace(abi, world, rays, False)
    d = np.abs(got["t"].view(np.int32).astype(np.int64) - want["t"].view(np.int32).astype(np.int64))
    assert d.max() <= 2 and (d > 0).mean() < 0.01
    assert ((got["primID"] != want["primID"]).mean()) < 0.01


@pytest.mark.parametrize("kind", ["perspective", "panoramic"])
def test_pick_matches_oracle_primary_hit(osp, kind):
    """ospPick (Renderer.cpp:94-137, Renderer.ispc:43-84): camera ray through the screen position, centre of the lens, no
    pixel filter; the hit primitive, model and instance equal the oracle's primary hit for the same ray and the world
    position is org + t * dir."""
    from oracle import pyoracle as O
    spec = scenes.instanced_boxes_and_spheres()
    world, keep = T.to_ospray(spec)
    w, h = 64, 48
    cam, pose = T.make_camera(world, w, h, kind=kind)
    ren = T.make_renderer("scivis")
    fb = osp.FrameBuffer(w, h, osp.OSP_FB_RGBA32F, osp.OSP_FB_COLOR)
    sc = O.Scene(spec)
    ocam = O.make_camera(pose["position"], pose["direction"], pose["up"], aspect=np.float32(w) / np.float32(h), kind=kind)
    oren = O.RendererParams("scivis", pixelFilter="point")
    ofb = O.HostFrameBuffer(w, h, ("color",))
    step = 8 if kind == "perspective" else 1  # (the scene covers a few per cent of the panorama)
    px, py = np.meshgrid(np.arange(step // 2, w, step), np.arange(step // 2, h, step))
    px, py = px.ravel().astype(np.uint32), py.ravel().astype(np.uint32)
    rays = O.camera_rays(ocam, oren, ofb, px, py)  # point filter: rays through the pixel centres
    want = O.trace(sc, rays.copy())
    hits = 0
    for i in range(len(px)):
        has, pos, inst, model, prim = osp.pick(fb, ren, cam, world, (px[i] + 0.5) / w, (py[i] + 0.5) / h)
        hit = want["geomID"][i] != O.INVALID_ID
        assert has == hit, (i, has, hit)
        if hit:
            hits += 1
            assert prim == want["primID"][i]
            np.testing.assert_allclose(pos, want["org"][i] + want["dir"][i] * want["t"][i], rtol=1e-5, atol=1e-5)
            assert inst and model
    assert (3 if kind == "perspective" else 0) < hits < len(px)


@pytest.mark.gpu
def test_debug_renderer_refuses_unsupported_methods(osp):
    """the volume method needs data this device does not carry: the commit fails through the API's error convention
    (API.cpp:45-66) instead of rendering something else; unknown names fall back to the test frame like the reference
    (DebugRenderer.cpp:18-52)"""
    for method in ("volume",):
        r = osp.Renderer("debug")
        r.set("method", method)
        with pytest.raises(osp.OSPRayError, match="not supported"):
            r.commit()
    osp.Renderer("debug").set("method", "noSuchMethod").commit()


@pytest.mark.gpu
def test_builder_reports_the_traversal_stack_need(osp, abi):
    """every hierarchy the builder hands out states the deepest traversal stack it can ask for; the device compares it with
    OSPCUDA_TRAVERSAL_STACK at commit (a scene beyond it is refused instead of traversed with dropped subtrees)"""
    abi.ospcuda_bvh_peak_stack_need.restype = C.c_uint32
    abi.ospcuda_bvh_peak_stack_need(1)
    world, keep = T.to_ospray(T.box_mesh(6))      # 2592 triangles: 4-wide depth 6..7 (+ a one-leaf top level)
    need = abi.ospcuda_bvh_peak_stack_need(1)
    assert 6 <= need <= 30, need
    world, keep = T.to_ospray(T.cornell_box())    # 34 records
    assert 1 <= abi.ospcuda_bvh_peak_stack_need(1) <= 12


def _clustered_mesh(levels=20, per_level=6):
    """triangles in clusters that shrink geometrically along the three axes: cluster (j, a) sits at 2^-j on axis a and is
    2^-j wide, so that every bit of the 63-bit Morton key peels exactly one cluster off: the radix hierarchy is one long
    chain (3 x levels deep)."""
    rng = np.random.default_rng(5)
    verts, index = [], []

    def cluster(c, s):
        for k in range(per_level):
            tri = c + 0.2 * s * rng.uniform(0, 1, (3, 3))
            index.append([len(verts), len(verts) + 1, len(verts) + 2])
            verts.extend(tri)
    cluster(np.array([0.9, 0.9, 0.9]), 0.5)              # anchors the scene bounds at the unit cube
    for j in range(1, levels + 1):
        for a in range(3):
            c = np.zeros(3)
            c[a] = 2.0 ** -j
            cluster(c, 2.0 ** -j)
    geom = T.GeometrySpec("mesh", vertex=np.array(verts, np.float32), index=np.array(index, np.uint32))
    return T.SceneSpec(instances=[T.InstanceSpec(T.GroupSpec(models=[T.ModelSpec(geom, T.MaterialSpec())]))], lights=[])


@pytest.mark.gpu
def test_clustered_mesh_deep_hierarchy_is_rebuilt_not_refused(osp, abi):
    """a hierarchy that would need more traversal stack than the device keeps (OSPCUDA_TRAVERSAL_STACK) is built again over a
    coarser Morton grid with larger leaves instead of refusing the scene; the rebuilt one answers like brute force"""
    from oracle import pyoracle as O
    import os
    abi.ospcuda_bvh_peak_stack_need.restype = C.c_uint32
    abi.ospcuda_bvh_last_stack_need.restype = C.c_uint32
    abi.ospcuda_bvh_peak_stack_need(1)
    spec = _clustered_mesh()
    # the chain needs ~38 entries with the default build (treelet restructuring already flattens it; the device keeps 64):
    # a limit of 28 makes the device take the fall-back for it, while traversal can still walk whatever comes out
    os.environ["OSPCUDA_STACK_LIMIT"] = "28"
    try:
        world, keep = T.to_ospray(spec)                   # would raise if the group were refused
    finally:
        del os.environ["OSPCUDA_STACK_LIMIT"]
    peak = abi.ospcuda_bvh_peak_stack_need(1)
    assert peak > 28, peak                                # the first build was too deep: the fall-back ran
    sc = O.Scene(spec)
    rays = _random_rays(*sc.bounds, 4000, seed=3)
    rays["org"][:2000] = np.float32(1.5)                  # half the rays look at the corner the clusters shrink towards
    d = -np.ones((2000, 3), np.float32) + np.random.default_rng(1).uniform(-0.3, 0.3, (2000, 3)).astype(np.float32)
    rays["dir"][:2000] = d / np.linalg.norm(d, axis=1, keepdims=True)
    want = O.trace(sc, rays.copy())
    got = _gpu_trace(abi, world, rays, False)
    assert ((want["geomID"] != O.INVALID_ID) == (got["geomID"] != O.INVALID_ID)).all()
    assert (got["t"].view(np.uint32) == want["t"].view(np.uint32)).all()
    assert (want["geomID"] != O.INVALID_ID).mean() > 0.02   # thin clusters: few of the random rays hit, 150 is plenty


@pytest.mark.gpu
def test_model_and_transfer_function_commits_reach_the_next_frame(osp):
    """OSPRay semantics: committing a geometric model (here: another material) or a transfer function takes effect in the
    next frame without committing the group / world again (the reference's Group holds pointers to its models' structs,
    Group.cpp:139-157); the device refreshes its snapshots before the frame"""
    w, h = 64, 48
    spec = T.cornell_box()
    world, keep = T.to_ospray(spec)
    cam, _ = T.make_camera(world, w, h)
    ren = T.make_renderer("scivis", aoSamples=0)
    fb = osp.FrameBuffer(w, h, osp.OSP_FB_RGBA32F, osp.OSP_FB_COLOR)
    fb.render_frame(ren, cam, world).wait()
    before = fb.map(osp.OSP_FB_COLOR)
    model = [o for o in keep if isinstance(o, osp.Object) and o.kind == osp.OSP_GEOMETRIC_MODEL][0]
    dark = osp.Material("obj")
    dark.set("kd", (0.05, 0.05, 0.05)).commit()
    model.set("material", dark).commit()                  # only the model is committed
    fb.render_frame(ren, cam, world).wait()
    after = fb.map(osp.OSP_FB_COLOR)
    assert after[..., :3].mean() < 0.5 * before[..., :3].mean()
