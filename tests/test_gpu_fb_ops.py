"""Framebuffer pixel ops through the libospcuda C-ABI (device pointers from torch tensors) vs the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from ospray_b200 import lib_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(osp):
    lib = C.CDLL(lib_path("libospcuda.so"), mode=C.RTLD_GLOBAL)
    lib.ospcuda_last_error.restype = C.c_char_p
    return lib


def _dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    torch.cuda.synchronize()
    return t


def _zeros(*a, **kw):
    """torch fills on ITS stream; libospcuda works on its own: the fill must have finished before the library writes"""
    t = torch.zeros(*a, **kw)
    torch.cuda.synchronize()
    return t


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (64, 64), (333, 517), (1024, 1024)])
def test_accumulate_matches_oracle(L, shape):
    from oracle import pyoracle as O
    h, w = shape
    rng = np.random.default_rng(w * 7 + h)
    host = O.HostFrameBuffer(w, h, ("color", "depth"), variance=True)
    color = torch.full((h * w * 4,), 99.0, dtype=torch.float32, device="cuda")
    var = _zeros(h * w * 4, dtype=torch.float32, device="cuda")
    depth = _zeros(h * w, dtype=torch.float32, device="cuda")
    dfb = O.FrameBuffer()
    dfb.sizeX, dfb.sizeY = w, h
    dfb.color, dfb.variance, dfb.depth = color.data_ptr(), var.data_ptr(), depth.data_ptr()
    L.ospcuda_fb_accumulate.argtypes = [C.POINTER(O.FrameBuffer), C.c_void_p, C.c_void_p]
    for i in range(4):
        s = rng.uniform(0, 2, (h, w, 4)).astype(np.float32)
        z = rng.uniform(0.5, 9, (h, w)).astype(np.float32)
        if i == 2:
            z[::2] = np.inf
        host.begin_frame()
        O.lib().oracle_fb_accumulate(C.byref(host.sh), s.ctypes.data, z.ctypes.data)
        dfb.frameID, dfb.accumulateVariance = host.sh.frameID, host.sh.accumulateVariance
        ds, dz = _dev(s), _dev(z)
        assert L.ospcuda_fb_accumulate(C.byref(dfb), ds.data_ptr(), dz.data_ptr()) == 0, L.ospcuda_last_error()
        assert L.ospcuda_synchronize() == 0
    # lerp(f, old, new) = (1-f)*old + f*new: the GPU contracts to fma, the oracle does not -> <= 1 ulp
    np.testing.assert_allclose(color.cpu().numpy().reshape(h, w, 4), host.buf["color"], rtol=2.5e-7, atol=1e-7)
    np.testing.assert_allclose(var.cpu().numpy().reshape(h, w, 4), host.buf["variance"], rtol=2.5e-7, atol=1e-7)
    np.testing.assert_array_equal(depth.cpu().numpy().reshape(h, w), host.buf["depth"])


@pytest.mark.parametrize("n", [1, 1000, 1 << 20])
@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_color_convert_matches_oracle(L, n, fmt):
    from oracle import pyoracle as O
    rng = np.random.default_rng(n + fmt)
    x = rng.uniform(-0.2, 1.4, (n, 4)).astype(np.float32)
    x[: min(n, 8)] = [0, 0.0031308, 0.5, 1.0]
    dx = _dev(x)
    out = _zeros(n * (16 if fmt == 3 else 4), dtype=torch.uint8, device="cuda")
    L.ospcuda_fb_color_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    assert L.ospcuda_fb_color_convert(dx.data_ptr(), out.data_ptr(), n, fmt) == 0, L.ospcuda_last_error()
    assert L.ospcuda_synchronize() == 0
    want = O.color_convert(x, fmt)
    if fmt == 3:
        np.testing.assert_array_equal(out.cpu().numpy().view(np.float32).reshape(n, 4), want)
    elif fmt == 1:
        np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), want)   # pure rounding: bit-exact
    else:
        d = np.abs(out.cpu().numpy().reshape(n, 4).astype(int) - want.view(np.uint8).reshape(n, 4).astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 5e-3  # powf(c, 1/2.4) differs in the last bit between libm and CUDA


def test_tonemap_matches_oracle(L):
    from oracle import pyoracle as O
    n = 100003
    x = np.random.default_rng(3).uniform(0, 6, (n, 4)).astype(np.float32)
    tm = O.default_tonemapper(exposure=1.3)
    dx = _dev(x)
    out = _zeros(n * 4, dtype=torch.float32, device="cuda")
    L.ospcuda_fb_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(O.ToneMapper)]
    assert L.ospcuda_fb_tonemap(dx.data_ptr(), out.data_ptr(), n, C.byref(tm)) == 0, L.ospcuda_last_error()
    assert L.ospcuda_synchronize() == 0
    np.testing.assert_allclose(out.cpu().numpy().reshape(n, 4), O.tonemap(x, tm), atol=3e-6, rtol=1e-5)


@pytest.mark.parametrize("shape", [(8, 8), (70, 45), (1024, 768)])
def test_variance_matches_oracle(L, shape):
    from oracle import pyoracle as O
    h, w = shape
    rng = np.random.default_rng(5)
    c = rng.uniform(0, 1, (h, w, 4)).astype(np.float32)
    v = (c + rng.normal(0, 0.05, (h, w, 4))).astype(np.float32)
    c[: h // 2, : w // 3] = [0, 0, 0, 1]  # den2 == 0 pixels are skipped
    tasks = ((w + 7) // 8) * ((h + 7) // 8)
    out = _zeros(tasks, dtype=torch.float32, device="cuda")
    dc, dv = _dev(c), _dev(v)
    L.ospcuda_fb_variance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    assert L.ospcuda_fb_variance(dc.data_ptr(), dv.data_ptr(), w, h, 8, out.data_ptr()) == 0, L.ospcuda_last_error()
    assert L.ospcuda_synchronize() == 0
    np.testing.assert_allclose(out.cpu().numpy(), O.variance(c, v), rtol=2e-5, atol=1e-6)  # summation order differs


@pytest.mark.parametrize("size", [(64, 64), (200, 136), (1920, 1080)])
@pytest.mark.parametrize("bpp", [4, 16])
def test_tile_pack_unpack_roundtrip(L, size, bpp):
    """pack -> unpack over every rank reproduces the image (the NCCL gather staging), and the packed layout is
    the one the oracle defines"""
    from oracle import pyoracle as O
    O.lib().oracle_fb_pack_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    from ospray_b200 import multigpu
    w, h = size
    n = 3
    img = np.random.default_rng(1).integers(0, 255, (h, w, bpp), dtype=np.uint8)
    dimg = _dev(img)
    dst = torch.zeros_like(dimg)
    torch.cuda.synchronize()
    for fn in (L.ospcuda_fb_pack_tiles, L.ospcuda_fb_unpack_tiles):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    for r in range(n):
        shard = multigpu.max_owned(w, h, n) * 64 * 64 * bpp
        packed = _zeros(shard, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        assert L.ospcuda_fb_pack_tiles(packed.data_ptr(), dimg.data_ptr(), w, h, bpp, n, r) == 0
        assert L.ospcuda_synchronize() == 0
        want = np.zeros(shard, np.uint8)
        O.lib().oracle_fb_pack_tiles(want.ctypes.data, img.ctypes.data, w, h, bpp, n, r, 1)
        np.testing.assert_array_equal(packed.cpu().numpy(), want)
        assert L.ospcuda_fb_unpack_tiles(dst.data_ptr(), packed.data_ptr(), w, h, bpp, n, r) == 0
        assert L.ospcuda_synchronize() == 0
    np.testing.assert_array_equal(dst.cpu().numpy(), img)
