"""texture2d lookups on the device against the REFERENCE's own Texture2D_get (oracle/_ref), through the batch entry point
ospcuda_texture_sample of the C-ABI: every texel format, both filters, the three wrap modes, 2D coordinate transforms and
footprints across the whole MIP chain (texture/Texture2D.ispc:13-158, Texture2DFormats.ispc, TextureParam.ispc:10-22);
the MIP chain the device builds at ospCommit against texture/MipMapGeneration.ispc as restated in the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
from ospray_b200 import lib_path, testing as T  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

_FMT = {  # format -> (channels, dtype)
    "RGBA8": (4, np.uint8), "SRGBA": (4, np.uint8), "RGBA32F": (4, np.float32), "RGB8": (3, np.uint8), "SRGB": (3, np.uint8),
    "RGB32F": (3, np.float32), "R8": (1, np.uint8), "R32F": (1, np.float32), "L8": (1, np.uint8), "RA8": (2, np.uint8), "LA8": (2, np.uint8),
    "RGBA16": (4, np.uint16), "RGB16": (3, np.uint16), "RA16": (2, np.uint16), "R16": (1, np.uint16), "RA32F": (2, np.float32),
    "RGBA16F": (4, np.float16), "RGB16F": (3, np.float16), "RA16F": (2, np.float16), "R16F": (1, np.float16)}


def _image(fmt, h, w, seed):
    c, dt = _FMT[fmt]
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.0, 1.0, (h, w, c))
    if dt == np.uint8: a = (a * 255).astype(np.uint8)
    elif dt == np.uint16: a = (a * 65535).astype(np.uint16)
    else: a = (a * 2.0 - 0.5).astype(dt)
    return np.ascontiguousarray(a[..., 0] if c == 1 else a)


def _abi():
    L = C.CDLL(lib_path("libospcuda.so"), mode=C.RTLD_GLOBAL)
    L.ospcuda_texture_sample.argtypes = [C.POINTER(O.TextureParam), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    return L


def _device_sample(osp, ms, tp_host, st, fp):
    import torch
    tex = T.make_texture(ms.texture)
    tp = O.TextureParam()
    C.memmove(C.byref(tp), C.byref(tp_host), C.sizeof(tp))  # same transform as the checker computed
    tp.tex = C.cast(osp.ext().ospcudaTextureDevicePointer(tex.h), C.POINTER(O.Texture2D))
    d_st, d_fp = torch.from_numpy(st).cuda(), torch.from_numpy(fp).cuda()
    out = torch.zeros((len(fp), 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    assert _abi().ospcuda_texture_sample(C.byref(tp), d_st.data_ptr(), d_fp.data_ptr(), len(fp), out.data_ptr()) == 0
    from ospray_b200 import capi
    capi.lib()  # the render stream is the library's own: synchronise through it
    L = C.CDLL(lib_path("libospcuda.so"), mode=C.RTLD_GLOBAL)
    L.ospcuda_synchronize()
    return out.cpu().numpy()


def _coords(n, size, seed):
    rng = np.random.default_rng(seed)
    st = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
    st[: n // 8] = rng.integers(-2, 3, (n // 8, 2)).astype(np.float32)      # exactly on the borders
    fp = (2.0 ** rng.uniform(-12, 1, n) / max(size)).astype(np.float32)     # footprints from below a texel to the whole image
    fp[::7] = np.float32(np.finfo(np.float32).tiny)                          # flt_min: what a bare Texture2D lookup uses
    fp[::11] = 0.0
    return st, fp


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libospref.so not built (needs /root/reference)")
@pytest.mark.parametrize("fmt", list(_FMT))
@pytest.mark.parametrize("filt", ["linear", "nearest"])
def test_texture_lookup_matches_reference(osp, fmt, filt):
    wraps = [("repeat", "repeat"), ("mirror", "clamp"), ("clamp", "mirror")]
    k = list(_FMT).index(fmt)
    h, w = [(16, 16), (9, 7), (5, 12), (1, 8)][k % 4]
    ts = T.TextureSpec(_image(fmt, h, w, k), fmt, filter=filt, wrap=wraps[k % 3])
    ms = T.MapSpec(ts) if k % 2 else T.MapSpec(ts, rotation=33.0, scale=(0.7, 1.3), translation=(0.25, -0.4))
    st, fp = _coords(4096, (w, h), k)
    if fmt in ("RGB16F", "R16F"):
        # the reference's scalar (SYCL) form reads these two formats through half_to_float(int16) with a sycl::half ARGUMENT
        # (Texture2DFormats.ispc:28-33,354-371): the value is converted to an integer first, which is not what its ISPC form
        # (half = int16 bit pattern) does. The device follows the ISPC form; the checker here is the reference's lookup of
        # the same texels stored as floats, on the finest level (the coarser levels round to half on the device side)
        fp[:] = np.float32(np.finfo(np.float32).tiny)
        want, tp = O.ref_texture_sample(T.MapSpec(T.TextureSpec(ts.data.astype(np.float32), fmt.replace("16F", "32F"), filter=filt, wrap=ts.wrap),
                                                  ms.translation, ms.scale, ms.rotation), st, fp)
    else:
        want, tp = O.ref_texture_sample(ms, st, fp)
    got = _device_sample(osp, ms, tp, st, fp)
    # nearest: a coordinate that lands within rounding of a texel border may pick the neighbour (fma contraction in the
    # transform / wrap arithmetic); everything else agrees to float rounding (libm powf in the sRGB decode: 2 ulp)
    close = np.isclose(got, want, rtol=2e-5, atol=1e-5).all(axis=1)  # the float images hold values of both signs
    assert close.mean() >= (0.995 if filt == "nearest" else 0.999), (fmt, filt, close.mean(), got[~close][:4], want[~close][:4])
    if filt == "linear":
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4)


def test_uncommitted_texture_is_refused(osp):
    t = osp.Texture("texture2d")
    t.set("format", osp.Typed(osp.OSP_UINT, 0))
    with pytest.raises(osp.OSPRayError):
        t.commit()


def test_principled_takes_every_parameter_of_the_reference(osp):
    """principled: every factor, map and normal-map scale of Principled.cpp:33-150 is consumed (no 'unused parameter' warning)"""
    m = osp.Material("principled")
    tex = osp.texture2d(np.full((2, 2, 3), 0.5, np.float32), "RGB32F", 0, (0, 0))
    for name in O.PRMAP_NAMES:
        m.set("map_" + name, tex)
    m.set("baseColor", (0.2, 0.4, 0.6)).set("sheenColor", (1.0, 0.5, 0.5)).set("backlight", 0.3).set("thin", True)
    m.set("normal", 0.5).set("baseNormal", 0.7).set("coatNormal", 1.5).set("map_normal.rotation", 20.0)
    m.commit()
    assert not [e for e in osp.pending_errors()]


def test_microfacet_albedo_tables_match_reference():
    """the energy-compensation tables the device integrates (albedoTablesDirKernel / AvgKernel) against
    precomputeMicrofacetAlbedoTables of the reference (golden copy written by tests/golden/gen_ref_golden.py): same 1024
    stratified samples per entry, libm differences only"""
    import torch
    L = _abi()
    L.ospcuda_microfacet_albedo_tables.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
    ptr, n = C.c_void_p(), C.c_uint32()
    assert L.ospcuda_microfacet_albedo_tables(C.byref(ptr), C.byref(n)) == 0
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_albedo_tables.npy"))
    assert n.value == want.size
    got = torch.zeros(n.value, dtype=torch.float32, device="cuda")
    L.ospcuda_synchronize()
    cudart = C.CDLL("libcudart.so.12")
    assert cudart.cudaMemcpy(C.c_void_p(got.data_ptr()), ptr, C.c_size_t(4 * n.value), C.c_int(3)) == 0  # device to device
    got = got.cpu().numpy()
    # entries at cosThetaO = 0 (clamped to 1e-6 by the integrand) amplify last-bit differences of sincosf / powf / sqrtf between
    # CUDA and glibc by 1 / cosThetaO: a handful of them differ by up to 0.5 %; everything else agrees to 2e-4
    close = np.isclose(got, want, rtol=2e-4, atol=2e-6)
    assert close.mean() >= 0.995, (close.mean(), np.nonzero(~close)[0][:40])
    np.testing.assert_allclose(got, want, rtol=1e-2, atol=1e-4)
