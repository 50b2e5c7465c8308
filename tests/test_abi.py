"""The drop-in boundary: every symbol declared in include/*.h is exported by the matching shared library
(no compute calls here -- this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ospray_b200", "lib")


def declared(header, marker):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"#define[^\n]*", "", src)
    return sorted(set(re.findall(marker + r"[^;(]*?\b(\w+)\s*\(", src)))


def test_ospcuda_c_abi_symbols():
    names = declared("ospcuda.h", "OSPCUDA_API")
    assert len(names) >= 30 and "ospcuda_render_frame" in names and "ospcuda_bvh_build_group" in names
    L = C.CDLL(os.path.join(LIB, "libospcuda.so"))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_ospray_public_api_symbols():
    # the 51 functions of ospray.h (reference: ospray/include/ospray/ospray.h:151-346) + ospray_util.h helpers
    api = declared("ospray/ospray.h", "OSPRAY_INTERFACE")
    util = declared("ospray/ospray_util.h", "OSPRAY_INTERFACE")
    assert len(api) >= 50, len(api)
    L = C.CDLL(os.path.join(LIB, "libospray.so"), mode=C.RTLD_GLOBAL)
    missing = [n for n in api + util if not hasattr(L, n)]
    assert not missing, missing


def test_module_entry_point_and_extensions():
    # OSPCommon.cpp:749-782: dlopen "ospray_module_<name>" then dlsym "ospray_module_init_<name>"
    C.CDLL(os.path.join(LIB, "libospray.so"), mode=C.RTLD_GLOBAL)
    M = C.CDLL(os.path.join(LIB, "libospray_module_cuda.so"), mode=C.RTLD_GLOBAL)
    assert hasattr(M, "ospray_module_init_cuda")
    for n in declared("ospray_cuda_ext.h", "OSPRAY_INTERFACE"):
        assert hasattr(M, n), n
    init = M.ospray_module_init_cuda
    init.restype, init.argtypes = C.c_int, [C.c_int16, C.c_int16, C.c_int16]
    assert init(3, 3, 0) == 0          # OSP_NO_ERROR
    assert init(2, 9, 0) != 0          # OSP_VERSION_MISMATCH (ospray/common/OSPCommon.h:126-133)


def test_pod_struct_sizes_match_ctypes_mirror():
    # the oracle's ctypes mirror must agree with the C structs byte for byte: compile-time sizes via a tiny C probe
    import subprocess
    import tempfile
    from oracle import pyoracle as O
    names = ["OcData1D", "OcGeometry", "OcModel", "OcMaterial", "OcNode4", "OcBVH", "OcTransferFunction", "OcVolume", "OcGroup",
             "OcInstance", "OcLight", "OcWorld", "OcCamera", "OcRenderer", "OcFrameBuffer", "OcToneMapper", "OcRayHit"]
    src = '#include <stdio.h>\n#include "ospcuda.h"\nint main(){' + "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        out = dict(l.split() for l in subprocess.check_output([os.path.join(d, "p")]).decode().splitlines())
    mirror = {"OcData1D": O.Data1D, "OcGeometry": O.Geometry, "OcModel": O.Model, "OcMaterial": O.Material, "OcBVH": O.BVH,
              "OcTransferFunction": O.TransferFunction, "OcVolume": O.Volume, "OcGroup": O.Group, "OcInstance": O.Instance,
              "OcLight": O.Light, "OcWorld": O.World, "OcCamera": O.Camera, "OcRenderer": O.Renderer, "OcFrameBuffer": O.FrameBuffer,
              "OcToneMapper": O.ToneMapper, "OcRayHit": O.RayHit}
    for n, t in mirror.items():
        assert int(out[n]) == C.sizeof(t), (n, out[n], C.sizeof(t))
    assert int(out["OcNode4"]) == 128  # one cache line per wide node
