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


REF_INCLUDE = "/root/reference/ospray/include/ospray"


def _prototypes(path):
    """{function name: normalised parameter type list} of the OSPRAY_INTERFACE declarations of a header"""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"OSPRAY_INTERFACE\s+([^;{]*?)\b(\w+)\s*\(([^;{]*?)\)\s*(?:OSP_DEPRECATED[^;]*)?;", src, flags=re.S):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        plist = []
        for prm in params.split(","):
            prm = re.sub(r"OSP_DEFAULT_VAL\([^)]*\)", "", prm)
            prm = re.sub(r"=\s*[^,]*", "", prm)
            toks = re.findall(r"[A-Za-z_][A-Za-z_0-9]*|\*", prm)
            # drop the parameter name (last identifier) when a type precedes it
            if len([t for t in toks if t != "*"]) > 1 and toks[-1] != "*":
                toks = toks[:-1]
            toks = [t for t in toks if t not in ("const",)]
            if toks and toks != ["void"]:
                plist.append(" ".join(toks))
        out[name] = (" ".join(re.findall(r"[A-Za-z_][A-Za-z_0-9]*|\*", ret.replace("const", ""))), tuple(plist))
    return out


@pytest.mark.skipif(not os.path.isdir(REF_INCLUDE), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("header", ["ospray.h", "ospray_util.h"])
def test_public_headers_declare_what_the_reference_declares(header):
    """the repo's include/ospray/*.h against the REFERENCE's headers: the same function names with the same return and
    parameter types (a renamed or missing entry point would pass a self-referential check), and every one exported"""
    ref = _prototypes(os.path.join(REF_INCLUDE, header))
    ours = _prototypes(os.path.join(ROOT, "include", "ospray", header))
    assert len(ref) >= 20
    missing = sorted(set(ref) - set(ours))
    extra = sorted(set(ours) - set(ref))
    assert not missing, missing
    assert not extra, extra
    wrong = {n: (ref[n], ours[n]) for n in ref if ref[n] != ours[n]}
    assert not wrong, wrong
    L = C.CDLL(os.path.join(LIB, "libospray.so"), mode=C.RTLD_GLOBAL)
    assert not [n for n in ref if not hasattr(L, n)]


@pytest.mark.skipif(not os.path.isdir(REF_INCLUDE), reason="the reference tree is only mounted in the build container")
def test_enums_have_the_reference_values():
    """every enumerator of the reference's OSPEnums.h exists in the repo's header with the same value: names are read from
    the reference header, values come from the C compiler on both sides (the repo's header builds its tables with macros)"""
    import subprocess
    import tempfile
    src = open(os.path.join(REF_INCLUDE, "OSPEnums.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = []
    for body in re.findall(r"enum[^{;]*\{([^}]*)\}", src, flags=re.S):
        for item in body.split(","):
            m = re.match(r"\s*(OSP_\w+)", item)
            if m:
                names.append(m.group(1))
    assert len(names) > 150

    def values(include_dir, header):
        prog = '#include <stdio.h>\n#include <stdint.h>\n#include "%s"\nint main(){' % header
        prog += "".join('printf("%s %%lld\\n", (long long)%s);' % (n, n) for n in names) + "return 0;}"
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "e.c"), "w").write(prog)
            r = subprocess.run(["gcc", "-I", include_dir, os.path.join(d, "e.c"), "-o", os.path.join(d, "e")], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[:2000]
            return dict(l.split() for l in subprocess.check_output([os.path.join(d, "e")]).decode().splitlines())
    ref = values(os.path.dirname(REF_INCLUDE), "ospray/OSPEnums.h")
    ours = values(os.path.join(ROOT, "include"), "ospray/OSPEnums.h")
    wrong = {n: (ref[n], ours[n]) for n in names if ref[n] != ours[n]}
    assert not wrong, wrong


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
