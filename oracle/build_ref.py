#!/usr/bin/env python3
"""Build oracle/_ref/libospref.so: the REFERENCE's own render kernels, compiled where they lie.

TEST INFRASTRUCTURE ONLY.  The reference (OSPRay v3.3.0) cannot be built with its own build system here (needs
ISPC, rkcommon, Embree, Open VKL, TBB; none installed, no network).  But its kernel sources
(/root/reference/modules/cpu/**/*.ispc + .ih) are written to compile as plain scalar C++ as well: that is how the
reference's `ospray_module_gpu_kernels` target builds them (modules/cpu/CMakeLists.txt:412-480: LANGUAGE CXX,
`uniform= varying= unmasked= export= programCount=1 programIndex=0 OSPRAY_TARGET_SYCL OSPRAY_GLOBAL=static
OSPRAY_BEGIN_ISPC_NAMESPACE=namespace ispc{ rtcIntersectV=rtcIntersect1 ...`).  This recipe runs g++ with exactly
those definitions on the reference's source list (OSPRAY_KERNEL_SOURCES, modules/cpu/CMakeLists.txt:120-226, plus
the volume kernels when --volumes) straight from /root/reference; nothing is copied into the repo, objects go to a
temporary directory, and the only output is oracle/_ref/libospref.so (git-ignored, travels with gpurun).

The externals the kernels include are NOT in /root/reference; oracle/refshim/ provides stand-ins:
  rkcommon/math/*.ih  -> refshim/shim_core.h       (vec/linear/affine/box helpers + ISPC stdlib names, restated)
  embree4/rtcore.h    -> refshim/embree4/rtcore.h  + oracle/ref_embree_stub.cpp (brute-force scene queries)
  openvkl/*.h         -> refshim/openvkl/*         + oracle/ref_vkl_stub.cpp    (structuredRegular sampler/iterator)
  sycl/sycl.hpp       -> refshim/sycl/sycl.hpp     (host-thread parallel_for, specialization constants)
  OSPConfig.h         -> refshim/OSPConfig.h       (CMake-generated in the reference build)
oracle/ref_driver.cpp fills the reference's shared structs from the Oc* PODs and calls the reference's launchers.

  python oracle/build_ref.py [--reference /root/reference] [--jobs N]
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))

# modules/cpu/CMakeLists.txt:120-226 (OSPRAY_KERNEL_SOURCES), in the reference's order
KERNEL_SOURCES = """
ISPCDevice.ispc common/World.ispc fb/FrameBufferDispatch.ispc fb/LocalFB.ispc fb/SparseFB.ispc fb/frame_ops/Blur.ispc
fb/frame_ops/ColorConversion.ispc fb/frame_ops/Debug.ispc fb/frame_ops/ToneMapper.ispc fb/frame_ops/Variance.ispc
pf/PixelFilterDispatch.ispc pf/LUTPixelFilter.ispc camera/CameraDispatch.ispc camera/PerspectiveCamera.ispc
camera/OrthographicCamera.ispc camera/PanoramicCamera.ispc geometry/GeometryDispatch.ispc geometry/GeometricModel.ispc
geometry/Boxes.ispc geometry/Curves.ispc geometry/Spheres.ispc geometry/Planes.ispc geometry/Mesh.ispc
lights/LightDispatch.ispc lights/Light.ispc lights/AmbientLight.ispc lights/DirectionalLight.ispc lights/PointLight.ispc
lights/SpotLight.ispc lights/QuadLight.ispc lights/CylinderLight.ispc lights/HDRILight.ispc math/Distribution1D.ispc
math/Distribution2D.ispc math/spectrum.ispc render/Material.ispc render/MaterialDispatch.ispc render/Renderer.ispc
render/util.ispc render/debug/DebugRenderer.ispc render/ao/AORenderer.ispc render/ao/surfaces.ispc render/scivis/SciVis.ispc
render/scivis/surfaces.ispc render/scivis/lightAlpha.ispc render/pathtracer/PathTracer.ispc render/pathtracer/PathSampler.ispc
render/pathtracer/GeometryLight.ispc render/pathtracer/VirtualLight.ispc render/pathtracer/TransparentShadow.ispc
render/pathtracer/NextEventEstimation.ispc render/bsdfs/BSDF.ispc render/bsdfs/MicrofacetAlbedoTables.ispc
render/materials/OBJ.ispc render/materials/Principled.ispc render/materials/CarPaint.ispc render/materials/Velvet.ispc
render/materials/Metal.ispc render/materials/Alloy.ispc render/materials/ThinGlass.ispc render/materials/Glass.ispc
render/materials/MetallicPaint.ispc render/materials/Plastic.ispc render/materials/Luminous.ispc render/materials/Mix.ispc
texture/TextureDispatch.ispc texture/Texture2DFormats.ispc texture/Texture2D.ispc texture/TextureParam.ispc
""".split()
# modules/cpu/CMakeLists.txt:229-241 (OSPRAY_ENABLE_VOLUMES additions)
VOLUME_SOURCES = """
geometry/Isosurfaces.ispc volume/Volume.ispc volume/transferFunction/LinearTransferFunction.ispc
volume/transferFunction/TransferFunctionDispatch.ispc render/ao/volumes.ispc render/scivis/volumes.ispc
render/pathtracer/volumes/VolumeSampler.ispc texture/TextureVolume.ispc
""".split()


def flags(ref, volumes):
    cpu = os.path.join(ref, "modules", "cpu")
    f = ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-x", "c++", "-w",
         # modules/cpu/CMakeLists.txt:455-480
         "-DOSPRAY_TARGET_SYCL", "-DOSPRAY_BEGIN_ISPC_NAMESPACE=namespace ispc{", "-DOSPRAY_END_ISPC_NAMESPACE=}",
         "-DOSPRAY_GLOBAL=static", "-DSYCL_EXTERNAL=", "-Duniform=", "-Dvarying=", "-Dunmasked=", "-Dexport=",
         "-DprogramCount=1", "-DprogramIndex=0", "-DrtcIntersectV=rtcIntersect1", "-DrtcOccludedV=rtcOccluded1",
         "-DrtcInterpolateV1=rtcInterpolate1", "-DvklComputeSampleV=vklComputeSample", "-DvklComputeGradientV=vklComputeGradient",
         "-DvklInitIntervalIteratorV=vklInitIntervalIterator", "-DvklIterateIntervalV=vklIterateInterval",
         "-DvklInitHitIteratorV=vklInitHitIterator", "-DvklIterateHitV=vklIterateHit",
         "-I" + os.path.join(HERE, "refshim"), "-I" + HERE, "-I" + cpu, "-I" + os.path.join(ref, "ospray", "include"),
         "-I" + os.path.join(ref, "ospray"), "-I" + ref]
    if volumes:
        f.append("-DOSPRAY_ENABLE_VOLUMES")
    return f


def compile_one(args):
    cxx, fl, src, obj = args
    r = subprocess.run([cxx] + fl + ["-c", src, "-o", obj], capture_output=True, text=True)
    return src, r.returncode, r.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--no-volumes", action="store_true")
    ap.add_argument("--cxx", default=os.environ.get("CXX", "g++"))
    ap.add_argument("--debug", action="store_true", help="-g for addr2line on a crash backtrace")
    a = ap.parse_args()
    cpu = os.path.join(a.reference, "modules", "cpu")
    if not os.path.isdir(cpu):
        print(f"build_ref: {cpu} not present; keeping any prebuilt oracle/_ref/libospref.so", file=sys.stderr)
        return 0
    volumes = not a.no_volumes
    fl = flags(a.reference, volumes) + (["-g"] if a.debug else [])
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="ospref_build_")
    try:
        jobs = []
        for s in KERNEL_SOURCES + (VOLUME_SOURCES if volumes else []):
            jobs.append((a.cxx, fl, os.path.join(cpu, s), os.path.join(tmp, s.replace("/", "_") + ".o")))
        own = ["ref_driver.cpp", "ref_embree_stub.cpp"] + (["ref_vkl_stub.cpp"] if volumes else [])
        for s in own:
            # the stubs are ordinary C++ (no ISPC keyword macros); the driver is compiled like a kernel TU
            f = fl if s == "ref_driver.cpp" else ["-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-w", "-I" + HERE,
                                                   "-I" + os.path.join(HERE, "refshim")] + (["-DOSPRAY_ENABLE_VOLUMES"] if volumes else [])
            jobs.append((a.cxx, f, os.path.join(HERE, s), os.path.join(tmp, s + ".o")))
        failed = False
        with cf.ThreadPoolExecutor(a.jobs) as ex:
            for src, rc, err in ex.map(compile_one, jobs):
                if rc != 0:
                    failed = True
                    print(f"build_ref: FAILED {src}\n{err[:4000]}", file=sys.stderr)
        if failed:
            return 1
        objs = [j[3] for j in jobs]
        so = os.path.join(out_dir, "libospref.so")
        r = subprocess.run([a.cxx, "-shared", "-o", so] + objs + ["-lpthread", "-Wl,--no-undefined"], capture_output=True, text=True)
        if r.returncode != 0:
            print("build_ref: link failed\n" + r.stderr[:4000], file=sys.stderr)
            return 1
        print(f"build_ref: wrote {so} ({len(objs)} objects, volumes={'on' if volumes else 'off'})")
        return 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
