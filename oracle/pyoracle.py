"""Python front end of the CPU oracle (liboracle.so).  TEST INFRASTRUCTURE ONLY -- see oracle.cpp.

Builds the host-pointer scene description (the POD structs of include/ospcuda.h, mirrored here with
ctypes) from an `ospray_b200.testing.SceneSpec`, following the reference's commit() functions:
  geometry/Mesh.cpp:37-204 (soup index = iota), Spheres.cpp:46-106 (interleave xyz+r),
  geometry/GeometricModel.cpp:21-87, render/materials/OBJ.cpp:30-66 (energy normalisation),
  volume/Volume.cpp:60-222 + VolumetricModel.cpp:29-82, common/World.cpp:70-188,
  render/pathtracer/PathTracerData.cpp:32-82 (uniform light CDF), render/scivis/SciVisData.cpp:17-105,
  render/Renderer.cpp:38-92, PathTracer.cpp:29-51, SciVis.cpp:27-36, AORenderer.cpp:26-35.
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

INVALID_ID = 0xFFFFFFFF
OC_GEOM_TRIANGLES, OC_GEOM_QUADS, OC_GEOM_SPHERES, OC_GEOM_BOXES = 0, 1, 2, 3
OC_GEOMFLAG_NORMAL_FV, OC_GEOMFLAG_COLOR_FV, OC_GEOMFLAG_TEX_FV, OC_GEOMFLAG_HAS_ALPHA, OC_GEOMFLAG_COLOR_VEC4 = 1, 2, 4, 8, 16
OC_MAT_NONE, OC_MAT_FROM_MODEL, OC_MAT_FROM_RENDERER = 0, 1, 2
OC_LIGHT_AMBIENT, OC_LIGHT_DISTANT, OC_LIGHT_QUAD = 0, 2, 5
OC_RENDERER = {"pathtracer": 0, "scivis": 1, "ao": 2, "debug": 3}
OC_DEBUG = {n: i for i, n in enumerate(["testFrame", "rayDir", "eyeLight", "Ng", "Ns", "color", "texCoord", "dPds", "dPdt", "primID", "geomID",
                                        "instID", "backfacing_Ng", "backfacing_Ns", "volume"])}  # ispc::DebugRendererType order
OC_PF_POINT, OC_PF_BOX, OC_PF_LUT = 0, 1, 2


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Affine(C.Structure):
    _fields_ = [("vx", Vec3), ("vy", Vec3), ("vz", Vec3), ("p", Vec3)]


class Data1D(C.Structure):
    _fields_ = [("addr", C.c_void_p), ("byteStride", C.c_int64), ("numItems", C.c_uint32), ("pad", C.c_uint32)]


class Geometry(C.Structure):
    _fields_ = [("type", C.c_uint32), ("numPrims", C.c_uint32), ("flags", C.c_uint32), ("curveBasis", C.c_uint32),
                ("index", Data1D), ("vertex", Data1D), ("normal", Data1D), ("color", Data1D), ("texcoord", Data1D),
                ("sphere", Data1D), ("box", Data1D)]


class Model(C.Structure):
    _fields_ = [("geom", C.POINTER(Geometry)), ("color", Data1D), ("index", Data1D), ("material", C.POINTER(C.c_uint32)),
                ("numMaterials", C.c_uint32), ("materialMode", C.c_uint32), ("invertedNormals", C.c_uint32), ("userID", C.c_uint32)]


class Texture2D(C.Structure):
    _fields_ = [("data", C.c_void_p * 32), ("sizeX", C.c_int32), ("sizeY", C.c_int32), ("maxLevel", C.c_int32), ("format", C.c_uint32),
                ("filter", C.c_uint32), ("wrapS", C.c_uint32), ("wrapT", C.c_uint32), ("hasAlpha", C.c_uint32)]


class TextureParam(C.Structure):
    _fields_ = [("tex", C.POINTER(Texture2D)), ("transformFlags", C.c_uint32), ("scale2f", C.c_float), ("xform2f", C.c_float * 6)]


PRMAP_NAMES = ["baseColor", "edgeColor", "metallic", "diffuse", "specular", "ior", "transmission", "transmissionColor", "transmissionDepth",
               "roughness", "anisotropy", "rotation", "normal", "baseNormal", "coat", "coatIor", "coatColor", "coatThickness", "coatRoughness",
               "coatNormal", "sheen", "sheenColor", "sheenTint", "sheenRoughness", "opacity", "backlight", "thickness", "emissiveColor"]  # OC_PRMAP_*


class PrincipledMaps(C.Structure):  # OcPrincipledMaps
    _fields_ = [("map", TextureParam * len(PRMAP_NAMES)), ("normalRot", C.c_float * 4), ("baseNormalRot", C.c_float * 4),
                ("coatNormalRot", C.c_float * 4), ("normalScale", C.c_float), ("baseNormalScale", C.c_float), ("coatNormalScale", C.c_float)]


CPMAP_NAMES = ["baseColor", "roughness", "normal", "flakeColor", "flakeScale", "flakeDensity", "flakeSpread", "flakeJitter", "flakeRoughness",
               "coat", "coatIor", "coatColor", "coatThickness", "coatRoughness", "coatNormal", "flipflopColor", "flipflopFalloff"]  # OC_CPMAP_*
CARPAINT_DEFAULTS = {"baseColor": (0.8, 0.8, 0.8), "roughness": 0.0, "normal": 1.0, "flakeColor": (0.0, 0.0, 0.0), "flakeScale": 100.0,
                     "flakeDensity": 0.0, "flakeSpread": 0.3, "flakeJitter": 0.75, "flakeRoughness": 0.3, "coat": 1.0, "coatIor": 1.5,
                     "coatColor": (1.0, 1.0, 1.0), "coatThickness": 1.0, "coatRoughness": 0.0, "coatNormal": 1.0,
                     "flipflopColor": (1.0, 1.0, 1.0), "flipflopFalloff": 1.0}  # CarPaint.cpp:27-50


class CarPaint(C.Structure):  # OcCarPaint
    _fields_ = [("baseColor", Vec3), ("roughness", C.c_float), ("useFlakeColor", C.c_uint32), ("flakeColor", Vec3), ("flakeScale", C.c_float),
                ("flakeDensity", C.c_float), ("flakeSpread", C.c_float), ("flakeJitter", C.c_float), ("flakeRoughness", C.c_float),
                ("coat", C.c_float), ("coatIor", C.c_float), ("coatColor", Vec3), ("coatThickness", C.c_float), ("coatRoughness", C.c_float),
                ("flipflopColor", Vec3), ("flipflopFalloff", C.c_float), ("normalScale", C.c_float), ("coatNormalScale", C.c_float),
                ("normalRot", C.c_float * 4), ("coatNormalRot", C.c_float * 4), ("map", TextureParam * len(CPMAP_NAMES))]


class Principled(C.Structure):  # OcPrincipled
    _fields_ = [("baseColor", Vec3), ("edgeColor", Vec3), ("metallic", C.c_float), ("diffuse", C.c_float), ("specular", C.c_float),
                ("specularMetallic", C.c_uint32), ("ior", C.c_float), ("transmission", C.c_float), ("transmissionColor", Vec3),
                ("transmissionDepth", C.c_float), ("roughness", C.c_float), ("anisotropy", C.c_float), ("rotation", C.c_float),
                ("coat", C.c_float), ("coatIor", C.c_float), ("coatColor", Vec3), ("coatThickness", C.c_float),
                ("coatRoughness", C.c_float), ("opacity", C.c_float), ("outsideIor", C.c_float), ("outsideAttenuation", Vec3),
                ("sheen", C.c_float), ("sheenColor", Vec3), ("sheenTint", C.c_float), ("sheenRoughness", C.c_float), ("thin", C.c_uint32),
                ("backlight", C.c_float), ("thickness", C.c_float), ("maps", C.POINTER(PrincipledMaps))]


# parameters of the principled material this build takes (ospray name -> default), Principled.cpp:33-77
PRINCIPLED_DEFAULTS = {
    "baseColor": (0.8, 0.8, 0.8), "edgeColor": (1.0, 1.0, 1.0), "metallic": 0.0, "diffuse": 1.0, "specular": 1.0,
    "specularMetallic": True, "ior": 1.0, "transmission": 0.0, "transmissionColor": (1.0, 1.0, 1.0), "transmissionDepth": 1.0,
    "roughness": 0.0, "anisotropy": 0.0, "rotation": 0.0, "coat": 0.0, "coatIor": 1.5, "coatColor": (1.0, 1.0, 1.0),
    "coatThickness": 1.0, "coatRoughness": 0.0, "opacity": 1.0, "outsideIor": 1.0, "outsideTransmissionColor": (1.0, 1.0, 1.0),
    "outsideTransmissionDepth": 1.0, "emissiveColor": (0.0, 0.0, 0.0), "intensity": 1.0,
    "sheen": 0.0, "sheenColor": (1.0, 1.0, 1.0), "sheenTint": 0.0, "sheenRoughness": 0.2, "thin": False, "backlight": 0.0, "thickness": 1.0,
}


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("d", C.c_float), ("Kd", Vec3), ("Ks", Vec3), ("Ns", C.c_float), ("Tf", Vec3),
                ("eta", C.c_float), ("attenuation", Vec3), ("attenuationScale", C.c_float), ("emission", Vec3),
                ("transparency", C.c_float), ("isEmissive", C.c_uint32), ("microfacetTables", C.c_void_p), ("hasTextures", C.c_uint32), ("cutout", C.c_uint32),
                ("bumpRot", C.c_float * 4), ("mapD", TextureParam), ("mapKd", TextureParam), ("mapKs", TextureParam),
                ("mapNs", TextureParam), ("mapBump", TextureParam), ("principled", Principled), ("carPaint", C.POINTER(CarPaint))]


class BVH(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("prims", C.c_void_p), ("numNodes", C.c_uint32), ("numRecords", C.c_uint32),
                ("bounds", C.c_float * 6)]


class TransferFunction(C.Structure):
    _fields_ = [("color", C.c_void_p), ("opacity", C.c_void_p), ("numColors", C.c_uint32), ("numOpacities", C.c_uint32),
                ("valueLower", C.c_float), ("valueUpper", C.c_float)]


class Volume(C.Structure):
    _fields_ = [("voxels", C.c_void_p), ("dims", C.c_uint32 * 3), ("mcDims", C.c_uint32 * 3), ("mcRange", C.c_void_p),
                ("gridOrigin", Vec3), ("gridSpacing", Vec3), ("bounds", C.c_float * 6), ("tf", TransferFunction),
                ("densityScale", C.c_float), ("anisotropy", C.c_float), ("gradientShadingScale", C.c_float),
                ("userID", C.c_uint32), ("numRanges", C.c_uint32), ("ranges", C.c_float * 16)]


class Group(C.Structure):
    _fields_ = [("models", C.POINTER(Model)), ("numModels", C.c_uint32), ("numVolumes", C.c_uint32),
                ("volumes", C.POINTER(Volume)), ("bvh", BVH)]


class Instance(C.Structure):
    _fields_ = [("xfm", Affine), ("rcp_xfm", Affine), ("group", C.POINTER(Group)), ("userID", C.c_uint32),
                ("isIdentity", C.c_uint32), ("bounds", C.c_float * 6), ("modelAreaPDF", C.POINTER(C.c_float))]


class Light(C.Structure):
    _fields_ = [("type", C.c_uint32), ("isVisible", C.c_uint32), ("radiance", Vec3), ("position", Vec3), ("edge1", Vec3),
                ("edge2", Vec3), ("nnormal", Vec3), ("ppdf", C.c_float), ("fx", Vec3), ("fy", Vec3), ("fz", Vec3),
                ("cosAngle", C.c_float), ("pdf", C.c_float), ("intensity", Vec3), ("radius", C.c_float),
                ("cosAngleScale", C.c_float), ("innerRadius", C.c_float), ("hdriMap", C.c_void_p), ("hdriCdfX", C.c_void_p), ("hdriCdfY", C.c_void_p),
                ("wx", Vec3), ("wy", Vec3), ("wz", Vec3), ("geoModel", C.c_void_p), ("geoInstance", C.c_void_p),
                ("geoPrimIDs", C.c_void_p), ("geoDistribution", C.c_void_p), ("geoNumPrimitives", C.c_uint32), ("geoPdf", C.c_float)]


class World(C.Structure):
    _fields_ = [("instances", C.POINTER(Instance)), ("numInstances", C.c_uint32), ("hasVolumes", C.c_uint32), ("tlas", BVH),
                ("materials", C.POINTER(Material)), ("numMaterials", C.c_uint32), ("numLights", C.c_uint32),
                ("numGeoLights", C.c_uint32), ("shadowOpaque", C.c_uint32), ("lights", C.POINTER(Light)), ("lightsCDF", C.POINTER(C.c_float)), ("scivisLights", C.POINTER(Light)),
                ("numScivisLights", C.c_uint32), ("numScivisLightsVisibleOnly", C.c_uint32), ("aoColorPi", Vec3),
                ("bounds", C.c_float * 6)]


class Camera(C.Structure):
    _fields_ = [("org", Vec3), ("dir_00", Vec3), ("du_size", Vec3), ("dv_up", Vec3), ("direction", Vec3),
                ("nearClip", C.c_float), ("subImage", C.c_float * 4), ("scaledAperture", C.c_float), ("aspect", C.c_float),
                ("imgPlaneSizeY", C.c_float), ("needLensSample", C.c_uint32), ("needTimeSample", C.c_uint32), ("type", C.c_uint32),
                ("stereoMode", C.c_uint32), ("ipdOffset", Vec3)]


class Renderer(C.Structure):
    _fields_ = [("type", C.c_uint32), ("spp", C.c_int32), ("maxDepth", C.c_uint32), ("minContribution", C.c_float),
                ("bgColor", C.c_float * 4), ("pfType", C.c_uint32), ("pfWidth", C.c_float), ("pfLutW", C.c_uint32),
                ("pfLutH", C.c_uint32), ("pfCdfX", C.POINTER(C.c_float)), ("pfCdfY", C.POINTER(C.c_float)),
                ("materials", C.POINTER(C.c_uint32)), ("numMaterials", C.c_uint32), ("rouletteDepth", C.c_int32),
                ("maxScatteringEvents", C.c_int32), ("maxRadiance", C.c_float), ("numLightSamples", C.c_int32),
                ("limitIndirectLightSamples", C.c_uint32), ("backgroundRefraction", C.c_uint32),
                ("shadowsEnabled", C.c_uint32), ("visibleLights", C.c_uint32), ("aoSamples", C.c_int32),
                ("aoRadius", C.c_float), ("aoIntensity", C.c_float), ("volumeSamplingRate", C.c_float),
                ("worldNumLights", C.c_uint32), ("debugMethod", C.c_uint32), ("worldHasVolumes", C.c_uint32),
                ("mipBiasFactor", C.c_float), ("backplate", C.POINTER(Texture2D)), ("maxDepthTexture", C.POINTER(Texture2D)),
                ("hasTextures", C.c_uint32), ("geometryLights", C.c_uint32), ("worldShadowOpaque", C.c_uint32)]


class FrameBuffer(C.Structure):
    _fields_ = [("sizeX", C.c_int32), ("sizeY", C.c_int32), ("frameID", C.c_int32), ("channels", C.c_uint32),
                ("accumulateVariance", C.c_uint32), ("projectedDepth", C.c_uint32), ("targetFrames", C.c_int32),
                ("colorFormat", C.c_uint32), ("color", C.c_void_p), ("variance", C.c_void_p), ("depth", C.c_void_p),
                ("normal", C.c_void_p), ("firstNormal", C.c_void_p), ("albedo", C.c_void_p), ("position", C.c_void_p),
                ("primID", C.c_void_p), ("objID", C.c_void_p), ("instID", C.c_void_p), ("numRanks", C.c_uint32),
                ("rank", C.c_uint32), ("taskError", C.c_void_p), ("errorThreshold", C.c_float),
                ("rayCounters", C.c_void_p), ("control", C.c_void_p)]


class ToneMapper(C.Structure):
    _fields_ = [("exposure", C.c_float), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float), ("d", C.c_float),
                ("acesColor", C.c_uint32)]


class RayHit(C.Structure):
    _fields_ = [("org", C.c_float * 3), ("t0", C.c_float), ("dir", C.c_float * 3), ("t", C.c_float), ("Ng", C.c_float * 3),
                ("u", C.c_float), ("v", C.c_float), ("primID", C.c_uint32), ("geomID", C.c_uint32), ("instID", C.c_uint32)]


RAYHIT_DTYPE = np.dtype([("org", "<f4", 3), ("t0", "<f4"), ("dir", "<f4", 3), ("t", "<f4"), ("Ng", "<f4", 3), ("u", "<f4"),
                         ("v", "<f4"), ("primID", "<u4"), ("geomID", "<u4"), ("instID", "<u4")])
assert RAYHIT_DTYPE.itemsize == C.sizeof(RayHit) == 64


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = C.CDLL(p)
        L.oracle_render_frame.argtypes = [C.POINTER(World), C.POINTER(Camera), C.POINTER(Renderer), C.POINTER(FrameBuffer), C.c_int]
        L.oracle_trace.argtypes = [C.POINTER(World), C.c_void_p, C.c_uint32, C.c_int]
        L.oracle_camera_rays.argtypes = [C.POINTER(Camera), C.POINTER(Renderer), C.POINTER(FrameBuffer), C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_int, C.c_void_p]
        L.oracle_sampler_kat.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_gauss_pixelfilter_lut.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        F3 = C.c_float * 3
        L.oracle_commit_camera.argtypes = [F3, F3, F3, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(Camera)]
        L.oracle_commit_ortho_camera.argtypes = [F3, F3, F3, C.c_float, C.c_float, C.POINTER(Camera)]
        L.oracle_commit_spot_light.argtypes = [F3, C.c_float, C.c_int, F3, F3, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(Light)]
        L.oracle_commit_sphere_light.argtypes = [F3, C.c_float, C.c_int, F3, C.c_float, C.POINTER(Light)]
        L.oracle_commit_light.argtypes = [C.c_int, F3, C.c_float, C.c_int, F3, F3, F3, F3, C.c_float, C.POINTER(Light)]
        L.oracle_fb_accumulate.argtypes = [C.POINTER(FrameBuffer), C.c_void_p, C.c_void_p]
        L.oracle_fb_variance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_fb_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(ToneMapper)]
        L.oracle_fb_color_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.oracle_fb_gather_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_hardware_threads.restype = C.c_int
        L.oracle_texture_texel_bytes.restype = C.c_uint32
        L.oracle_texture_texel_bytes.argtypes = [C.c_uint32]
        L.oracle_texture_mip_level.argtypes = [C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


_ref = None


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libospref.so"))


def ref_lib():
    """oracle/_ref/libospref.so: the REFERENCE's own kernels (modules/cpu/**/*.ispc compiled in their scalar C++ mode by
    oracle/build_ref.py) behind the same entry-point signatures as liboracle.so, prefix ospref_."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libospref.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run `python oracle/build_ref.py` where /root/reference is present")
        L = C.CDLL(p)
        L.ospref_render_frame.argtypes = [C.POINTER(World), C.POINTER(Camera), C.POINTER(Renderer), C.POINTER(FrameBuffer), C.c_int]
        L.ospref_sampler_kat.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ospref_gauss_pixelfilter_lut.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.ospref_hardware_threads.restype = C.c_int
        L.ospref_fb_accumulate.argtypes = [C.POINTER(FrameBuffer), C.c_void_p, C.c_void_p]
        L.ospref_fb_variance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ospref_fb_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(ToneMapper)]
        L.ospref_fb_color_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.ospref_ray_counters.argtypes = [C.c_void_p, C.c_int]
        L.ospref_texture_sample.argtypes = [C.POINTER(TextureParam), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _ref = L
    return _ref


def curve_eval(basis, v, t):
    """float32, operation by operation as oracle/curve_capsule.h curveEval: v = (4, n, 4) control points, t scalar"""
    f = np.float32
    t = f(t)
    s = f(1.0) - t
    if basis == 0:
        return s * v[0] + t * v[1]
    s2, t2 = s * s, t * t
    s2t, t2s, s3, t3 = s2 * t, t2 * s, s2 * s, t2 * t
    if basis == 1:    # Bezier
        b0, b1, b2, b3, c = s3, f(3.0) * s2t, f(3.0) * t2s, t3, f(1.0)
    elif basis == 2:  # B-spline
        b0, b1 = s3, (f(4.0) * s3 + t3) + (f(12.0) * s2t + f(6.0) * t2s)
        b2, b3, c = (f(4.0) * t3 + s3) + (f(12.0) * t2s + f(6.0) * s2t), t3, f(1.0) / f(6.0)
    else:             # Catmull-Rom
        b0, b1, b2, b3, c = -s2t, f(2.0) + t2 * (f(3.0) * t - f(5.0)), f(2.0) + s2 * (f(3.0) * s - f(5.0)), -t2s, f(0.5)
    return c * (((b0 * v[0] + b1 * v[1]) + b2 * v[2]) + b3 * v[3])


def curve_bounds(pos_radius, index, basis, subdiv=8):
    """bounds of the capsule tessellation (end points -+ mean end radius), as the BVH builder reports them"""
    pr = np.asarray(pos_radius, np.float32)
    v = np.stack([pr[index + k] for k in range(4 if basis != 0 else 2)] + ([pr[index]] * 2 if basis == 0 else []))
    lo, hi = np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
    for k in range(subdiv):
        a, b = curve_eval(basis, v, np.float32(k) * np.float32(1.0 / subdiv)), curve_eval(basis, v, np.float32(k + 1) * np.float32(1.0 / subdiv))
        rad = np.maximum(np.float32(0), np.float32(0.5) * (a[:, 3] + b[:, 3]))[:, None]
        lo = np.minimum(lo, (np.minimum(a[:, :3], b[:, :3]) - rad).min(axis=0))
        hi = np.maximum(hi, (np.maximum(a[:, :3], b[:, :3]) + rad).max(axis=0))
    return lo.astype(np.float32), hi.astype(np.float32)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _data1d(arr, keep):
    a = np.ascontiguousarray(arr)
    keep.append(a)
    stride = a.strides[0] if a.ndim > 0 else a.itemsize
    return Data1D(a.ctypes.data, stride, a.shape[0], 0)


def _affine(m):
    m = np.asarray(m, np.float32).reshape(4, 3)
    return Affine(*[Vec3(*[float(x) for x in m[i]]) for i in range(4)])


def _rcp_affine(m):
    """rkcommon rcp(AffineSpace3f): adjoint / determinant, every op rounded to float32 in source order"""
    f = np.float32
    m = np.asarray(m, np.float32).reshape(4, 3)
    vx, vy, vz, p = m

    def cross(a, b):
        return np.array([f(f(a[1] * b[2]) - f(a[2] * b[1])), f(f(a[2] * b[0]) - f(a[0] * b[2])), f(f(a[0] * b[1]) - f(a[1] * b[0]))], f)
    c0, c1, c2 = cross(vy, vz), cross(vz, vx), cross(vx, vy)
    det = f(f(f(vx[0] * c0[0]) + f(vx[1] * c0[1])) + f(vx[2] * c0[2]))
    r = f(f(1.0) / det)
    lx = np.array([c0[0], c1[0], c2[0]], f) * r
    ly = np.array([c0[1], c1[1], c2[1]], f) * r
    lz = np.array([c0[2], c1[2], c2[2]], f) * r
    t = (f(p[0]) * lx).astype(f)
    t = (t + (f(p[1]) * ly).astype(f)).astype(f)
    t = (t + (f(p[2]) * lz).astype(f)).astype(f)
    return np.stack([lx, ly, lz, -t]).astype(f)


def _xfm_bounds(m, lo, hi):
    m = np.asarray(m, np.float32).reshape(4, 3)
    pts = []
    for i in range(8):
        c = np.array([hi[0] if i & 1 else lo[0], hi[1] if i & 2 else lo[1], hi[2] if i & 4 else lo[2]], np.float32)
        pts.append(c[0] * m[0] + c[1] * m[1] + c[2] * m[2] + m[3])
    pts = np.array(pts, np.float32)
    return pts.min(axis=0), pts.max(axis=0)


_TEX_FORMATS = ["RGBA8", "SRGBA", "RGBA32F", "RGB8", "SRGB", "RGB32F", "R8", "R32F", "L8", "RA8", "LA8", "RGBA16", "RGB16", "RA16", "R16",
                "RA32F", "RGBA16F", "RGB16F", "RA16F", "R16F"]  # OSPTextureFormat order (OSPEnums.h)
_TEX_ALPHA = {"RGBA8", "SRGBA", "RA8", "LA8", "RGBA32F", "RGBA16", "RA16", "RA16F", "RA32F", "RGBA16F"}  # Texture2D.cpp:28-32
_TEX_FILTER = {"linear": 0, "nearest": 1}
_TEX_WRAP = {"repeat": 0, "mirror": 1, "clamp": 2}


def make_texture2d(ts, keep, mipmaps=True):
    """ispc::Texture2D of a TextureSpec: level 0 = the image, then the MIP chain of Texture2D::commit (texture/Texture2D.cpp:
    103-141), every level made from the finer one by oracle_texture_mip_level (texture/MipMapGeneration.ispc restated)"""
    L = lib()
    fmt = _TEX_FORMATS.index(ts.format)
    a = np.ascontiguousarray(ts.data)
    h, w = a.shape[:2]
    assert a.nbytes == w * h * L.oracle_texture_texel_bytes(fmt), "texture data does not match its format"
    t = Texture2D()
    levels = [a]
    if mipmaps and ts.mipmaps:
        sx, sy = w, h
        while (sx, sy) != (1, 1):
            dx, dy = max(sx // 2, 1), max(sy // 2, 1)
            dst = np.zeros(dx * dy * L.oracle_texture_texel_bytes(fmt), np.uint8)
            L.oracle_texture_mip_level(fmt, levels[-1].ctypes.data, sx, sy, dst.ctypes.data, dx, dy)
            levels.append(dst)
            sx, sy = dx, dy
    keep += levels
    for l, lv in enumerate(levels):
        t.data[l] = lv.ctypes.data
    t.sizeX, t.sizeY, t.maxLevel, t.format = w, h, len(levels) - 1, fmt
    t.filter, t.wrapS, t.wrapT = _TEX_FILTER[ts.filter], _TEX_WRAP[ts.wrap[0]], _TEX_WRAP[ts.wrap[1]]
    t.hasAlpha = int(ts.format in _TEX_ALPHA)
    keep.append(t)
    return t


def _a2_mul(a, b):
    """affine2f product a * b; an affine map is (l00, l10, l01, l11, px, py) with l.vx = (l00, l10), l.vy = (l01, l11)"""
    f = np.float32
    return (f(a[0] * b[0] + a[2] * b[1]), f(a[1] * b[0] + a[3] * b[1]), f(a[0] * b[2] + a[2] * b[3]), f(a[1] * b[2] + a[3] * b[3]),
            f(f(a[0] * b[4] + a[2] * b[5]) + a[4]), f(f(a[1] * b[4] + a[3] * b[5]) + a[5]))


def texture_param(mp, keep, mipmaps=True):
    """ispc::TextureParam of a MapSpec: Material::getTextureParam + getTextureTransform2f (render/Material.cpp:73-165)"""
    f = np.float32
    tp = TextureParam()
    one = (f(1), f(0), f(0), f(1), f(0), f(0))
    x, have = one, False
    tr = lambda tx, ty: (f(1), f(0), f(0), f(1), f(tx), f(ty))
    if mp is not None:
        if mp.translation is not None:
            x, have = tr(-f(mp.translation[0]), -f(mp.translation[1])), True
        if mp.scale is not None:
            x = _a2_mul(_a2_mul(_a2_mul(x, tr(0.5, 0.5)), (f(1) / f(mp.scale[0]), f(0), f(0), f(1) / f(mp.scale[1]), f(0), f(0))), tr(-0.5, -0.5))
            have = True
        if mp.rotation is not None:
            r = -(f(mp.rotation) * f(np.pi / 180.0))
            sn, cs = f(np.sin(r, dtype=np.float32)), f(np.cos(r, dtype=np.float32))
            x = _a2_mul(_a2_mul(_a2_mul(x, tr(0.5, 0.5)), (cs, sn, -sn, cs, f(0), f(0))), tr(-0.5, -0.5))
            have = True
        if mp.transform is not None:
            t = [f(v) for v in mp.transform]
            x, have = _a2_mul(x, (t[0], t[1], t[2], t[3], f(0), f(0))), True
        tex = make_texture2d(mp.texture, keep, mipmaps)
        tp.tex = C.pointer(tex)
    tp.transformFlags = 1 if have else 0
    for k in range(6):
        tp.xform2f[k] = float(x[k])
    tp.scale2f = float(np.sqrt(f(f(x[0] * x[3]) - f(x[2] * x[1]))))
    return tp


def ref_texture_sample(map_spec, st, footprint, mipmaps=True):
    """the reference's own get4f(TextureParam, dg) (oracle/_ref) for (n, 2) coordinates and n footprints -> (n, 4)"""
    keep = []
    tp = texture_param(map_spec, keep, mipmaps)
    st, fp = np.ascontiguousarray(st, np.float32), np.ascontiguousarray(footprint, np.float32)
    out = np.zeros((len(fp), 4), np.float32)
    ref_lib().ospref_texture_sample(C.byref(tp), st.ctypes.data, fp.ctypes.data, len(fp), out.ctypes.data)
    return out, tp


def ref_albedo_tables():
    """the reference's microfacet albedo tables (oracle/_ref) as one block in ospcuda_microfacet_albedo_tables' order"""
    L = ref_lib()
    L.ospref_albedo_tables.restype = C.c_uint32
    n = L.ospref_albedo_tables(None, 0)
    out = np.zeros(n, np.float32)
    L.ospref_albedo_tables(C.c_void_p(out.ctypes.data), n)
    return out


def linear2f_orthogonal(m):
    """rkcommon LinearSpace2::orthogonal() (un-vendored; restated): rotation part by m <- (m + m^-T) / 2; m = (vx.x, vx.y, vy.x, vy.y)"""
    f = np.float32
    m = [f(v) for v in m]
    mirror = f(1)
    if f(m[0] * m[3]) - f(m[2] * m[1]) < 0:
        m[0], m[1], mirror = -m[0], -m[1], f(-1)
    for _ in range(99):
        ta, tb, tc, td = m[0], m[2], m[1], m[3]
        r = f(1) / f(f(ta * td) - f(tc * tb))
        inv = (f(td * r), f(-tb * r), f(-tc * r), f(ta * r))
        n = [f(f(0.5) * f(m[k] + inv[k])) for k in range(4)]
        d = [f(n[k] - m[k]) for k in range(4)]
        m = n
        if max(f(f(d[0] * d[0]) + f(d[1] * d[1])), f(f(d[2] * d[2]) + f(d[3] * d[3]))) < f(1e-8):
            break
    return [float(m[0] * mirror), float(m[1] * mirror), float(m[2]), float(m[3])]


class Scene:
    """host-pointer OcWorld built from a SceneSpec; keeps every numpy / ctypes buffer alive"""

    def __init__(self, spec):
        self.mipmaps = spec.uses_mipmaps() if hasattr(spec, "uses_mipmaps") else True
        self.keep = []
        self.materials = []
        L = lib()
        insts = []
        wlo, whi = np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
        has_volumes = 0
        for ispec in spec.instances:
            models = []
            glo, ghi = np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
            for ms in ispec.group.models:
                g, lo, hi = self._geometry(ms.geometry)
                glo, ghi = np.minimum(glo, lo), np.maximum(ghi, hi)
                m = Model()
                m.geom = C.pointer(g)
                self.keep.append(g)
                m.materialMode = OC_MAT_NONE
                if ms.material is not None:
                    slot = self._material(ms.material)
                    arr = (C.c_uint32 * 1)(slot)
                    self.keep.append(arr)
                    m.material = C.cast(arr, C.POINTER(C.c_uint32))
                    m.numMaterials = 1
                    m.materialMode = OC_MAT_FROM_MODEL
                if ms.color is not None:
                    m.color = _data1d(np.asarray(ms.color, np.float32), self.keep)
                m.userID = INVALID_ID if ms.id is None else ms.id
                models.append(m)
            vols = []
            for vs in ispec.group.volumes:
                v, lo, hi = self._volume(vs)
                glo, ghi = np.minimum(glo, lo), np.maximum(ghi, hi)
                vols.append(v)
                has_volumes = 1
            grp = Group()
            if models:
                marr = (Model * len(models))(*models)
                self.keep.append(marr)
                grp.models = C.cast(marr, C.POINTER(Model))
                grp.numModels = len(models)
            if vols:
                varr = (Volume * len(vols))(*vols)
                self.keep.append(varr)
                grp.volumes = C.cast(varr, C.POINTER(Volume))
                grp.numVolumes = len(vols)
            self.keep.append(grp)
            inst = Instance()
            ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.float32)
            xfm = ident if ispec.xfm is None else np.asarray(ispec.xfm, np.float32).ravel()
            inst.xfm = _affine(xfm)
            inst.rcp_xfm = _affine(_rcp_affine(xfm))
            inst.group = C.pointer(grp)
            inst.userID = INVALID_ID if ispec.id is None else ispec.id
            inst.isIdentity = int(np.array_equal(xfm, ident))
            lo, hi = (glo, ghi) if inst.isIdentity else _xfm_bounds(xfm, glo, ghi)
            for k in range(3):
                inst.bounds[k], inst.bounds[3 + k] = float(lo[k]), float(hi[k])
            wlo, whi = np.minimum(wlo, lo), np.maximum(whi, hi)
            insts.append(inst)
        w = World()
        if insts:
            iarr = (Instance * len(insts))(*insts)
            self.keep.append(iarr)
            w.instances = C.cast(iarr, C.POINTER(Instance))
            w.numInstances = len(insts)
        w.hasVolumes = has_volumes
        if self.materials:
            marr = (Material * len(self.materials))(*self.materials)
            self.keep.append(marr)
            w.materials = C.cast(marr, C.POINTER(Material))
            w.numMaterials = len(self.materials)
        # lights: PathTracerData (world lights in order, uniform CDF) and SciVisData (ambient first)
        pt, sv, visible_only = [], [], 0
        ao = np.zeros(3, np.float32)
        # world lights, then the lights of the instances' groups moved into world space (World.cpp:120-150, *Light::createSh)
        all_lights = [(ls, None) for ls in spec.lights]
        for inst in spec.instances:
            for ls in getattr(inst.group, "lights", []):
                all_lights.append((ls, None if inst.xfm is None else np.ascontiguousarray(inst.xfm, np.float32)))
        for ls, lxfm in all_lights:
            l = Light()
            p = ls.params
            t = {"ambient": OC_LIGHT_AMBIENT, "distant": OC_LIGHT_DISTANT, "quad": OC_LIGHT_QUAD, "sphere": 4, "spot": 6, "hdri": 3, "cylinder": 1}[ls.type]
            if ls.type == "cylinder":  # lights/CylinderLight.cpp:22-97
                f = np.float32
                p0, p1 = np.asarray(p.get("position0", (0, 0, 0)), f), np.asarray(p.get("position1", (0, 0, 1)), f)
                radius = f(p.get("radius", 1.0))
                ci = (np.asarray(p.get("color", (1, 1, 1)), f) * f(p.get("intensity", 1.0))).astype(f)
                zmax = np.sqrt(np.dot(p1 - p0, p1 - p0), dtype=f)
                area = f(f(f(2.0) * f(np.pi)) * zmax * radius)
                iq = p.get("intensityQuantity", 0)  # OSP_INTENSITY_QUANTITY_RADIANCE = 0, IRRADIANCE = 1, INTENSITY = 2, POWER = 3
                iq = getattr(iq, "value", iq)
                rad = ci if iq == 0 else (f(np.pi) * ci / area).astype(f) if iq == 2 else (ci / f(f(np.pi) * area)).astype(f) if iq == 3 else np.zeros(3, f)
                if zmax <= 0 or radius <= 0 or (iq in (2, 3) and area <= 0):
                    rad = np.zeros(3, f)
                l.type, l.isVisible, l.radiance, l.radius = 1, int(p.get("visible", True)), Vec3(*map(float, rad)), float(radius)
                l.position, l.edge1 = Vec3(*map(float, p0)), Vec3(*map(float, p1))
                t = 1
            elif ls.type == "hdri":  # lights/HDRILight.cpp:73-97 (frame from direction / up; the distribution is built by the checker's own code)
                f = np.float32
                nrm = lambda v: (v / np.sqrt(np.dot(v, v), dtype=np.float32)).astype(np.float32)
                d, up = np.asarray(p.get("direction", (0, 0, 1)), f), np.asarray(p.get("up", (0, 1, 0)), f)
                vx = nrm(-d); vy = nrm(np.cross(vx, up).astype(f)); vz = np.cross(vx, vy).astype(f)
                scale = (np.asarray(p.get("color", (1, 1, 1)), f) * f(p.get("intensity", 1.0))).astype(f)
                l.type, l.isVisible, l.radiance = 3, int(p.get("visible", True)), Vec3(*map(float, scale))
                l.fx, l.fy, l.fz = Vec3(*map(float, vx)), Vec3(*map(float, vy)), Vec3(*map(float, vz))
                tex = make_texture2d(p["map"], self.keep, self.mipmaps)
                self.keep.append(tex)
                l.hdriMap = C.cast(C.pointer(tex), C.c_void_p)
            elif ls.type == "spot":
                L.oracle_commit_spot_light(_f3(p.get("color", (1, 1, 1))), float(p.get("intensity", 1.0)), int(p.get("visible", True)),
                                           _f3(p.get("position", (0, 0, 0))), _f3(p.get("direction", (0, 0, 1))),
                                           float(p.get("openingAngle", 180.0)), float(p.get("penumbraAngle", 5.0)),
                                           float(p.get("radius", 0.0)), float(p.get("innerRadius", 0.0)), C.byref(l))
            elif ls.type == "sphere":
                L.oracle_commit_sphere_light(_f3(p.get("color", (1, 1, 1))), float(p.get("intensity", 1.0)), int(p.get("visible", True)),
                                             _f3(p.get("position", (0, 0, 0))), float(p.get("radius", 0.0)), C.byref(l))
            else:
                L.oracle_commit_light(t, _f3(p.get("color", (1, 1, 1))), float(p.get("intensity", 1.0)), int(p.get("visible", True)),
                                  _f3(p.get("position", (0, 0, 0))), _f3(p.get("edge1", (1, 0, 0))), _f3(p.get("edge2", (0, 1, 0))),
                                  _f3(p.get("direction", (0, 0, 1))), float(p.get("angularDiameter", 0.0)), C.byref(l))
            if lxfm is not None:
                L.oracle_transform_light(C.byref(l), lxfm.ctypes.data_as(C.POINTER(C.c_float)))
            pt.append(l)
            if t == OC_LIGHT_AMBIENT or t == 3:  # ambient and hdri: visible only under scivis (SciVisData.cpp:17-52)
                if visible_only == len(sv):
                    sv.append(l)
                else:
                    sv.append(sv[visible_only])
                    sv[visible_only] = l
                visible_only += 1
                if t == OC_LIGHT_AMBIENT:
                    ao += np.array([l.radiance.x, l.radiance.y, l.radiance.z], np.float32)
            else:
                sv.append(l)
        if pt:
            parr, sarr = (Light * len(pt))(*pt), (Light * len(sv))(*sv)
            n = len(pt)
            cdf = np.cumsum(np.ones(n, np.float32), dtype=np.float32)
            total = cdf[-1]
            cdf = np.where(cdf >= total, np.float32(1.0000001192092896), cdf * (np.float32(1.0) / total)).astype(np.float32)
            self.keep += [parr, sarr, cdf]
            w.lights = C.cast(parr, C.POINTER(Light))
            w.scivisLights = C.cast(sarr, C.POINTER(Light))
            w.lightsCDF = cdf.ctypes.data_as(C.POINTER(C.c_float))
        w.numLights = len(pt)
        w.numScivisLights = len(sv)
        w.numScivisLightsVisibleOnly = visible_only
        aopi = (ao * np.float32(np.pi)).astype(np.float32)
        w.aoColorPi = Vec3(float(aopi[0]), float(aopi[1]), float(aopi[2]))
        for k in range(3):
            w.bounds[k], w.bounds[3 + k] = float(wlo[k]), float(whi[k])
        self.world = w
        self.bounds = (wlo, whi)

    def _material(self, ms):
        m = Material()
        for k in range(4):
            m.bumpRot[k] = (1.0, 0.0, 0.0, 1.0)[k]
        for f in ("mapD", "mapKd", "mapKs", "mapNs", "mapBump"):
            setattr(m, f, texture_param(None, self.keep))
        maps = {"mapD": ms.map_d, "mapKd": ms.map_kd if ms.type == "obj" else ms.map_attenuation_color, "mapKs": ms.map_ks,
                "mapNs": ms.map_ns, "mapBump": ms.map_bump}
        for f, mp in maps.items():
            if mp is not None:
                setattr(m, f, texture_param(mp, self.keep, self.mipmaps))
                m.hasTextures = 1
        if ms.map_bump is not None:  # OBJ.cpp:43-44: orthogonal().transposed() of the 2D transform's linear part
            rot = linear2f_orthogonal([m.mapBump.xform2f[k] for k in range(4)])
            m.bumpRot[0], m.bumpRot[1], m.bumpRot[2], m.bumpRot[3] = rot[0], rot[2], rot[1], rot[3]
        if ms.type == "obj":
            kd, ks, tf = (np.asarray(x, np.float32) for x in (ms.kd, ms.ks, ms.tf))
            total = np.float32(np.max(kd + ks + tf))
            if total > 1.0:
                kd, ks, tf = kd / total, ks / total, tf / total
            m.type, m.d, m.Ns = 0, float(ms.d), float(ms.ns)
            m.Kd, m.Ks, m.Tf = Vec3(*map(float, kd)), Vec3(*map(float, ks)), Vec3(*map(float, tf))
        elif ms.type == "metal":  # render/materials/Metal.cpp:30-76, RGB form (Kd = eta, Ks = k, Ns = roughness)
            m.type = 4
            m.Kd, m.Ks, m.Ns = Vec3(*map(float, np.asarray(ms.metal_eta, np.float32))), Vec3(*map(float, np.asarray(ms.metal_k, np.float32))), float(np.float32(ms.roughness))
            if ms.map_roughness is not None:
                m.mapNs = texture_param(ms.map_roughness, self.keep, self.mipmaps)
                m.hasTextures = 1
        elif ms.type == "plastic":  # render/materials/Plastic.cpp:28-38
            m.type = 5
            m.Kd, m.eta, m.Ns = Vec3(*map(float, np.asarray(ms.pigment_color, np.float32))), float(np.float32(1.0) / np.float32(ms.eta)), float(np.float32(ms.roughness))
        elif ms.type == "glass":  # render/materials/Glass.cpp:35-63 (inside medium in eta / attenuation, outside one in principled.outside*)
            f = np.float32
            m.type = 7
            dist = max(f(ms.attenuation_distance), f(1e-5))
            m.eta = float(f(ms.eta))
            m.attenuation = Vec3(*map(float, (np.log(np.asarray(ms.attenuation_color, np.float32)).astype(np.float32) / dist)))
            m.principled.outsideIor = float(f(ms.eta_outside))
            m.principled.outsideAttenuation = Vec3(*map(float, (np.log(np.asarray(ms.attenuation_color_outside, np.float32)).astype(np.float32) / dist)))
        elif ms.type == "alloy":  # render/materials/Alloy.cpp:28-40 (Kd = color, Ks = edgeColor, Ns = roughness)
            m.type = 8
            m.Kd, m.Ks, m.Ns = Vec3(*map(float, np.asarray(ms.alloy_color, np.float32))), Vec3(*map(float, np.asarray(ms.edge_color, np.float32))), float(np.float32(ms.roughness))
            for f_, mp in (("mapKd", ms.map_color), ("mapKs", ms.map_edge_color), ("mapNs", ms.map_roughness)):
                if mp is not None:
                    setattr(m, f_, texture_param(mp, self.keep, self.mipmaps))
                    m.hasTextures = 1
        elif ms.type == "carPaint":  # render/materials/CarPaint.cpp:27-87
            f = np.float32
            q = dict(CARPAINT_DEFAULTS)
            given = dict(ms.params)
            map_specs = {k[4:]: given.pop(k) for k in list(given) if k.startswith("map_")}
            assert set(given) <= set(q) and set(map_specs) <= set(CPMAP_NAMES), ms.params
            for name in map_specs:
                q[name] = 1.0 if not isinstance(q[name], tuple) else (1.0, 1.0, 1.0)
            q.update(given)
            cp = CarPaint()
            for i, name in enumerate(CPMAP_NAMES):
                cp.map[i] = texture_param(map_specs.get(name), self.keep, self.mipmaps)
            for fld, name in (("normalRot", "normal"), ("coatNormalRot", "coatNormal")):
                t = cp.map[CPMAP_NAMES.index(name)]
                rot = linear2f_orthogonal([t.xform2f[k] for k in range(4)])
                getattr(cp, fld)[0], getattr(cp, fld)[1], getattr(cp, fld)[2], getattr(cp, fld)[3] = rot[0], rot[2], rot[1], rot[3]
            for name in ("baseColor", "flakeColor", "coatColor", "flipflopColor"):
                setattr(cp, name, Vec3(*map(float, np.asarray(q[name], f))))
            for name in ("roughness", "flakeScale", "flakeDensity", "flakeSpread", "flakeJitter", "flakeRoughness", "coat", "coatIor", "coatThickness",
                         "coatRoughness", "flipflopFalloff"):
                setattr(cp, name, float(f(q[name])))
            cp.normalScale, cp.coatNormalScale = float(f(q["normal"])), float(f(q["coatNormal"]))
            cp.useFlakeColor = int("flakeColor" in given)
            self.keep.append(cp)
            m.type = 11
            m.carPaint = C.pointer(cp)
        elif ms.type == "velvet":  # render/materials/Velvet.cpp:27-39
            f = np.float32
            q = {"reflectance": (0.4, 0.0, 0.0), "backScattering": 0.5, "horizonScatteringColor": (0.75, 0.1, 0.1), "horizonScatteringFallOff": 10.0}
            assert set(ms.params) <= set(q), ms.params
            q.update(ms.params)
            m.type = 9
            m.Kd, m.Ks = Vec3(*map(float, np.asarray(q["reflectance"], f))), Vec3(*map(float, np.asarray(q["horizonScatteringColor"], f)))
            m.Ns, m.eta = float(f(q["backScattering"])), float(f(q["horizonScatteringFallOff"]))
        elif ms.type == "metallicPaint":  # render/materials/MetallicPaint.cpp:27-40
            f = np.float32
            q = {"baseColor": (0.8, 0.8, 0.8), "flakeAmount": 0.3, "flakeColor": (0.909, 0.916, 0.923), "flakeSpread": 0.5, "eta": 1.5}
            given = dict(ms.params)
            mp = given.pop("map_baseColor", None)
            assert set(given) <= set(q), ms.params
            if mp is not None:
                q["baseColor"] = (1.0, 1.0, 1.0)
                m.mapKd = texture_param(mp, self.keep, self.mipmaps)
                m.hasTextures = 1
            q.update(given)
            m.type = 10
            fa = f(q["flakeAmount"])
            m.Kd = Vec3(*map(float, np.asarray(q["baseColor"], f) * (f(1.0) - fa)))
            m.Ks = Vec3(*map(float, np.asarray(q["flakeColor"], f) * fa))
            m.attenuationScale, m.Ns, m.eta = float(fa), float(f(q["flakeSpread"])), float(f(1.0) / f(q["eta"]))
        elif ms.type == "principled":  # render/materials/Principled.cpp:33-150 (solid, no sheen, no maps)
            f = np.float32
            prm = dict(PRINCIPLED_DEFAULTS)
            given = dict(ms.principled)
            map_specs = {k[4:]: given.pop(k) for k in list(given) if k.startswith("map_")}
            scales = {k: float(given.pop(k)) for k in ("normal", "baseNormal", "coatNormal") if k in given}
            unknown = (set(given) - set(prm)) | (set(map_specs) - set(PRMAP_NAMES))
            if unknown:
                raise ValueError(f"principled parameters not supported: {sorted(unknown)}")
            for name in map_specs:  # a factor with a map defaults to 1 (Material.cpp:103-126)
                if name in prm:
                    prm[name] = 1.0 if not isinstance(prm[name], tuple) else (1.0, 1.0, 1.0)
            prm.update(given)
            if map_specs:
                pm = PrincipledMaps()
                for i, name in enumerate(PRMAP_NAMES):
                    pm.map[i] = texture_param(map_specs.get(name), self.keep, self.mipmaps)
                for fld, name in (("normalRot", "normal"), ("baseNormalRot", "baseNormal"), ("coatNormalRot", "coatNormal")):
                    t = pm.map[PRMAP_NAMES.index(name)]
                    rot = linear2f_orthogonal([t.xform2f[k] for k in range(4)])
                    getattr(pm, fld)[0], getattr(pm, fld)[1], getattr(pm, fld)[2], getattr(pm, fld)[3] = rot[0], rot[2], rot[1], rot[3]
                pm.normalScale, pm.baseNormalScale, pm.coatNormalScale = (scales.get(k, 1.0) for k in ("normal", "baseNormal", "coatNormal"))
                self.keep.append(pm)
                m.principled.maps = C.pointer(pm)
            m.type = 6
            q = m.principled
            for name in ("baseColor", "edgeColor", "transmissionColor", "coatColor", "sheenColor"):
                setattr(q, name, Vec3(*map(float, np.asarray(prm[name], np.float32))))
            for name in ("metallic", "diffuse", "specular", "ior", "transmission", "transmissionDepth", "roughness", "anisotropy",
                         "rotation", "coat", "coatIor", "coatThickness", "coatRoughness", "opacity", "sheen", "sheenTint", "sheenRoughness",
                         "backlight", "thickness"):
                setattr(q, name, float(f(prm[name])))
            q.specularMetallic = int(bool(prm["specularMetallic"]))
            q.thin = int(bool(prm["thin"]))
            oi = f(prm["outsideIor"])
            q.outsideIor = float(oi if oi >= 1.0 else f(1.0) / oi)
            otc = np.log(np.asarray(prm["outsideTransmissionColor"], np.float32)).astype(np.float32) / f(prm["outsideTransmissionDepth"])
            q.outsideAttenuation = Vec3(*map(float, otc))
            em = (np.asarray(prm["emissiveColor"], np.float32) * f(prm["intensity"])).astype(np.float32)
            m.emission, m.isEmissive = Vec3(*map(float, em)), int(em.max() > 0)
        elif ms.type == "luminous":  # render/materials/Luminous.cpp:33-45
            m.type = 3
            em = (np.asarray(ms.color, np.float32) * np.float32(ms.intensity)).astype(np.float32)
            m.emission, m.transparency, m.isEmissive = Vec3(*map(float, em)), float(ms.transparency), int(em.max() > 0)
            if ms.map_color is not None:
                m.mapKd = texture_param(ms.map_color, self.keep, self.mipmaps)
                m.hasTextures = 1
        elif ms.type == "thinGlass":
            # render/materials/ThinGlass.cpp:33-48 in float32 (std::log of a float is the float overload)
            f = np.float32
            m.type = 2
            m.eta = float(f(1.0) / f(ms.eta))
            scale = f(ms.thickness) * (f(1.0) / max(f(ms.attenuation_distance), f(1e-5)))
            m.attenuationScale = float(scale)
            att = np.log(np.asarray(ms.attenuation_color, np.float32)).astype(np.float32) * scale
            m.attenuation = Vec3(*map(float, att))
        else:
            m.type = 1
        self.materials.append(m)
        return len(self.materials) - 1

    def _geometry(self, gs):
        g = Geometry()
        if gs.type == "mesh":
            v = np.asarray(gs.vertex, np.float32)
            if gs.index is not None:
                idx = np.asarray(gs.index, np.uint32)
            else:
                e = 4 if gs.quad_soup else 3
                idx = np.arange((len(v) // e) * e, dtype=np.uint32).reshape(-1, e)
            g.type = OC_GEOM_TRIANGLES if idx.shape[1] == 3 else OC_GEOM_QUADS
            g.numPrims = len(idx)
            g.index, g.vertex = _data1d(idx, self.keep), _data1d(v, self.keep)
            flags = OC_GEOMFLAG_TEX_FV  # texcoord absent: the device flags it face-varying by default
            if gs.vertex_normal is not None:
                g.normal = _data1d(np.asarray(gs.vertex_normal, np.float32), self.keep)
            else:
                flags |= OC_GEOMFLAG_NORMAL_FV
            if gs.vertex_color is not None:
                c = np.asarray(gs.vertex_color, np.float32)
                g.color = _data1d(c, self.keep)
                if c.shape[1] == 4:
                    flags |= OC_GEOMFLAG_HAS_ALPHA | OC_GEOMFLAG_COLOR_VEC4
            else:
                flags |= OC_GEOMFLAG_COLOR_FV
            if getattr(gs, "vertex_texcoord", None) is not None:
                g.texcoord = _data1d(np.ascontiguousarray(gs.vertex_texcoord, np.float32), self.keep)
                flags &= ~OC_GEOMFLAG_TEX_FV
            g.flags = flags
            used = v[idx.ravel()]
            return g, used.min(axis=0), used.max(axis=0)
        if gs.type == "sphere":
            pos = np.asarray(gs.sphere_position, np.float32)
            rad = np.asarray(gs.sphere_radius, np.float32) if gs.sphere_radius is not None else np.full(len(pos), gs.radius, np.float32)
            inter = np.concatenate([pos, rad[:, None]], axis=1).astype(np.float32)
            g.type, g.numPrims = OC_GEOM_SPHERES, len(pos)
            g.sphere = _data1d(inter, self.keep)
            return g, (pos - rad[:, None]).min(axis=0), (pos + rad[:, None]).max(axis=0)
        if gs.type == "curve":
            pr = np.ascontiguousarray(gs.curve_vertex, np.float32)
            idx = np.ascontiguousarray(gs.curve_index, np.uint32).reshape(-1, 1)
            g.type, g.numPrims, g.curveBasis = 4, len(idx), {"linear": 0, "bezier": 1, "bspline": 2, "catmull-rom": 4}[gs.curve_basis]
            g.sphere, g.index = _data1d(pr, self.keep), _data1d(idx, self.keep)
            if gs.vertex_color is not None:
                g.color = _data1d(np.asarray(gs.vertex_color, np.float32), self.keep)
                g.flags = OC_GEOMFLAG_HAS_ALPHA | OC_GEOMFLAG_COLOR_VEC4
            lo, hi = curve_bounds(pr, idx.ravel(), g.curveBasis)
            return g, lo, hi
        b = np.asarray(gs.boxes, np.float32)
        g.type, g.numPrims = OC_GEOM_BOXES, len(b)
        g.box = _data1d(b, self.keep)
        return g, b[:, :3].min(axis=0), b[:, 3:].max(axis=0)

    def _volume(self, vs):
        v = Volume()
        vox = np.ascontiguousarray(vs.voxels, np.float32)
        self.keep.append(vox)
        v.voxels = vox.ctypes.data
        dz, dy, dx = vox.shape
        v.dims[0], v.dims[1], v.dims[2] = dx, dy, dz
        go, gs_ = np.asarray(vs.grid_origin, np.float32), np.asarray(vs.grid_spacing, np.float32)
        v.gridOrigin, v.gridSpacing = Vec3(*map(float, go)), Vec3(*map(float, gs_))
        hi = (go + np.array([dx - 1, dy - 1, dz - 1], np.float32) * gs_).astype(np.float32)
        for k in range(3):
            v.bounds[k], v.bounds[3 + k] = float(go[k]), float(hi[k])
        col, opa = np.ascontiguousarray(vs.tf_colors, np.float32), np.ascontiguousarray(vs.tf_opacities, np.float32)
        self.keep += [col, opa]
        v.tf = TransferFunction(col.ctypes.data, opa.ctypes.data, len(col), len(opa), float(vs.value_range[0]), float(vs.value_range[1]))
        v.densityScale, v.anisotropy, v.gradientShadingScale = float(vs.density_scale), float(getattr(vs, "anisotropy", 0.0)), float(vs.gradient_shading_scale)
        v.userID = INVALID_ID
        v.numRanges = 1
        v.ranges[0], v.ranges[1] = -np.inf, np.inf
        return v, go, hi


def make_camera(position, direction=(0, 0, 1), up=(0, 1, 0), fovy=60.0, aspect=1.0, aperture_radius=0.0, focus_distance=1.0,
                kind="perspective", height=1.0, stereo_mode=0, interpupillary_distance=0.0635):
    cam = Camera()
    if kind == "panoramic":
        L = lib()
        L.oracle_commit_pano_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        a, b, c = (np.ascontiguousarray(v, np.float32) for v in (position, direction, up))
        L.oracle_commit_pano_camera(a.ctypes.data, b.ctypes.data, c.ctypes.data, int(stereo_mode), float(interpupillary_distance),
                                    C.cast(C.pointer(cam), C.c_void_p))
        return cam
    if kind == "orthographic":
        lib().oracle_commit_ortho_camera(_f3(position), _f3(direction), _f3(up), float(height), float(np.float32(aspect)), C.byref(cam))
        return cam
    L = lib()
    L.oracle_commit_camera_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
    p3 = lambda v: np.ascontiguousarray(v, np.float32)
    a, b, c = p3(position), p3(direction), p3(up)
    L.oracle_commit_camera_stereo(a.ctypes.data, b.ctypes.data, c.ctypes.data, fovy, float(np.float32(aspect)), aperture_radius, focus_distance,
                                  int(stereo_mode), float(interpupillary_distance), C.cast(C.pointer(cam), C.c_void_p))
    return cam


class RendererParams:
    """OcRenderer with the reference defaults (Renderer.cpp:38-92 etc.); keeps the pixel-filter LUT alive"""

    def __init__(self, kind, pixelSamples=1, pixelFilter="gauss", maxPathLength=20, minContribution=0.001, backgroundColor=(0, 0, 0, 0),
                 roulettePathLength=5, maxScatteringEvents=20, maxContribution=np.inf, lightSamples=-1, limitIndirectLightSamples=True,
                 backgroundRefraction=False, shadows=False, visibleLights=False, aoSamples=None, aoDistance=1e20, aoIntensity=1.0,
                 volumeSamplingRate=1.0, method="eyeLight", mipMapBias=0.0, backplate=None, maxDepthTexture=None, mipmaps=True,
                 geometryLights=True):
        r = Renderer()
        r.geometryLights = int(geometryLights)
        self.keep = []
        r.mipBiasFactor = float(np.power(np.float32(2.0), np.float32(mipMapBias)))  # Renderer.cpp:41,91
        if backplate is not None:
            r.backplate = C.pointer(make_texture2d(backplate, self.keep, mipmaps))
        if maxDepthTexture is not None:
            r.maxDepthTexture = C.pointer(make_texture2d(maxDepthTexture, self.keep, mipmaps))
        r.debugMethod = OC_DEBUG.get(method, 0)  # DebugRenderer.cpp:18-52: unknown names fall back to the test frame
        r.type, r.spp, r.maxDepth, r.minContribution = OC_RENDERER[kind], max(1, pixelSamples), maxPathLength, minContribution
        bg = list(backgroundColor) + [0.0] * (4 - len(backgroundColor))
        for k in range(4):
            r.bgColor[k] = float(bg[k])
        if pixelFilter == "point":
            r.pfType, r.pfWidth = OC_PF_POINT, 0.0
        elif pixelFilter == "box":
            r.pfType, r.pfWidth = OC_PF_BOX, 1.0
        else:
            lut = 15
            self.cdfx, self.cdfy = np.zeros(lut * lut, np.float32), np.zeros(lut, np.float32)
            lib().oracle_gauss_pixelfilter_lut(3.0, lut, self.cdfx.ctypes.data, self.cdfy.ctypes.data)
            r.pfType, r.pfWidth, r.pfLutW, r.pfLutH = OC_PF_LUT, 3.0, lut, lut
            r.pfCdfX = self.cdfx.ctypes.data_as(C.POINTER(C.c_float))
            r.pfCdfY = self.cdfy.ctypes.data_as(C.POINTER(C.c_float))
        r.rouletteDepth, r.maxScatteringEvents, r.maxRadiance = roulettePathLength, maxScatteringEvents, maxContribution
        r.numLightSamples, r.limitIndirectLightSamples, r.backgroundRefraction = lightSamples, int(limitIndirectLightSamples), int(backgroundRefraction)
        r.shadowsEnabled, r.visibleLights = int(shadows), int(visibleLights)
        r.aoSamples = (1 if kind == "ao" else 0) if aoSamples is None else aoSamples
        r.aoRadius, r.aoIntensity, r.volumeSamplingRate = aoDistance, aoIntensity, volumeSamplingRate
        self.sh = r


class HostFrameBuffer:
    """numpy-backed OcFrameBuffer (fb/LocalFB.cpp:49-180: buffers per requested channel)"""

    def __init__(self, width, height, channels=("color", "depth"), accumulate=True, variance=False, num_ranks=1, rank=0):
        self.w, self.h = width, height
        n = width * height
        fb = FrameBuffer()
        fb.sizeX, fb.sizeY, fb.frameID = width, height, -1
        fb.targetFrames = 0 if accumulate else 1
        self.accumulate = accumulate
        self.buf = {}
        spec = {"color": ("color", np.float32, 4), "variance": ("variance", np.float32, 4), "depth": ("depth", np.float32, 1),
                "normal": ("normal", np.float32, 3), "firstNormal": ("firstNormal", np.float32, 3), "albedo": ("albedo", np.float32, 3),
                "position": ("position", np.float32, 3), "primID": ("primID", np.uint32, 1), "objID": ("objID", np.uint32, 1),
                "instID": ("instID", np.uint32, 1)}
        chans = list(channels) + (["variance"] if variance else [])
        for cname in chans:
            field, dt, comps = spec[cname]
            a = np.zeros((height, width, comps) if comps > 1 else (height, width), dt)
            if cname == "depth":
                a[:] = np.inf
            self.buf[cname] = a
            setattr(fb, field, a.ctypes.data)
        fb.accumulateVariance = 0
        self.variance = variance
        fb.numRanks, fb.rank = num_ranks, rank
        self.counters = np.zeros(3, np.uint64)
        fb.rayCounters = self.counters.ctypes.data
        self.sh = fb

    def set_adaptive(self, task_error, threshold):
        """LocalFrameBuffer::getRenderTaskIDs (fb/LocalFB.cpp:238-261): render only the 8x8 tasks whose error is above the
        threshold; task_error None switches it off"""
        self.task_error = None if task_error is None else np.ascontiguousarray(task_error, np.float32)
        self.sh.taskError = None if task_error is None else self.task_error.ctypes.data
        self.sh.errorThreshold = float(threshold)

    def begin_frame(self):
        """FrameBuffer::beginFrame (fb/FrameBuffer.cpp:115-137); variance accumulates on odd frames"""
        self.sh.frameID = self.sh.frameID + 1 if self.accumulate else 0
        self.sh.accumulateVariance = int(self.variance and (self.sh.frameID & 1) == 1)
        self.counters[:] = 0


def render_frame(scene, camera, renderer, fb, threads=0, backend="oracle"):
    """backend "oracle": the restatement (liboracle.so); "ref": the reference's own kernels (oracle/_ref/libospref.so)"""
    fb.begin_frame()
    if backend == "ref":
        rc = ref_lib().ospref_render_frame(C.byref(scene.world), C.byref(camera), C.byref(renderer.sh), C.byref(fb.sh), threads)
        if rc != 0:
            raise RuntimeError("ospref_render_frame: image size must be a multiple of the 8x8 render task")
    else:
        lib().oracle_render_frame(C.byref(scene.world), C.byref(camera), C.byref(renderer.sh), C.byref(fb.sh), threads)
    return fb


def trace(scene, rays, occlusion=False):
    """rays: numpy structured array of RAYHIT_DTYPE (modified in place)"""
    assert rays.dtype == RAYHIT_DTYPE and rays.flags.c_contiguous
    lib().oracle_trace(C.byref(scene.world), rays.ctypes.data, len(rays), int(occlusion))
    return rays


def camera_rays(camera, renderer, fb, px, py, sample_index=0):
    px, py = np.ascontiguousarray(px, np.uint32), np.ascontiguousarray(py, np.uint32)
    out = np.zeros(len(px), RAYHIT_DTYPE)
    lib().oracle_camera_rays(C.byref(camera), C.byref(renderer.sh), C.byref(fb.sh), px.ctypes.data, py.ctypes.data, len(px), sample_index,
                             out.ctypes.data)
    return out


def sampler_kat(kind, seed, first, count, backend="oracle"):
    width = {0: 12, 1: 4, 2: 4, 3: 2, 4: 1}[kind]
    out = np.zeros((count, width), np.float32)
    (ref_lib().ospref_sampler_kat if backend == "ref" else lib().oracle_sampler_kat)(kind, seed, first, count, out.ctypes.data)
    return out.view(np.uint32) if kind == 4 else out


def color_convert(img_f32, fmt, backend="oracle"):
    a = np.ascontiguousarray(img_f32, np.float32).reshape(-1, 4)
    out = np.zeros((len(a), 4), np.float32) if fmt == 3 else np.zeros(len(a), np.uint32)
    (ref_lib().ospref_fb_color_convert if backend == "ref" else lib().oracle_fb_color_convert)(a.ctypes.data, out.ctypes.data, len(a), fmt)
    return out


def tonemap(img_f32, tm, backend="oracle"):
    a = np.ascontiguousarray(img_f32, np.float32).reshape(-1, 4)
    out = np.zeros_like(a)
    (ref_lib().ospref_fb_tonemap if backend == "ref" else lib().oracle_fb_tonemap)(a.ctypes.data, out.ctypes.data, len(a), C.byref(tm))
    return out


def variance(color, var, task=8, backend="oracle"):
    h, w = color.shape[:2]
    tx, ty = (w + task - 1) // task, (h + task - 1) // task
    out = np.zeros(tx * ty, np.float32)
    c, v = np.ascontiguousarray(color, np.float32), np.ascontiguousarray(var, np.float32)
    (ref_lib().ospref_fb_variance if backend == "ref" else lib().oracle_fb_variance)(c.ctypes.data, v.ctypes.data, w, h, task, out.ctypes.data)
    return out


def accumulate(fb, sample_color, sample_depth=None, backend="oracle"):
    """one LocalFB_accumulateSample pass of an externally produced sample image into fb (frameID advanced first)"""
    fb.begin_frame()
    sc = np.ascontiguousarray(sample_color, np.float32)
    sz = None if sample_depth is None else np.ascontiguousarray(sample_depth, np.float32)
    fn = ref_lib().ospref_fb_accumulate if backend == "ref" else lib().oracle_fb_accumulate
    fn(C.byref(fb.sh), sc.ctypes.data, None if sz is None else sz.ctypes.data)
    return fb


def gauss_pixelfilter_lut(width=3.0, lut=15, backend="oracle"):
    cx, cy = np.zeros(lut * lut, np.float32), np.zeros(lut, np.float32)
    (ref_lib().ospref_gauss_pixelfilter_lut if backend == "ref" else lib().oracle_gauss_pixelfilter_lut)(width, lut, cx.ctypes.data, cy.ctypes.data)
    return cx, cy


def default_tonemapper(exposure=1.0, contrast=1.6773, shoulder=0.9714, midIn=0.18, midOut=0.18, hdrMax=11.0785, acesColor=True):
    """ToneMapper::commit coefficients (fb/frame_ops/ToneMapper.cpp:14-51) in float32"""
    f = np.float32
    a, d, m, n, w = f(max(contrast, 0.0001)), f(min(max(shoulder, 0.0001), 1)), f(min(max(midIn, 0.0001), 1)), f(min(max(midOut, 0.0001), 1)), f(max(hdrMax, 1))
    p = lambda x, y: f(np.power(f(x), f(y)))
    b = -((p(m, -a * d) * (-p(m, a) + (n * (p(m, a * d) * n * p(w, a) - p(m, a) * p(w, a * d))) / (p(m, a * d) * n - n * p(w, a * d)))) / n)
    c = max((p(m, a * d) * n * p(w, a) - p(m, a) * p(w, a * d)) / (p(m, a * d) * n - n * p(w, a * d)), f(0))
    return ToneMapper(float(exposure), float(a), float(b), float(c), float(d), int(acesColor))
