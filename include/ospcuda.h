/* ospcuda.h — the thin C-ABI between the host-side `cuda` OSPRay device (C++) and the
 * hand-written sm_100a kernels (libospcuda.so).
 *
 * Plain C: pointers, sizes and POD structs only; no CUDA, torch or C++ types appear in
 * any signature, so the host module never includes a CUDA header.
 * Every entry point names the reference interface it stands in for (paths relative to
 * the OSPRay v3.3.0 tree):
 *
 *   ospcuda_bvh_build_group / ospcuda_bvh_build_tlas
 *        -> Embree rtcCommitScene as called from modules/cpu/common/Group.cpp:129-142
 *           and modules/cpu/common/World.cpp:167-171
 *   ospcuda_trace
 *        -> rtcIntersectV / rtcOccludedV as called from modules/cpu/common/World.ih:29-52,
 *           :311-335 (closest hit / occlusion for arbitrary ray batches; used by tests)
 *   ospcuda_render_frame
 *        -> devicert::Device::launchRendererKernel (ospray/common/DeviceRT.h:140-197) with
 *           PathTracer_renderTask (modules/cpu/render/pathtracer/PathTracer.ispc:218-249),
 *           SciVis_renderTask (render/scivis/SciVis.ispc:287-298) or AORenderer_renderTask
 *           (render/ao/AORenderer.ispc:155-166), plus LocalFB_accumulateSample
 *           (modules/cpu/fb/LocalFB.ispc:14-78) fused into the epilogue
 *   ospcuda_fb_variance      -> Variance_kernel   (fb/frame_ops/Variance.ispc:11-59)
 *   ospcuda_fb_tonemap       -> ToneMapper_kernel (fb/frame_ops/ToneMapper.ispc:12-41)
 *   ospcuda_fb_color_convert -> ColorConversion_kernel (fb/frame_ops/ColorConversion.ispc:11-36)
 *   ospcuda_fb_accumulate    -> LocalFB_accumulateSample for an externally produced sample
 *                               image (used by tests and by the roofline bench leg)
 *   ospcuda_fb_gather_tiles  -> LocalFrameBuffer::writeTiles (fb/LocalFB.cpp:286-353), the
 *                               multi-device untile step after the NCCL gather
 *   ospcuda_volume_macrocells-> Open VKL structuredRegular acceleration build done inside
 *                               vklCommit (modules/cpu/volume/Volume.cpp:77-94)
 *   ospcuda_malloc/free/memcpy/... -> devicert::Device::malloc/free/memcpy (DeviceRT.h:150-170)
 *
 * All functions return 0 on success, non-zero CUDA error code otherwise;
 * ospcuda_last_error() returns the message. Nothing here falls back to the CPU.        */
#ifndef OSPCUDA_H
#define OSPCUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSPCUDA_API __attribute__((visibility("default")))

/* ---- POD views --------------------------------------------------------------------- */
typedef struct { float x, y, z; } OcVec3f;
typedef struct { float x, y, z, w; } OcVec4f;
typedef struct { OcVec3f vx, vy, vz, p; } OcAffine3f; /* column-major 3x4 (OSP_AFFINE3F) */

/* strided 1-D array living in device memory (reference: ispc::Data1D, common/Data.h:15-21) */
typedef struct {
  const uint8_t *addr;
  int64_t byteStride;
  uint32_t numItems;
  uint32_t pad;
} OcData1D;

enum { OC_GEOM_TRIANGLES = 0, OC_GEOM_QUADS = 1, OC_GEOM_SPHERES = 2, OC_GEOM_BOXES = 3, OC_GEOM_CURVES = 4 };
/* round curves are tessellated into OC_CURVE_SUBDIV capsules per segment at BVH build */
#define OC_CURVE_SUBDIV 8u
enum { OC_CURVE_BASIS_LINEAR = 0, OC_CURVE_BASIS_BEZIER = 1, OC_CURVE_BASIS_BSPLINE = 2, OC_CURVE_BASIS_HERMITE = 3, OC_CURVE_BASIS_CATMULL_ROM = 4 }; /* = OSPCurveBasis */
enum { OC_GEOMFLAG_NORMAL_FACEVARYING = 1, OC_GEOMFLAG_COLOR_FACEVARYING = 2,
       OC_GEOMFLAG_TEXCOORD_FACEVARYING = 4, OC_GEOMFLAG_HAS_ALPHA = 8, OC_GEOMFLAG_COLOR_IS_VEC4 = 16,
       OC_GEOMFLAG_ALPHA_ALL_ONE = 32 /* every vec4f vertex colour has w == 1.0f (checked on the host at commit) */ };

/* geometry (reference: ispc::Mesh MeshShared.h:12-44, ispc::Spheres, ispc::Boxes) */
typedef struct {
  uint32_t type;
  uint32_t numPrims;
  uint32_t flags;
  uint32_t curveBasis; /* curves: OC_CURVE_BASIS_* (linear and Catmull-Rom are implemented) */
  OcData1D index;    /* vec3ui (triangles), vec4ui (quads) or uint32 first control point per segment (curves) */
  OcData1D vertex;   /* vec3f */
  OcData1D normal;   /* vec3f, optional */
  OcData1D color;    /* vec3f or vec4f, optional */
  OcData1D texcoord; /* vec2f, optional */
  OcData1D sphere;   /* vec4f (x,y,z,r) interleaved at commit (Spheres.cpp:58-96); curves: vertex.position_radius */
  OcData1D box;      /* box3f (6 floats) */
} OcGeometry;

enum { OC_MAT_NONE = 0, OC_MAT_FROM_MODEL = 1, OC_MAT_FROM_RENDERER = 2 };

/* geometric model (reference: ispc::GeometricModel, geometry/GeometricModel.ih:21-76) */
typedef struct {
  const OcGeometry *geom;
  OcData1D color;           /* vec4f per-prim colour, optional */
  OcData1D index;           /* uint8 indirection, optional */
  const uint32_t *material; /* indices into OcWorld.materials (mode FROM_MODEL) or renderer ids */
  uint32_t numMaterials;
  uint32_t materialMode;
  uint32_t invertedNormals;
  uint32_t userID;          /* 0xFFFFFFFF = unset */
} OcModel;

/* 2D texture (reference: ispc::Texture2D, texture/Texture2DShared.h:16-46; Texture2D.cpp:56-146: level 0 is the app's
 * image, levels 1..maxLevel the MIP chain down to 1x1, all in `format`); format / filter / wrap are the OSPTextureFormat /
 * OSPTextureFilter / OSPTextureWrapMode values of OSPEnums.h */
#define OC_MAX_MIPMAP_LEVEL 32
typedef struct {
  const void *data[OC_MAX_MIPMAP_LEVEL];
  int32_t sizeX, sizeY;
  int32_t maxLevel;
  uint32_t format, filter;
  uint32_t wrapS, wrapT;
  uint32_t hasAlpha;
} OcTexture2D;
/* texture + the optional 2D transform of its coordinates (reference: ispc::TextureParam, texture/TextureParamShared.h:21-38,
 * Material::getTextureParam, render/Material.cpp:73-108): xform2f = (l.vx, l.vy, p), scale2f = sqrt(det l) */
enum { OC_TEX_TRANSFORM_NONE = 0, OC_TEX_TRANSFORM_2D = 1 };
typedef struct {
  const OcTexture2D *tex;      /* NULL: no texture */
  uint32_t transformFlags;
  float scale2f;
  float xform2f[6];
} OcTextureParam;

enum { OC_MATERIAL_OBJ = 0, OC_MATERIAL_OTHER = 1, OC_MATERIAL_THINGLASS = 2, OC_MATERIAL_LUMINOUS = 3, OC_MATERIAL_METAL = 4,
       OC_MATERIAL_PLASTIC = 5, OC_MATERIAL_PRINCIPLED = 6, OC_MATERIAL_GLASS = 7, OC_MATERIAL_ALLOY = 8,
       OC_MATERIAL_VELVET = 9, OC_MATERIAL_METALLICPAINT = 10, OC_MATERIAL_CARPAINT = 11 };
/* texture maps of the principled material (PrincipledShared.h: one TextureParam per factor; the three normal maps come with the
 * orthogonal part of their 2D transform, transposed, and a scale: Material::getMaterialParam1f, render/Material.cpp:103-113) */
enum { OC_PRMAP_BASECOLOR = 0, OC_PRMAP_EDGECOLOR, OC_PRMAP_METALLIC, OC_PRMAP_DIFFUSE, OC_PRMAP_SPECULAR, OC_PRMAP_IOR,
       OC_PRMAP_TRANSMISSION, OC_PRMAP_TRANSMISSIONCOLOR, OC_PRMAP_TRANSMISSIONDEPTH, OC_PRMAP_ROUGHNESS, OC_PRMAP_ANISOTROPY,
       OC_PRMAP_ROTATION, OC_PRMAP_NORMAL, OC_PRMAP_BASENORMAL, OC_PRMAP_COAT, OC_PRMAP_COATIOR, OC_PRMAP_COATCOLOR,
       OC_PRMAP_COATTHICKNESS, OC_PRMAP_COATROUGHNESS, OC_PRMAP_COATNORMAL, OC_PRMAP_SHEEN, OC_PRMAP_SHEENCOLOR, OC_PRMAP_SHEENTINT,
       OC_PRMAP_SHEENROUGHNESS, OC_PRMAP_OPACITY, OC_PRMAP_BACKLIGHT, OC_PRMAP_THICKNESS, OC_PRMAP_EMISSION, OC_PRMAP_COUNT };
typedef struct {
  OcTextureParam map[OC_PRMAP_COUNT];
  float normalRot[4], baseNormalRot[4], coatNormalRot[4]; /* (vx, vy) */
  float normalScale, baseNormalScale, coatNormalScale;
} OcPrincipledMaps;
/* carPaint (reference: ispc::CarPaint, render/materials/CarPaintShared.h:14-100, CarPaint.cpp:27-87): factors as committed (a
 * factor with a map defaults to 1), one TextureParam per factor, the two normal maps with rotation and scale */
enum { OC_CPMAP_BASECOLOR = 0, OC_CPMAP_ROUGHNESS, OC_CPMAP_NORMAL, OC_CPMAP_FLAKECOLOR, OC_CPMAP_FLAKESCALE, OC_CPMAP_FLAKEDENSITY,
       OC_CPMAP_FLAKESPREAD, OC_CPMAP_FLAKEJITTER, OC_CPMAP_FLAKEROUGHNESS, OC_CPMAP_COAT, OC_CPMAP_COATIOR, OC_CPMAP_COATCOLOR,
       OC_CPMAP_COATTHICKNESS, OC_CPMAP_COATROUGHNESS, OC_CPMAP_COATNORMAL, OC_CPMAP_FLIPFLOPCOLOR, OC_CPMAP_FLIPFLOPFALLOFF,
       OC_CPMAP_COUNT };
typedef struct {
  OcVec3f baseColor;
  float roughness;
  uint32_t useFlakeColor;      /* 0: the flakes are aluminium (CarPaint.ispc:86-92) */
  OcVec3f flakeColor;
  float flakeScale, flakeDensity, flakeSpread, flakeJitter, flakeRoughness;
  float coat, coatIor;
  OcVec3f coatColor;
  float coatThickness, coatRoughness;
  OcVec3f flipflopColor;
  float flipflopFalloff;
  float normalScale, coatNormalScale;
  float normalRot[4], coatNormalRot[4];
  OcTextureParam map[OC_CPMAP_COUNT];
} OcCarPaint;
/* principled (reference: ispc::Principled, render/materials/PrincipledShared.h:17-170, Principled.cpp:33-150): factors as committed; outsideIor = outsideIor >= 1 ? outsideIor : 1 / outsideIor,
 * outsideAttenuation = log(outsideTransmissionColor) / outsideTransmissionDepth (the medium around the object) */
typedef struct {
  OcVec3f baseColor, edgeColor;
  float metallic, diffuse, specular;
  uint32_t specularMetallic;
  float ior, transmission;
  OcVec3f transmissionColor;
  float transmissionDepth;
  float roughness, anisotropy, rotation;
  float coat, coatIor;
  OcVec3f coatColor;
  float coatThickness, coatRoughness;
  float opacity;
  float outsideIor;
  OcVec3f outsideAttenuation;
  float sheen;
  OcVec3f sheenColor;
  float sheenTint, sheenRoughness;
  uint32_t thin;               /* thin-walled mode: no medium change, ThinDielectric glass, backlight = diffuse transmission */
  float backlight, thickness;
  const OcPrincipledMaps *maps; /* NULL: no texture maps (a factor with a map defaults to 1, Material.cpp:103-126) */
} OcPrincipled;
/* material (reference: ispc::OBJ, render/materials/OBJShared.h, OBJ.cpp:30-66; ispc::ThinGlass,
 * render/materials/ThinGlassShared.h, ThinGlass.cpp:33-48: eta = 1/ior, attenuationScale = thickness /
 * max(attenuationDistance, EPS), attenuation = log(attenuationColor) * attenuationScale) */
typedef struct {
  uint32_t type;
  float d;
  OcVec3f Kd;
  OcVec3f Ks;
  float Ns;
  OcVec3f Tf;
  float eta;
  OcVec3f attenuation;
  float attenuationScale;
  /* luminous (ispc::Luminous, render/materials/LuminousShared.h, Luminous.cpp:33-45): emission = color * intensity
   * (times map_color = mapKd), transparency; isEmissive = any emission (Material::isEmissive) */
  OcVec3f emission;
  float transparency;
  uint32_t isEmissive;
  /* metal (ispc::Metal, render/materials/MetalShared.h, Metal.cpp:30-76; RGB form): Kd = eta, Ks = k, Ns = roughness, mapNs =
   * map_roughness. plastic (ispc::Plastic, PlasticShared.h, Plastic.cpp): Kd = pigmentColor, eta, Ns = roughness.
   * microfacetTables: the device's albedo tables (ospcuda_microfacet_albedo_tables; MicrofacetAlbedoTables.ispc) */
  const float *microfacetTables;
  uint32_t hasTextures;        /* any of the maps below is set (kernels without texture support skip them all) */
  uint32_t cutout;             /* a map can make the surface transmit along shadow rays: map_d, or map_kd with an alpha channel */
  float bumpRot[4];            /* OBJ.cpp:43-44: orthogonal part of map_bump's 2D transform, transposed (vx, vy) */
  /* OBJ: map_d, map_kd, map_ks, map_ns, map_bump (OBJShared.h:19-31); thinGlass: mapKd = map_attenuationColor */
  OcTextureParam mapD, mapKd, mapKs, mapNs, mapBump;
  /* type OC_MATERIAL_PRINCIPLED (emission / isEmissive above: emissiveColor * intensity).
   * glass (ispc::Glass, render/materials/GlassShared.h, Glass.cpp:35-63): eta / attenuation above = the inside medium (ior,
   * log(attenuationColorInside) / max(attenuationDistance, EPS)), principled.outsideIor / outsideAttenuation = the outside one.
   * alloy (ispc::Alloy, AlloyShared.h, Alloy.cpp:28-40): Kd = color, Ks = edgeColor, Ns = roughness, mapKd / mapKs / mapNs their maps.
   * velvet (ispc::Velvet, VelvetShared.h, Velvet.cpp:27-39): Kd = reflectance, Ns = backScattering, Ks = horizonScatteringColor,
   * eta = horizonScatteringFallOff. metallicPaint (ispc::MetallicPaint, MetallicPaint.cpp:27-40): Kd = baseColor * (1 - flakeAmount),
   * mapKd = map_baseColor, Ks = flakeColor * flakeAmount, attenuationScale = flakeAmount, Ns = flakeSpread, eta = 1 / eta */
  OcPrincipled principled;
  const OcCarPaint *carPaint;  /* type OC_MATERIAL_CARPAINT (device memory) */
} OcMaterial;

/* 4-wide BVH node, 128 B = one cache line, fetched as 8 x 128-bit loads */
typedef struct {
  float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
  uint32_t child[4]; /* 0xFFFFFFFF empty; bit31: leaf (bits 27..30 count, bits 0..26 first record) */
  uint32_t pad[4];
} OcNode4;

enum { OC_PRIM_TRI = 0, OC_PRIM_TRI_QUAD2ND = 1, OC_PRIM_SPHERE = 2, OC_PRIM_BOX = 3, OC_PRIM_INSTANCE = 4,
       OC_PRIM_CAPSULE = 5 /* one piece of a round curve: (p0|primID, p1|geomID, (radius, u0, du, kind)) */ };

typedef struct {
  const OcNode4 *nodes;
  const OcVec4f *prims; /* 3 x float4 per record, in leaf order */
  uint32_t numNodes;
  uint32_t numRecords;
  float bounds[6];
} OcBVH;

/* linear transfer function (reference: LinearTransferFunction.ispc:11-43) */
typedef struct {
  const OcVec3f *color;
  const float *opacity;
  uint32_t numColors;
  uint32_t numOpacities;
  float valueLower, valueUpper;
} OcTransferFunction;

/* structuredRegular volume + its volumetric model
 * (reference: volume/Volume.cpp:60-222 -> Open VKL; VolumetricModel.cpp:29-82) */
typedef struct {
  const float *voxels;       /* x fastest */
  uint32_t dims[3];
  uint32_t mcDims[3];        /* macrocell grid, 16^3 cells each (Open VKL GridAccelerator) */
  const float *mcRange;      /* 2 floats (min,max) per macrocell */
  OcVec3f gridOrigin, gridSpacing;
  float bounds[6];
  OcTransferFunction tf;
  float densityScale, anisotropy, gradientShadingScale;
  uint32_t userID;
  uint32_t numRanges;        /* TF positive-opacity value ranges (VolumetricModel.cpp:46-62) */
  float ranges[2 * 8];
} OcVolume;

typedef struct {
  const OcModel *models;
  uint32_t numModels;
  uint32_t numVolumes;
  const OcVolume *volumes;
  OcBVH bvh;
} OcGroup;

/* instance (reference: ispc::Instance, common/InstanceShared.h; Instance.cpp:20-34) */
typedef struct OcInstanceS {
  OcAffine3f xfm, rcp_xfm;
  const OcGroup *group;
  uint32_t userID;
  uint32_t isIdentity;
  float bounds[6]; /* world-space */
  /* 1 / (world-space area of the emissive primitives) per geometric model of the group, NULL when the instance holds no
   * geometry light (what GeometricModel_gatherEmissivePrimIDs leaves in ispc::GeometricModel::areaPDF,
   * geometry/GeometricModel.ispc:48-53; kept per instance here because the area depends on the instance transform) */
  const float *modelAreaPDF;
} OcInstance;

enum { OC_LIGHT_AMBIENT = 0, OC_LIGHT_CYLINDER = 1, OC_LIGHT_DISTANT = 2, OC_LIGHT_HDRI = 3, OC_LIGHT_SPHERE = 4, OC_LIGHT_QUAD = 5, OC_LIGHT_SPOT = 6,
       OC_LIGHT_GEOMETRY = 8 }; /* values of ispc::LightType */
struct OcInstanceS;
/* lights (reference: lights/LightShared.h:79-110, QuadLightShared.h, DirectionalLightShared.h,
 * AmbientLightShared.h) */
typedef struct {
  uint32_t type;
  uint32_t isVisible;
  OcVec3f radiance;            /* ambient/quad: radiance; distant: irradiance */
  OcVec3f position, edge1, edge2, nnormal; /* quad */
  float ppdf;                  /* quad: 1/area */
  OcVec3f fx, fy, fz;          /* distant: frame (vz = direction towards light) */
  float cosAngle, pdf;         /* distant */
  OcVec3f intensity;           /* sphere ("sphere" = ispc::PointLight, lights/PointLightShared.h): radiant intensity (r = 0 path) */
  float radius;                /* sphere: radius; `radiance` is its surface radiance, `position` its centre */
  /* spot (ispc::SpotLight, lights/SpotLightShared.h): position, fx/fy/fz = c0/c90/direction, radiance, intensity,
   * cosAngle = cosAngleMax, ppdf = areaPdf (ring), radius, plus: */
  float cosAngleScale, innerRadius;
  /* geometry light: the proxy the pathtracer makes for every geometric model with emissive primitives
   * (ispc::GeometryLight, render/pathtracer/GeometryLightShared.h; PathTracerData.cpp:87-160): the model and the instance
   * it sits in, the emissive primitives, the CDF of their world-space areas, pdf = 1 / total area */
  /* cylinder (ispc::CylinderLight, lights/CylinderLightShared.h, CylinderLight.cpp:22-97): position = position0, edge1 =
   * position1, radius, radiance (0 when the cylinder is degenerate) */
  /* hdri (ispc::HDRILight, lights/HDRILightShared.h, HDRILight.cpp:20-97): fx / fy / fz = light2world, wx / wy / wz =
   * world2light, radiance = radianceScale, the latitude / longitude map and the CDFs of its importance distribution
   * (ospcuda_hdri_distribution: cdfX = sizeX * sizeY floats, row-wise; cdfY = sizeY floats) */
  const OcTexture2D *hdriMap;
  const float *hdriCdfX, *hdriCdfY;
  OcVec3f wx, wy, wz;
  const OcModel *geoModel;
  const struct OcInstanceS *geoInstance;
  const int32_t *geoPrimIDs;
  const float *geoDistribution;
  uint32_t geoNumPrimitives;
  float geoPdf;
} OcLight;

typedef struct {
  const OcInstance *instances;
  uint32_t numInstances;
  uint32_t hasVolumes;
  OcBVH tlas;                  /* top level over the instance bounds, one instance per leaf record (always built) */
  const OcMaterial *materials;
  uint32_t numMaterials;
  /* pathtracer light list (PathTracerData.cpp:32-82): geometry lights first, world, instance */
  uint32_t numLights;
  uint32_t numGeoLights;
  /* what a shadow ray can run into, decided on the host before a frame (World::prepareFrame): 2 = every surface of the
   * world takes a shadow-ray contribution to exactly 0 (no model colours, no vertex alpha, every material an OBJ with d = 1
   * and tf = 0 or none); 1 = the same provided the contribution is <= 1e6 * minContribution (vec4 vertex colours whose
   * alphas are all 1, see ptOccluderIsOpaque in pathtracer_wf.cuh); 0 = look at the occluder */
  uint32_t shadowOpaque;
  const OcLight *lights;
  const float *lightsCDF;
  /* scivis light list (SciVisData.cpp:74-105): visible-only lights first, ambient folded */
  const OcLight *scivisLights;
  uint32_t numScivisLights;
  uint32_t numScivisLightsVisibleOnly;
  OcVec3f aoColorPi;
  float bounds[6];
} OcWorld;

enum { OC_CAMERA_PERSPECTIVE = 0, OC_CAMERA_ORTHOGRAPHIC = 1, OC_CAMERA_PANORAMIC = 2 };
/* camera (reference: PerspectiveCameraShared.h, PerspectiveCamera.cpp:26-117; OrthographicCameraShared.h,
 * OrthographicCamera.cpp:25-49: for the orthographic type dir_00 is the (normalised) view direction, org the lower
 * left corner of the image plane, du_size / dv_up its extents, imgPlaneSizeY the `height` parameter) */
typedef struct {
  OcVec3f org, dir_00, du_size, dv_up, direction;
  float nearClip;
  float subImage[4];           /* lower.x, lower.y, upper.x, upper.y */
  float scaledAperture, aspect, imgPlaneSizeY;
  uint32_t needLensSample, needTimeSample;
  uint32_t type;
  /* perspective: OSPStereoMode (PerspectiveCamera.cpp:35-100: LEFT / RIGHT are folded into org at commit, SIDE_BY_SIDE /
   * TOP_BOTTOM pick the eye per pixel, PerspectiveCamera.ispc:28-40) and half the interpupillary distance along du */
  uint32_t stereoMode;
  OcVec3f ipdOffset;
} OcCamera;

enum { OC_RENDERER_PATHTRACER = 0, OC_RENDERER_SCIVIS = 1, OC_RENDERER_AO = 2, OC_RENDERER_DEBUG = 3 };
/* debug renderer methods: ispc::DebugRendererType (render/debug/DebugRendererShared.h:12-30), same order */
enum { OC_DEBUG_TEST_FRAME = 0, OC_DEBUG_RAY_DIR, OC_DEBUG_EYE_LIGHT, OC_DEBUG_NG, OC_DEBUG_NS, OC_DEBUG_COLOR, OC_DEBUG_TEX_COORD,
       OC_DEBUG_DPDS, OC_DEBUG_DPDT, OC_DEBUG_PRIM_ID, OC_DEBUG_GEOM_ID, OC_DEBUG_INST_ID, OC_DEBUG_BACKFACING_NG,
       OC_DEBUG_BACKFACING_NS, OC_DEBUG_VOLUME };
enum { OC_PF_POINT = 0, OC_PF_BOX = 1, OC_PF_LUT = 2 };

/* renderer (reference: Renderer.cpp:38-92, PathTracer.cpp:29-51, SciVis.cpp:27-36,
 * AORenderer.cpp:26-35, pf/PixelFilter.cpp:23-69) */
typedef struct {
  uint32_t type;
  int32_t spp;
  uint32_t maxDepth;
  float minContribution;
  float bgColor[4];
  /* pixel filter */
  uint32_t pfType;
  float pfWidth;
  uint32_t pfLutW, pfLutH;
  const float *pfCdfX;         /* pfLutW*pfLutH */
  const float *pfCdfY;         /* pfLutH */
  /* renderer material list (for OC_MAT_FROM_RENDERER) : indices into OcWorld.materials */
  const uint32_t *materials;
  uint32_t numMaterials;
  /* pathtracer */
  int32_t rouletteDepth;
  int32_t maxScatteringEvents;
  float maxRadiance;
  int32_t numLightSamples;
  uint32_t limitIndirectLightSamples;
  uint32_t backgroundRefraction;
  /* scivis / ao */
  uint32_t shadowsEnabled;
  uint32_t visibleLights;
  int32_t aoSamples;
  float aoRadius;
  float aoIntensity;
  float volumeSamplingRate;
  /* host-side copy of OcWorld.numLights of the world this frame renders (set per ospRenderFrame by the
   * device; PathTracerData is per world, PathTracerData.cpp:32-82): sizes the pathtracer's light-sample loop */
  uint32_t worldNumLights;
  /* debug renderer (DebugRenderer.cpp:69-75): OC_DEBUG_* */
  uint32_t debugMethod;
  /* host-side copy of OcWorld.hasVolumes of the world this frame renders (set per ospRenderFrame like worldNumLights):
   * the pathtracer then samples free paths / transmittance and carries a RandomSampler per path */
  uint32_t worldHasVolumes;
  /* Renderer.cpp:41,91: 2^mipMapBias; map_backplate / map_maxDepth (Renderer.cpp:61-80, Renderer.ispc:18-39), NULL = none */
  float mipBiasFactor;
  const OcTexture2D *backplate;
  const OcTexture2D *maxDepthTexture;
  /* host-side: the world / renderer of this frame reference textures or emissive materials (set per ospRenderFrame like
   * worldNumLights): the frame runs the kernels compiled with the extended material support (texture coordinates,
   * tangents, ray cones, emission / geometry lights) */
  uint32_t hasTextures;
  /* pathtracer parameter geometryLights (PathTracer.cpp:44): emissive geometry is sampled as light sources */
  uint32_t geometryLights;
  /* host-side copy of OcWorld.shadowOpaque of the world this frame renders (set per ospRenderFrame) */
  uint32_t worldShadowOpaque;
} OcRenderer;

/* framebuffer (reference: LocalFBShared.h, FrameBufferShared.h; LocalFB.cpp:49-180) */
typedef struct {
  int32_t sizeX, sizeY;
  int32_t frameID;             /* already incremented by beginFrame (FrameBuffer.cpp:115-118) */
  uint32_t channels;
  uint32_t accumulateVariance;
  uint32_t projectedDepth;
  int32_t targetFrames;
  uint32_t colorFormat;        /* OSPFrameBufferFormat */
  OcVec4f *color;
  OcVec4f *variance;
  float *depth;
  OcVec3f *normal, *firstNormal, *albedo, *position;
  uint32_t *primID, *objID, *instID;
  /* image-tile partition across ranks (reference: MultiDevice.cpp:403-416: tile % N == rank) */
  uint32_t numRanks, rank;
  /* adaptive accumulation (reference: LocalFrameBuffer::getRenderTaskIDs, fb/LocalFB.cpp:238-261): when taskError is
   * not NULL only the 8x8 render tasks with taskError[task] > errorThreshold are rendered this frame; task =
   * (y/8) * ceil(sizeX/8) + x/8; the errors are those of the Variance frame op (ospcuda_fb_variance) */
  const float *taskError;
  float errorThreshold;
  /* per-frame ray statistics: [0] primary, [1] continuation, [2] shadow/occlusion segments; [3] = how many of those the
   * pathtracer's tail kernel traced (bench.py: rays of the TRACE launches = total - [3]) */
  unsigned long long *rayCounters;
  /* frame control words in device memory (may be NULL): [0] cancel flag -- every frame kernel looks at it when it starts
   * and the persistent loops when they fetch work, a set flag makes them return (ospCancel; reference:
   * fb/FrameBuffer.cpp:180-187, render/RendererRenderTaskFn.inl:21); [1] work units finished so far, [2] work units of
   * the frame (ospGetProgress). Poked / peeked from the host through ospcuda_aux_* while the frame runs. */
  uint32_t *control;
} OcFrameBuffer;

typedef struct {
  float exposure, a, b, c, d;
  uint32_t acesColor;
} OcToneMapper;

/* a ray / hit record for ospcuda_trace (reference: struct Ray, common/Ray.ih:15-42) */
typedef struct {
  float org[3], t0, dir[3], t;
  float Ng[3], u, v;
  uint32_t primID, geomID, instID;
} OcRayHit;

/* ---- runtime ----------------------------------------------------------------------- */
OSPCUDA_API int ospcuda_init(int deviceOrdinal);
/* several CUDA devices in one process (the in-process multi-device, modules/multiDevice analogue): every device that was
 * initialised keeps its own render stream; make_current switches which one the other entry points work on */
OSPCUDA_API int ospcuda_make_current(int deviceOrdinal);
OSPCUDA_API int ospcuda_current_device(void);
/* let kernels of the current device read / write memory of `peerOrdinal` (NVLink peer access) */
OSPCUDA_API int ospcuda_enable_peer(int peerOrdinal);
OSPCUDA_API const char *ospcuda_last_error(void);
OSPCUDA_API int ospcuda_device_count(void);
OSPCUDA_API int ospcuda_malloc(void **ptr, size_t bytes);
OSPCUDA_API int ospcuda_free(void *ptr);
OSPCUDA_API int ospcuda_malloc_host(void **ptr, size_t bytes); /* pinned */
OSPCUDA_API int ospcuda_free_host(void *ptr);
OSPCUDA_API int ospcuda_memcpy_h2d(void *dst, const void *src, size_t bytes);
OSPCUDA_API int ospcuda_memcpy_d2h(void *dst, const void *src, size_t bytes);
OSPCUDA_API int ospcuda_memcpy_d2h_async(void *dst, const void *src, size_t bytes);
OSPCUDA_API int ospcuda_memset(void *dst, int value, size_t bytes);
OSPCUDA_API int ospcuda_synchronize(void);
/* one 32-bit word of device memory written / read on a side stream, concurrently with whatever the render stream runs */
OSPCUDA_API int ospcuda_aux_poke_u32(void *d_word, uint32_t value);
OSPCUDA_API int ospcuda_aux_peek_u32(const void *d_word, uint32_t *value);
/* render-stream events: create, record (on the render stream), wait, query, elapsed ms */
OSPCUDA_API int ospcuda_event_create(void **ev);
OSPCUDA_API int ospcuda_event_destroy(void *ev);
OSPCUDA_API int ospcuda_event_record(void *ev);
OSPCUDA_API int ospcuda_event_synchronize(void *ev);
OSPCUDA_API int ospcuda_stream_wait_event(void *ev); /* the current render stream waits for an event of any device */
OSPCUDA_API int ospcuda_event_query(void *ev); /* 1 done, 0 pending, <0 error */
OSPCUDA_API int ospcuda_event_elapsed_ms(void *start, void *stop, float *ms);
OSPCUDA_API void *ospcuda_stream(void); /* cudaStream_t of the render stream */
/* render on an externally owned cudaStream_t (NULL: back to the library's own stream); drains the old one */
OSPCUDA_API int ospcuda_set_stream(void *stream);
OSPCUDA_API unsigned long long ospcuda_kernel_launch_count(void);

/* ---- scene build ------------------------------------------------------------------- */
/* Build the group BVH over every primitive of models[0..numModels) (device pointers).
 * Allocates bvh->nodes / bvh->prims with ospcuda_malloc; free with ospcuda_bvh_free.  */
OSPCUDA_API int ospcuda_bvh_build_group(const OcModel *d_models, uint32_t numModels,
    const uint32_t *h_numPrims, const uint32_t *h_types, OcBVH *bvh, float *buildMs);
/* Build the top-level BVH over instance world bounds (h_bounds: 6 floats per instance). */
OSPCUDA_API int ospcuda_bvh_build_tlas(const float *h_bounds, uint32_t numInstances, OcBVH *bvh);
OSPCUDA_API int ospcuda_bvh_free(OcBVH *bvh);
/* Traversal keeps a per-thread stack of OSPCUDA_TRAVERSAL_STACK entries (the first 32 in shared memory, the rest in a
 * local-memory overflow array that costs nothing until a walk gets that deep); the top level and the group hierarchy it
 * enters share it. The builder knows the deepest stack a depth-first walk of the hierarchy it just
 * built can need (sum over the deepest path of children - 1 per wide node): the host compares and refuses a scene that
 * could overflow instead of letting traversal drop subtrees. Value of the most recent build on this thread. */
#ifndef OSPCUDA_TRAVERSAL_STACK
#define OSPCUDA_TRAVERSAL_STACK 64
#endif
OSPCUDA_API uint32_t ospcuda_bvh_last_stack_need(void);
/* options of the builds that follow (0 = default): Morton bits per axis (<= 21) and records per leaf (<= 15). The device
 * rebuilds a group with (10, 8) when the default hierarchy needs more stack than traversal keeps: equal keys are split by
 * index, so the coarser grid bounds the depth, at the price of a less tight hierarchy for that group */
OSPCUDA_API int ospcuda_bvh_set_build_options(uint32_t mortonBitsPerAxis, uint32_t leafMax);
/* largest value over the builds since the last call with reset != 0 (a world commit builds group and top-level hierarchies) */
OSPCUDA_API uint32_t ospcuda_bvh_peak_stack_need(int reset);
/* macrocell (16^3 cells) value ranges of a structured volume; d_ranges: 2*mcx*mcy*mcz floats */
OSPCUDA_API int ospcuda_volume_macrocells(const float *d_voxels, const uint32_t dims[3],
    float *d_ranges, const uint32_t mcDims[3]);

/* ---- traversal (batch API) --------------------------------------------------------- */
OSPCUDA_API int ospcuda_trace(const OcWorld *d_world, OcRayHit *d_rays, uint32_t numRays, int occlusionOnly);

/* ---- texture lookups (batch API) -----------------------------------------------------
 * get4f(TextureParam, dg) for n (st, pixelFootprint) pairs (reference: texture/TextureParam.ispc:10-22 ->
 * Texture2D_get, texture/Texture2D.ispc:133-158). `d_param->tex` and the level pointers inside it are device memory
 * (what a committed texture2d / material holds); d_st = 2n floats, d_footprint = n floats, d_out = 4n floats. */
/* the microfacet albedo tables of the energy compensation (reference: precomputeMicrofacetAlbedoTables,
 * render/bsdfs/MicrofacetAlbedoTables.ispc:148-158), computed once per device on first use; *d_tables = the float block
 * (layout: csrc/cuda/microfacet.cuh OC_MFT_*), numFloats its length */
OSPCUDA_API int ospcuda_microfacet_albedo_tables(const float **d_tables, uint32_t *numFloats);
/* the importance distribution of an hdri light (reference: HDRILight_initDistribution, lights/HDRILight.ispc:152-231 +
 * Distribution1D_create, math/Distribution1D.ispc:14-42): importance = sin(theta) * luminance(texel) per texel, row CDFs in
 * d_cdfX (sizeX * sizeY floats), the CDF of the row sums in d_cdfY (sizeY floats); d_map = a committed texture (device) */
OSPCUDA_API int ospcuda_hdri_distribution(const OcTexture2D *d_map, int32_t sizeX, int32_t sizeY, float *d_cdfX, float *d_cdfY);
OSPCUDA_API int ospcuda_texture_sample(const OcTextureParam *param, const float *d_st, const float *d_footprint, uint32_t n,
    float *d_out);

/* ---- frame ------------------------------------------------------------------------- */
OSPCUDA_API int ospcuda_render_frame(const OcWorld *d_world, const OcCamera *camera,
    const OcRenderer *renderer, const OcFrameBuffer *fb);

/* per-launch-class device timing of the wavefront pathtracer (bench.py roofline leg): while enabled every
 * launch is bracketed by CUDA events on the render stream; read returns and clears the summed milliseconds and
 * launch counts of class 0 = TRACE (traversal), 1 = SHADE, 2 = GEN (camera rays), 3 = TAIL (per-lane path completion). */
OSPCUDA_API int ospcuda_profile_enable(int on);
OSPCUDA_API int ospcuda_profile_read(float ms[4], uint32_t launches[4]);

/* ---- framebuffer ops --------------------------------------------------------------- */
OSPCUDA_API int ospcuda_fb_accumulate(const OcFrameBuffer *fb, const OcVec4f *d_sampleColor,
    const float *d_sampleDepth);
OSPCUDA_API int ospcuda_fb_variance(const OcVec4f *d_color, const OcVec4f *d_variance, int sizeX,
    int sizeY, int taskSize, float *d_taskError);
OSPCUDA_API int ospcuda_fb_tonemap(const OcVec4f *d_in, OcVec4f *d_out, uint32_t numPixels,
    const OcToneMapper *tm);
OSPCUDA_API int ospcuda_fb_color_convert(const OcVec4f *d_in, void *d_out, uint32_t numPixels,
    uint32_t colorFormat);
/* fused tonemap (optional, tm may be NULL) + colour conversion: one pass over the image */
OSPCUDA_API int ospcuda_fb_postprocess(const OcVec4f *d_in, OcVec4f *d_ppOut, void *d_out,
    uint32_t numPixels, uint32_t colorFormat, const OcToneMapper *tm);
/* the same pass restricted to the 64x64 tiles with tile % numRanks == rank (the sub-device's share of a multi-device
 * frame, MultiDevice.cpp:403-416): the rest of d_out is left to the tile exchange */
OSPCUDA_API int ospcuda_fb_postprocess_tiles(const OcVec4f *d_in, OcVec4f *d_ppOut, void *d_out, int sizeX, int sizeY,
    uint32_t colorFormat, const OcToneMapper *tm, uint32_t numRanks, uint32_t rank);
/* copy the 64x64 tiles owned by `srcRank` (tile % numRanks == srcRank) from a rank-local
 * full-size image into dst (both device, row-major, bytesPerPixel each) */
OSPCUDA_API int ospcuda_fb_gather_tiles(void *d_dst, const void *d_src, int sizeX, int sizeY,
    uint32_t bytesPerPixel, uint32_t numRanks, uint32_t srcRank);

/* NCCL tile gather staging (the multi-device path, modules/multiDevice/MultiDeviceLoadBalancer.cpp:41-47 +
 * fb/LocalFB.cpp:286-353 writeTiles): pack copies the 64x64 tiles owned by `rank` (tile % numRanks == rank)
 * from a row-major image into a tile-major buffer (owned tile k at k*64*64*bytesPerPixel, border tiles
 * padded); unpack scatters a received shard of `srcRank` into the row-major image on GPU 0. */
OSPCUDA_API int ospcuda_fb_pack_tiles(void *d_packed, const void *d_image, int sizeX, int sizeY,
    uint32_t bytesPerPixel, uint32_t numRanks, uint32_t rank);
OSPCUDA_API int ospcuda_fb_unpack_tiles(void *d_image, const void *d_packed, int sizeX, int sizeY,
    uint32_t bytesPerPixel, uint32_t numRanks, uint32_t srcRank);

/* Peer-memory tile exchange (peer_exchange.cu): the same composition (MultiDeviceLoadBalancer.cpp:41-47 +
 * LocalFB.cpp:286-353 writeTiles) without staging or a collective call, for one process per GPU on an NVLink/NVSwitch
 * box. Rank 0 exports its display image and its synchronisation block with ospcuda_ipc_export; the other ranks map
 * them (ospcuda_ipc_open) and push their owned tiles straight into rank 0's image. Per frame (epoch = 1, 2, ...):
 * rank 0 calls post_ready at the start of the frame (the previous image has been consumed in stream order) and
 * wait_arrive at its end; every other rank calls push_tiles after its colour conversion. All calls enqueue on the render
 * stream and return at once; spins on the device are bounded and report through ospcuda_peer_error. */
#define OSPCUDA_IPC_HANDLE_BYTES 64
#define OSPCUDA_PEER_BLOCK_WORDS 64
#define OSPCUDA_MAX_PEERS 16
OSPCUDA_API int ospcuda_peer_block_alloc(uint32_t **d_block); /* zeroed; free with ospcuda_free */
OSPCUDA_API int ospcuda_ipc_export(const void *d_ptr, unsigned char *handle /* OSPCUDA_IPC_HANDLE_BYTES */);
OSPCUDA_API int ospcuda_ipc_open(const unsigned char *handle, void **d_ptr);
OSPCUDA_API int ospcuda_ipc_close(void *d_ptr);
OSPCUDA_API int ospcuda_peer_post_ready(uint32_t *const *peerBlocks /* host array [numRanks], entry 0 unused */,
    uint32_t numRanks, uint32_t epoch);
OSPCUDA_API int ospcuda_peer_wait_arrive(uint32_t *d_rootBlock, uint32_t numRanks, uint32_t epoch);
OSPCUDA_API int ospcuda_peer_push_tiles(void *peerImage, const void *d_image, int sizeX, int sizeY, uint32_t bytesPerPixel,
    uint32_t numRanks, uint32_t rank, uint32_t *d_localBlock, uint32_t *peerRootBlock, uint32_t epoch);
OSPCUDA_API int ospcuda_peer_error(const uint32_t *d_block, uint32_t *timedOut); /* synchronous read of the block's error word */

#ifdef __cplusplus
}
#endif
#endif /* OSPCUDA_H */
