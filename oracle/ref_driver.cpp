// ref_driver.cpp -- runs the REFERENCE's own render kernels (compiled from /root/reference/modules/cpu/**/*.ispc in
// the scalar C++ mode its `ospray_module_gpu_kernels` target uses, see oracle/build_ref.py) on the scene PODs of
// include/ospcuda.h.  TEST INFRASTRUCTURE ONLY: oracle/_ref/libospref.so is loaded by tests/, smoke() and
// bench.py's reference arm, never by the product.
//
// What this file does is the reference's HOST side for this path, restated: it fills the reference's shared
// structs (ispc::World, Instance, Group, GeometricModel, Mesh/Spheres/Boxes, OBJ, QuadLight/DirectionalLight/
// AmbientLight, PathTracerData/SciVisData, PathTracer/SciVis/AORenderer, PerspectiveCamera, LocalFB) from the
// already-committed Oc* values the way ISPCDevice's commit code does (file:line beside each block), then calls
// the reference's own launchers (PathTracer_renderTaskLauncher etc., DEFINE_RENDERER_KERNEL_LAUNCHER,
// common/DeviceRT.ih:60-94).  No reference source is copied: the kernels are compiled where they lie.
// std headers first: common/OSPCommon.ih defines the macro `in` for its scalar foreach_unique emulation
#include <algorithm>
#include <deque>
#include <iostream>
#include <memory>
#include <thread>
#include <vector>
#include "ref_embree_stub.h"
#include "ospray/OSPEnums.h"
#include "common/OSPCommon.ih"
#include "common/Data.ih"
#include "common/World.ih"
#include "common/GroupShared.h"
#include "geometry/GeometricModelShared.h"
#include "geometry/MeshShared.h"
#include "geometry/SpheresShared.h"
#include "geometry/BoxesShared.h"
#include "geometry/CurvesShared.h"
#include "render/MaterialShared.h"
#include "render/materials/OBJShared.h"
#include "render/materials/ThinGlassShared.h"
#include "render/materials/LuminousShared.h"
#include "render/materials/MetalShared.h"
#include "render/materials/PlasticShared.h"
#include "render/materials/PrincipledShared.h"
#include "render/materials/GlassShared.h"
#include "render/materials/AlloyShared.h"
#include "render/materials/VelvetShared.h"
#include "render/materials/CarPaintShared.h"
#include "render/materials/MetallicPaintShared.h"
#include "render/bsdfs/MicrofacetAlbedoTablesShared.h"
#include "render/pathtracer/GeometryLightShared.h"
#include "texture/Texture2DShared.h"
#include "common/DifferentialGeometry.ih"
#include "texture/TextureParamShared.h"
#include "lights/QuadLightShared.h"
#include "lights/PointLightShared.h"
#include "lights/SpotLightShared.h"
#include "lights/DirectionalLightShared.h"
#include "lights/AmbientLightShared.h"
#include "lights/HDRILightShared.h"
#include "lights/CylinderLightShared.h"
#include "math/Distribution2DShared.h"
#include "render/pathtracer/PathTracerShared.h"
#include "render/scivis/SciVisShared.h"
#include "render/ao/AORendererShared.h"
#include "render/debug/DebugRendererShared.h"
#include "pf/LUTPixelFilterShared.h"
#include "camera/PerspectiveCameraShared.h"
#include "camera/OrthographicCameraShared.h"
#include "camera/PanoramicCameraShared.h"
#include "fb/LocalFBShared.h"
#include "fb/FrameBufferView.ih"
#include "fb/RenderTaskDesc.ih"
#include "fb/frame_ops/ColorConversionShared.h"
#include "fb/frame_ops/ToneMapperShared.h"
#include "fb/frame_ops/VarianceShared.h"
#include "render/ScreenSample.ih"
#include "math/random.ih"
#include "math/halton.ih"
#include "render/util.ih"
#ifdef OSPRAY_ENABLE_VOLUMES
#include "volume/VolumeShared.h"
#include "volume/VolumetricModelShared.h"
#include "volume/transferFunction/LinearTransferFunctionShared.h"
#include "ref_vkl_stub.h"
#endif

#define REF_API extern "C" __attribute__((visibility("default")))

using namespace ispc;

// the reference's launchers (extern "C", defined by DEFINE_*_KERNEL_LAUNCHER in the kernel TUs)
extern "C" {
void PathTracer_renderTaskLauncher(void *, void *, const vec3ui &, Renderer *, FrameBuffer *, Camera *, World *, const uint32_t *, const ospray::FeatureFlags &);
void SciVis_renderTaskLauncher(void *, void *, const vec3ui &, Renderer *, FrameBuffer *, Camera *, World *, const uint32_t *, const ospray::FeatureFlags &);
void AORenderer_renderTaskLauncher(void *, void *, const vec3ui &, Renderer *, FrameBuffer *, Camera *, World *, const uint32_t *, const ospray::FeatureFlags &);
void DebugRenderer_renderTaskLauncher(void *, void *, const vec3ui &, Renderer *, FrameBuffer *, Camera *, World *, const uint32_t *, const ospray::FeatureFlags &);
void ColorConversion_kernelLauncher(void *, void *, const vec2ui &, const FrameBufferView *);
void ToneMapper_kernelLauncher(void *, void *, const vec2ui &, const FrameBufferView *);
void Variance_kernelLauncher(void *, void *, const vec2ui &, const FrameBufferView *);
void ospref_ray_counters(uint64_t out[2], int reset);
}
namespace ispc {
void LocalFB_accumulateSample(FrameBuffer *fb, const ScreenSample &screenSample, RenderTaskDesc &);
void LUTPixelFilter_buildLUT(void *self);
float Distribution1D_create(const int size, float *cdf);
int GeometricModel_gatherEmissivePrimIDs(void *model, void *renderer, void *instance, int *primIDs, float *distribution, float &pdf);
void precomputeMicrofacetAlbedoTables(void *tables);
void HDRILight_initDistribution(const void *map, void *distribution);
void Mesh_getAreas(const Geometry *const, const int32 *const, const int32, const affine3f &, float *const);
void Spheres_getAreas(const Geometry *const, const int32 *const, const int32, const affine3f &, float *const);
}

namespace {

inline vec3f v3(const OcVec3f &o) { return make_vec3f(o.x, o.y, o.z); }
inline Data1D d1(const OcData1D &d)
{
  Data1D r; r.addr = (uint8 *)d.addr; r.byteStride = d.byteStride; r.numItems = d.numItems; r.huge = false;
  if (!d.addr) { r.byteStride = 0; r.numItems = 0; }
  return r;
}
inline affine3f a3(const OcAffine3f &a) { return affine3f(linear3f(v3(a.vx), v3(a.vy), v3(a.vz)), v3(a.p)); }

ospray::FeatureFlags allFeatures()
{
  ospray::FeatureFlags ff;
  ff.geometry = ospray::FFG_ALL;
#ifdef OSPRAY_ENABLE_VOLUMES
  ff.volume = VKL_FEATURE_FLAGS_DEFAULT;
#endif
  ff.other = ospray::FFO_ALL;
  return ff;
}

#ifdef OSPRAY_ENABLE_VOLUMES
// volume/Volume.cpp:60-222 (structuredRegular -> VKL handles), volume/VolumetricModel.cpp:29-82,
// volume/transferFunction/LinearTransferFunction.cpp:22-33 and :110-160 (precomputeMaxOpacityRanges: the table the path
// tracer's delta tracking reads through LinearTransferFunction_getMaxOpacity)
VolumetricModel *refMakeVolumetricModel(std::deque<Volume> &vols, std::deque<VolumetricModel> &vms, std::deque<LinearTransferFunction> &tfs,
    std::deque<RefVklVolume> &vklVols, std::deque<RefVklSampler> &, std::deque<RefVklIntervalContext> &, const OcVolume &ov)
{
  vklVols.push_back(RefVklVolume{&ov});
  vols.emplace_back();
  Volume &vol = vols.back();
  vol.vklVolume.host = &vklVols.back(); vol.vklSampler.host = &vklVols.back();
  vol.boundingBox = box3f(make_vec3f(ov.bounds[0], ov.bounds[1], ov.bounds[2]), make_vec3f(ov.bounds[3], ov.bounds[4], ov.bounds[5]));
  tfs.emplace_back();
  LinearTransferFunction &tf = tfs.back();
  memset(tf.maxOpacityInRange, 0, sizeof(tf.maxOpacityInRange));
  if (ov.tf.numOpacities) {
    const float *opacities = ov.tf.opacity;
    const int N = PRECOMPUTED_OPACITY_SUBRANGE_COUNT, maxOpacityDim = (int)ov.tf.numOpacities - 1, maxPrecomputedDim = N - 1;
    for (int i = 0; i < N; i++) { // the diagonal
      const int lo = (int)floorf(maxOpacityDim * (float)i / maxPrecomputedDim), hi = (int)ceilf(maxOpacityDim * (float)i / maxPrecomputedDim);
      float m = opacities[lo];
      for (int k = lo; k <= hi; k++) m = std::max(m, opacities[k]);
      tf.maxOpacityInRange[i][i] = m;
    }
    for (int i = 0; i < N; i++) // each column from the diagonal up
      for (int j = i + 1; j < N; j++) {
        const int lo = (int)floorf(maxOpacityDim * (float)i / maxPrecomputedDim), hi = (int)ceilf(maxOpacityDim * (float)j / maxPrecomputedDim);
        float m = tf.maxOpacityInRange[i][i];
        for (int k = lo; k <= hi; k++) m = std::max(m, opacities[k]);
        tf.maxOpacityInRange[i][j] = m;
      }
  }
  tf.super.valueRange = box1f(ov.tf.valueLower, ov.tf.valueUpper);
  tf.color.addr = (uint8 *)ov.tf.color; tf.color.byteStride = 12; tf.color.numItems = ov.tf.numColors; tf.color.huge = false;
  tf.opacity.addr = (uint8 *)ov.tf.opacity; tf.opacity.byteStride = 4; tf.opacity.numItems = ov.tf.numOpacities; tf.opacity.huge = false;
  vms.emplace_back();
  VolumetricModel &vm = vms.back();
  vm.volume = &vol; vm.transferFunction = &tf.super; vm.vklIntervalContext.host = &vklVols.back();
  vm.boundingBox = vol.boundingBox; vm.densityScale = ov.densityScale; vm.anisotropy = ov.anisotropy;
  vm.gradientShadingScale = ov.gradientShadingScale; vm.userID = ov.userID;
  return &vm;
}
#endif

// Everything one frame needs, with stable addresses (deques never move their elements)
struct Scene
{
  std::deque<Mesh> meshes; std::deque<Spheres> spheres; std::deque<Boxes> boxes; std::deque<Curves> curves;
  std::deque<GeometricModel> models; std::deque<Group> groups; std::deque<Instance> instances;
  std::deque<std::vector<GeometricModel *>> modelPtrs;
  std::deque<std::vector<Material *>> materialPtrs;
  std::deque<OBJ> objs; std::deque<ThinGlass> glasses; std::vector<Material *> allMaterials;
  std::deque<Texture2D> textures;
  std::deque<Luminous> luminous; std::deque<GeometryLight> geoLights;
  std::deque<Metal> metals; std::deque<Plastic> plastics; std::deque<Principled> principleds; std::deque<Glass> solidGlasses; std::deque<Alloy> alloys; std::deque<Velvet> velvets; std::deque<MetallicPaint> paints; std::deque<CarPaint> carPaints;
  std::deque<std::vector<int>> geoPrimIDs; std::deque<std::vector<float>> geoDistributions;
  std::vector<float> geoCDF;
  std::vector<Instance *> instancePtrs;
  std::deque<QuadLight> quads; std::deque<PointLight> points; std::deque<SpotLight> spots; std::deque<DirectionalLight> dirs; std::deque<AmbientLight> ambients; std::deque<HDRILight> hdris; std::deque<CylinderLight> cylinders; std::deque<Distribution2D> hdriDists; std::deque<std::vector<float>> hdriCdfs;
  std::vector<Light *> ptLights, svLights;
  Instance identityInstance; // stands in for the NULL instance of world lights: the scalar light dispatch reads
                             // self->instance->motionBlur unconditionally (lights/LightDispatch.ispc:36-60)
  PathTracerData ptData; SciVisData svData;
  World world;
  RTCSceneTy geomScene, volScene;
#ifdef OSPRAY_ENABLE_VOLUMES
  std::deque<Volume> volumes; std::deque<VolumetricModel> vmodels; std::deque<LinearTransferFunction> tfs;
  std::deque<std::vector<VolumetricModel *>> vmodelPtrs;
  std::deque<RefVklVolume> vklVolumes; std::deque<RefVklSampler> vklSamplers; std::deque<RefVklIntervalContext> vklContexts;
#endif
};

Geometry *makeGeometry(Scene &s, const OcGeometry *g)
{
  if (g->type == OC_GEOM_TRIANGLES || g->type == OC_GEOM_QUADS) {
    // geometry/Mesh.cpp:137-196
    s.meshes.emplace_back();
    Mesh &m = s.meshes.back();
    const bool isTri = g->type == OC_GEOM_TRIANGLES;
    m.super.type = isTri ? GEOMETRY_TYPE_TRIANGLE_MESH : GEOMETRY_TYPE_QUAD_MESH;
    m.super.numPrimitives = g->numPrims;
    m.index = d1(g->index); m.vertex = d1(g->vertex); m.normal = d1(g->normal); m.color = d1(g->color); m.texcoord = d1(g->texcoord);
    m.has_alpha = (g->flags & OC_GEOMFLAG_COLOR_IS_VEC4) != 0 && g->color.addr;
    m.is_triangleMesh = isTri;
    m.isColorFaceVarying = (g->flags & OC_GEOMFLAG_COLOR_FACEVARYING) != 0;
    m.isTexcoordFaceVarying = (g->flags & OC_GEOMFLAG_TEXCOORD_FACEVARYING) != 0;
    m.isNormalFaceVarying = (g->flags & OC_GEOMFLAG_NORMAL_FACEVARYING) != 0;
    m.super.getAreas = Mesh_getAreas; // geometry/Mesh.cpp:26-29
    m.flagMask = -1;
    if (!g->normal.addr) m.flagMask &= int64(~DG_NS);
    if (!g->color.addr) m.flagMask &= int64(~DG_COLOR);
    if (!g->texcoord.addr) m.flagMask &= int64(~DG_TEXCOORD);
    return &m.super;
  }
  if (g->type == OC_GEOM_SPHERES) {
    // geometry/Spheres.cpp:46-106 (the interleaved vec4f array is the committed state)
    s.spheres.emplace_back();
    Spheres &sp = s.spheres.back();
    sp.super.numPrimitives = g->numPrims;
    sp.sphere = d1(g->sphere); sp.texcoord = d1(g->texcoord);
    Data1D_Constructor(&sp.normalData);
    sp.sphereType = OSP_SPHERE;
    sp.super.getAreas = Spheres_getAreas; // geometry/Spheres.cpp:32-38
    return &sp.super;
  }
  if (g->type == OC_GEOM_CURVES) {
    // geometry/Curves.cpp:84-194 (round curves; the Embree geometry handle is not used by Curves_postIntersect)
    s.curves.emplace_back();
    Curves &c = s.curves.back();
    c.super.numPrimitives = g->numPrims;
    c.curveType = OSP_ROUND; c.curveBasis = (OSPCurveBasis)g->curveBasis;
    c.index = d1(g->index); c.color = d1(g->color); c.texcoord = d1(g->texcoord);
    c.flagMask = -1;
    if (!g->color.addr) c.flagMask &= int64(~DG_COLOR);
    if (!g->texcoord.addr) c.flagMask &= int64(~DG_TEXCOORD);
    return &c.super;
  }
  // geometry/Boxes.cpp:33-42
  s.boxes.emplace_back();
  Boxes &b = s.boxes.back();
  b.super.numPrimitives = g->numPrims;
  b.boxes = d1(g->box);
  return &b.super;
}

// ispc::Texture2D::set (texture/Texture2D.cpp:12-39) from the committed levels; ispc::TextureParam as Material::getTextureParam
// leaves it (render/Material.cpp:73-108)
Texture2D *makeTexture(std::deque<Texture2D> &store, const OcTexture2D *t)
{
  if (!t) return nullptr;
  store.emplace_back();
  Texture2D &r = store.back();
  r.size = make_vec2i(t->sizeX, t->sizeY);
  r.maxLevel = t->maxLevel;
  r.format = (OSPTextureFormat)t->format; r.filter = (OSPTextureFilter)t->filter;
  r.wrapMode = make_vec2ui(t->wrapS, t->wrapT);
  for (int l = 0; l <= t->maxLevel; l++) r.data[l] = (void *)t->data[l];
  r.super.hasAlpha = t->hasAlpha != 0;
  return &r;
}
TextureParam makeTextureParam(std::deque<Texture2D> &store, const OcTextureParam &p)
{
  TextureParam r;
  Texture2D *t = makeTexture(store, p.tex);
  r.ptr = t ? &t->super : nullptr;
  r.transformFlags = (p.transformFlags & OC_TEX_TRANSFORM_2D) ? TRANSFORM_FLAG_2D : TRANSFORM_FLAG_NONE;
  r.scale2f = p.scale2f;
  r.xform2f = affine2f(linear2f(make_vec2f(p.xform2f[0], p.xform2f[1]), make_vec2f(p.xform2f[2], p.xform2f[3])), make_vec2f(p.xform2f[4], p.xform2f[5]));
  return r;
}

// the reference's microfacet albedo tables, computed once per process by the reference's own precompute routine
// (render/bsdfs/MicrofacetAlbedoTables.ispc:148-158; storage as MicrofacetAlbedoTables.cpp allocates it)
MicrofacetAlbedoTables *refAlbedoTables()
{
  static MicrofacetAlbedoTables t;
  static std::vector<float> store;
  static bool done = false;
  if (!done) {
    const size_t a = MICROFACET_ALBEDO_TABLE_SIZE, d = MICROFACET_DIELECTRIC_ALBEDO_TABLE_SIZE, sh = MICROFACET_SHEEN_ALBEDO_TABLE_SIZE;
    store.assign(a * a + a + 3 * (d * d * d + d * d) + sh * sh, 0.f);
    float *p = store.data();
    t.albedo_dir = p; p += a * a; t.albedo_avg = p; p += a;
    t.dielectricAlbedo_dir = p; p += d * d * d; t.dielectricAlbedo_avg = p; p += d * d;
    t.dielectricRcpEtaAlbedo_dir = p; p += d * d * d; t.dielectricRcpEtaAlbedo_avg = p; p += d * d;
    t.dielectricReflectionAlbedo_dir = p; p += d * d * d; t.dielectricReflectionAlbedo_avg = p; p += d * d;
    t.sheenAlbedo_dir = p;
    precomputeMicrofacetAlbedoTables(&t);
    done = true;
  }
  return &t;
}

void buildScene(Scene &s, const OcWorld *w, const OcRenderer *ren)
{
  // materials: render/materials/OBJ.cpp:30-66 (values arrive already normalised by the commit)
  for (uint32_t i = 0; i < w->numMaterials; i++) {
    const OcMaterial &m = w->materials[i];
    if (m.type == OC_MATERIAL_THINGLASS) { // render/materials/ThinGlass.cpp:33-48 (committed values)
      s.glasses.emplace_back();
      ThinGlass &g = s.glasses.back();
      g.eta = m.eta; g.attenuation = v3(m.attenuation); g.attenuationScale = m.attenuationScale;
      g.attenuationColorMap = makeTextureParam(s.textures, m.mapKd);
      s.allMaterials.push_back(&g.super);
      continue;
    }
    if (m.type == OC_MATERIAL_METAL) { // render/materials/Metal.cpp:30-76, RGB form
      s.metals.emplace_back();
      Metal &mt = s.metals.back();
      mt.spectral = false; mt.etaRGB = v3(m.Kd); mt.kRGB = v3(m.Ks); mt.roughness = m.Ns;
      mt.roughnessMap = makeTextureParam(s.textures, m.mapNs);
      mt.super.microfacetAlbedoTables = refAlbedoTables();
      s.allMaterials.push_back(&mt.super);
      continue;
    }
    if (m.type == OC_MATERIAL_PLASTIC) { // render/materials/Plastic.cpp:28-38 (eta arrives as 1 / ior)
      s.plastics.emplace_back();
      Plastic &pl = s.plastics.back();
      pl.pigmentColor = v3(m.Kd); pl.eta = m.eta; pl.roughness = m.Ns;
      pl.super.microfacetAlbedoTables = refAlbedoTables();
      s.allMaterials.push_back(&pl.super);
      continue;
    }
    if (m.type == OC_MATERIAL_GLASS) { // render/materials/Glass.cpp:35-63 (committed values)
      s.solidGlasses.emplace_back();
      Glass &gl = s.solidGlasses.back();
      gl.mediumInside.ior = m.eta; gl.mediumInside.attenuation = v3(m.attenuation);
      gl.mediumOutside.ior = m.principled.outsideIor; gl.mediumOutside.attenuation = v3(m.principled.outsideAttenuation);
      s.allMaterials.push_back(&gl.super);
      continue;
    }
    if (m.type == OC_MATERIAL_ALLOY) { // render/materials/Alloy.cpp:28-40
      s.alloys.emplace_back();
      Alloy &al = s.alloys.back();
      al.color = v3(m.Kd); al.edgeColor = v3(m.Ks); al.roughness = m.Ns;
      al.colorMap = makeTextureParam(s.textures, m.mapKd); al.edgeColorMap = makeTextureParam(s.textures, m.mapKs);
      al.roughnessMap = makeTextureParam(s.textures, m.mapNs);
      al.super.microfacetAlbedoTables = refAlbedoTables();
      s.allMaterials.push_back(&al.super);
      continue;
    }
    if (m.type == OC_MATERIAL_CARPAINT) { // render/materials/CarPaint.cpp:27-87 (committed values)
      s.carPaints.emplace_back();
      CarPaint &c = s.carPaints.back();
      const OcCarPaint &q = *m.carPaint;
      auto tp = [&](int i) { return makeTextureParam(s.textures, q.map[i]); };
      auto rot = [](const float *r) { return linear2f(make_vec2f(r[0], r[1]), make_vec2f(r[2], r[3])); };
      c.baseColor = v3(q.baseColor); c.baseColorMap = tp(OC_CPMAP_BASECOLOR);
      c.roughness = q.roughness; c.roughnessMap = tp(OC_CPMAP_ROUGHNESS);
      c.normal = q.normalScale; c.normalMap = tp(OC_CPMAP_NORMAL); c.normalRot = rot(q.normalRot);
      c.useFlakeColor = q.useFlakeColor != 0; c.flakeColor = v3(q.flakeColor); c.flakeColorMap = tp(OC_CPMAP_FLAKECOLOR);
      c.flakeScale = q.flakeScale; c.flakeScaleMap = tp(OC_CPMAP_FLAKESCALE);
      c.flakeDensity = q.flakeDensity; c.flakeDensityMap = tp(OC_CPMAP_FLAKEDENSITY);
      c.flakeSpread = q.flakeSpread; c.flakeSpreadMap = tp(OC_CPMAP_FLAKESPREAD);
      c.flakeJitter = q.flakeJitter; c.flakeJitterMap = tp(OC_CPMAP_FLAKEJITTER);
      c.flakeRoughness = q.flakeRoughness; c.flakeRoughnessMap = tp(OC_CPMAP_FLAKEROUGHNESS);
      c.coat = q.coat; c.coatMap = tp(OC_CPMAP_COAT); c.coatIor = q.coatIor; c.coatIorMap = tp(OC_CPMAP_COATIOR);
      c.coatColor = v3(q.coatColor); c.coatColorMap = tp(OC_CPMAP_COATCOLOR);
      c.coatThickness = q.coatThickness; c.coatThicknessMap = tp(OC_CPMAP_COATTHICKNESS);
      c.coatRoughness = q.coatRoughness; c.coatRoughnessMap = tp(OC_CPMAP_COATROUGHNESS);
      c.coatNormal = q.coatNormalScale; c.coatNormalMap = tp(OC_CPMAP_COATNORMAL); c.coatNormalRot = rot(q.coatNormalRot);
      c.flipflopColor = v3(q.flipflopColor); c.flipflopColorMap = tp(OC_CPMAP_FLIPFLOPCOLOR);
      c.flipflopFalloff = q.flipflopFalloff; c.flipflopFalloffMap = tp(OC_CPMAP_FLIPFLOPFALLOFF);
      c.super.microfacetAlbedoTables = refAlbedoTables();
      s.allMaterials.push_back(&c.super);
      continue;
    }
    if (m.type == OC_MATERIAL_VELVET) { // render/materials/Velvet.cpp:27-39
      s.velvets.emplace_back();
      Velvet &ve = s.velvets.back();
      ve.reflectance = v3(m.Kd); ve.backScattering = m.Ns; ve.horizonScatteringColor = v3(m.Ks); ve.horizonScatteringFallOff = m.eta;
      s.allMaterials.push_back(&ve.super);
      continue;
    }
    if (m.type == OC_MATERIAL_METALLICPAINT) { // render/materials/MetallicPaint.cpp:27-40 (committed values)
      s.paints.emplace_back();
      MetallicPaint &mp = s.paints.back();
      mp.baseColor = v3(m.Kd); mp.baseColorMap = makeTextureParam(s.textures, m.mapKd);
      mp.flakeAmount = m.attenuationScale; mp.flakeColor = v3(m.Ks); mp.flakeSpread = m.Ns; mp.eta = m.eta;
      mp.super.microfacetAlbedoTables = refAlbedoTables();
      s.allMaterials.push_back(&mp.super);
      continue;
    }
    if (m.type == OC_MATERIAL_PRINCIPLED) { // render/materials/Principled.cpp:33-150 (committed values, no maps)
      s.principleds.emplace_back();
      Principled &pr = s.principleds.back();
      const OcPrincipled &q = m.principled;
      pr.baseColor = v3(q.baseColor); pr.edgeColor = v3(q.edgeColor);
      pr.metallic = q.metallic; pr.diffuse = q.diffuse; pr.specular = q.specular; pr.specularMetallic = q.specularMetallic != 0;
      pr.ior = q.ior; pr.transmission = q.transmission; pr.transmissionColor = v3(q.transmissionColor);
      pr.transmissionDepth = q.transmissionDepth; pr.roughness = q.roughness; pr.anisotropy = q.anisotropy; pr.rotation = q.rotation;
      pr.coat = q.coat; pr.coatIor = q.coatIor; pr.coatColor = v3(q.coatColor); pr.coatThickness = q.coatThickness;
      pr.coatRoughness = q.coatRoughness; pr.opacity = q.opacity;
      pr.sheen = q.sheen; pr.sheenColor = v3(q.sheenColor); pr.sheenTint = q.sheenTint; pr.sheenRoughness = q.sheenRoughness;
      pr.thin = q.thin != 0; pr.backlight = q.backlight; pr.thickness = q.thickness;
      pr.outsideMedium.ior = q.outsideIor; pr.outsideMedium.attenuation = v3(q.outsideAttenuation);
      pr.emission = v3(m.emission);
      if (const OcPrincipledMaps *mp = q.maps) { // PrincipledShared.h: one TextureParam per factor, rotations of the normal maps
        auto tp = [&](int i) { return makeTextureParam(s.textures, mp->map[i]); };
        auto rot = [](const float *r) { return linear2f(make_vec2f(r[0], r[1]), make_vec2f(r[2], r[3])); };
        pr.baseColorMap = tp(OC_PRMAP_BASECOLOR); pr.edgeColorMap = tp(OC_PRMAP_EDGECOLOR); pr.metallicMap = tp(OC_PRMAP_METALLIC);
        pr.diffuseMap = tp(OC_PRMAP_DIFFUSE); pr.specularMap = tp(OC_PRMAP_SPECULAR); pr.iorMap = tp(OC_PRMAP_IOR);
        pr.transmissionMap = tp(OC_PRMAP_TRANSMISSION); pr.transmissionColorMap = tp(OC_PRMAP_TRANSMISSIONCOLOR);
        pr.transmissionDepthMap = tp(OC_PRMAP_TRANSMISSIONDEPTH); pr.roughnessMap = tp(OC_PRMAP_ROUGHNESS);
        pr.anisotropyMap = tp(OC_PRMAP_ANISOTROPY); pr.rotationMap = tp(OC_PRMAP_ROTATION);
        pr.normalMap = tp(OC_PRMAP_NORMAL); pr.normalRot = rot(mp->normalRot); pr.normal = mp->normalScale;
        pr.baseNormalMap = tp(OC_PRMAP_BASENORMAL); pr.baseNormalRot = rot(mp->baseNormalRot); pr.baseNormal = mp->baseNormalScale;
        pr.coatMap = tp(OC_PRMAP_COAT); pr.coatIorMap = tp(OC_PRMAP_COATIOR); pr.coatColorMap = tp(OC_PRMAP_COATCOLOR);
        pr.coatThicknessMap = tp(OC_PRMAP_COATTHICKNESS); pr.coatRoughnessMap = tp(OC_PRMAP_COATROUGHNESS);
        pr.coatNormalMap = tp(OC_PRMAP_COATNORMAL); pr.coatNormalRot = rot(mp->coatNormalRot); pr.coatNormal = mp->coatNormalScale;
        pr.sheenMap = tp(OC_PRMAP_SHEEN); pr.sheenColorMap = tp(OC_PRMAP_SHEENCOLOR); pr.sheenTintMap = tp(OC_PRMAP_SHEENTINT);
        pr.sheenRoughnessMap = tp(OC_PRMAP_SHEENROUGHNESS); pr.opacityMap = tp(OC_PRMAP_OPACITY); pr.backlightMap = tp(OC_PRMAP_BACKLIGHT);
        pr.thicknessMap = tp(OC_PRMAP_THICKNESS); pr.emissionMap = tp(OC_PRMAP_EMISSION);
      }
      pr.super.isEmissive = m.isEmissive != 0;
      pr.super.microfacetAlbedoTables = refAlbedoTables();
      s.allMaterials.push_back(&pr.super);
      continue;
    }
    if (m.type == OC_MATERIAL_LUMINOUS) { // render/materials/Luminous.cpp:33-45 (committed values)
      s.luminous.emplace_back();
      Luminous &lm = s.luminous.back();
      lm.radiance = v3(m.emission); lm.transparency = m.transparency;
      lm.emissionMap = makeTextureParam(s.textures, m.mapKd);
      lm.super.isEmissive = m.isEmissive != 0;
      s.allMaterials.push_back(&lm.super);
      continue;
    }
    s.objs.emplace_back();
    OBJ &o = s.objs.back();
    o.d = m.d; o.Kd = v3(m.Kd); o.Ks = v3(m.Ks); o.Ns = m.Ns; o.Tf = v3(m.Tf);
    o.dMap = makeTextureParam(s.textures, m.mapD); o.KdMap = makeTextureParam(s.textures, m.mapKd);
    o.KsMap = makeTextureParam(s.textures, m.mapKs); o.NsMap = makeTextureParam(s.textures, m.mapNs);
    o.bumpMap = makeTextureParam(s.textures, m.mapBump);
    o.bumpRot = linear2f(make_vec2f(m.bumpRot[0], m.bumpRot[1]), make_vec2f(m.bumpRot[2], m.bumpRot[3]));
    s.allMaterials.push_back(&o.super);
  }
  bool anyGeometry = false, anyVolume = false;
  s.geomScene.world = w; s.volScene.world = w; s.volScene.volumes = true;
  s.geomScene.userPtr.resize(w->numInstances); s.volScene.userPtr.resize(w->numInstances);
  for (uint32_t i = 0; i < w->numInstances; i++) {
    const OcInstance &oi = w->instances[i];
    const OcGroup *og = oi.group;
    // common/Group.cpp:69-209
    s.groups.emplace_back();
    Group &grp = s.groups.back();
    s.modelPtrs.emplace_back();
    for (uint32_t m = 0; m < og->numModels; m++) {
      const OcModel &om = og->models[m];
      // geometry/GeometricModel.cpp:20-88
      s.models.emplace_back();
      GeometricModel &gm = s.models.back();
      gm.geom = makeGeometry(s, om.geom);
      gm.color = d1(om.color); gm.index = d1(om.index);
      gm.numMaterials = om.numMaterials;
      if (om.materialMode == OC_MAT_FROM_MODEL && om.numMaterials) {
        s.materialPtrs.emplace_back();
        for (uint32_t k = 0; k < om.numMaterials; k++) s.materialPtrs.back().push_back(s.allMaterials[om.material[k]]);
        gm.material = s.materialPtrs.back().data();
      } else if (om.materialMode == OC_MAT_FROM_RENDERER && om.numMaterials) {
        gm.materialID = (uint32 *)om.material;
      } else
        gm.numMaterials = 0;
      gm.invertedNormals = om.invertedNormals != 0;
      gm.userID = om.userID;
      s.modelPtrs.back().push_back(&gm);
      s.geomScene.userPtr[i].push_back(gm.geom);
      anyGeometry = true;
    }
    grp.geometricModels = s.modelPtrs.back().empty() ? nullptr : s.modelPtrs.back().data();
    grp.numGeometricModels = og->numModels;
#ifdef OSPRAY_ENABLE_VOLUMES
    s.vmodelPtrs.emplace_back();
    for (uint32_t v = 0; v < og->numVolumes; v++) {
      VolumetricModel *vm = refMakeVolumetricModel(s.volumes, s.vmodels, s.tfs, s.vklVolumes, s.vklSamplers, s.vklContexts, og->volumes[v]);
      s.vmodelPtrs.back().push_back(vm);
      s.volScene.userPtr[i].push_back(vm->volume);
      anyVolume = true;
    }
    grp.volumetricModels = s.vmodelPtrs.back().empty() ? nullptr : s.vmodelPtrs.back().data();
    grp.numVolumetricModels = og->numVolumes;
#endif
    // common/Instance.cpp:20-34
    s.instances.emplace_back();
    Instance &inst = s.instances.back();
    inst.group = &grp; inst.scene = &s.geomScene; inst.geomID = i;
    inst.xfm = a3(oi.xfm); inst.rcp_xfm = a3(oi.rcp_xfm);
    inst.userID = oi.userID;
  }
  for (auto &inst : s.instances) s.instancePtrs.push_back(&inst);

  // lights: lights/QuadLight.cpp:17-59, DirectionalLight.cpp, AmbientLight.cpp (committed values)
  auto makeLight = [&](const OcLight &l) -> Light * {
    if (l.type == OC_LIGHT_QUAD) {
      s.quads.emplace_back();
      QuadLight &q = s.quads.back();
      q.super.isVisible = l.isVisible != 0; q.super.instance = &s.identityInstance;
      q.radiance = v3(l.radiance);
      q.pre.position = v3(l.position); q.pre.edge1 = v3(l.edge1); q.pre.edge2 = v3(l.edge2);
      q.pre.nnormal = v3(l.nnormal); q.pre.ppdf = l.ppdf;
      q.pre.c90 = normalize(cross(q.pre.nnormal, q.pre.edge2));
      q.pre.c0 = cross(q.pre.c90, q.pre.nnormal);
      return &q.super;
    }
    if (l.type == OC_LIGHT_SPHERE) { // lights/PointLight.cpp:22-62
      s.points.emplace_back();
      PointLight &pl = s.points.back();
      pl.super.isVisible = l.isVisible != 0; pl.super.instance = &s.identityInstance;
      pl.radiance = v3(l.radiance); pl.intensity = v3(l.intensity); pl.radius = l.radius;
      pl.pre.position = v3(l.position);
      return &pl.super;
    }
    if (l.type == OC_LIGHT_SPOT) { // lights/SpotLight.cpp:22-62
      s.spots.emplace_back();
      SpotLight &sl = s.spots.back();
      sl.super.isVisible = l.isVisible != 0; sl.super.instance = &s.identityInstance;
      sl.pre.position = v3(l.position); sl.pre.c0 = v3(l.fx); sl.pre.c90 = v3(l.fy); sl.pre.direction = v3(l.fz);
      sl.radiance = v3(l.radiance); sl.intensity = v3(l.intensity);
      sl.cosAngleMax = l.cosAngle; sl.cosAngleScale = l.cosAngleScale;
      sl.radius = l.radius; sl.innerRadius = l.innerRadius; sl.areaPdf = l.ppdf;
      return &sl.super;
    }
    if (l.type == OC_LIGHT_DISTANT) {
      s.dirs.emplace_back();
      DirectionalLight &d = s.dirs.back();
      d.super.isVisible = l.isVisible != 0; d.super.instance = &s.identityInstance;
      d.frame = linear3f(v3(l.fx), v3(l.fy), v3(l.fz));
      d.irradiance = v3(l.radiance); d.cosAngle = l.cosAngle; d.pdf = l.pdf;
      return &d.super;
    }
    if (l.type == OC_LIGHT_CYLINDER) { // lights/CylinderLight.cpp:22-55 (committed values)
      s.cylinders.emplace_back();
      CylinderLight &c = s.cylinders.back();
      c.super.isVisible = l.isVisible != 0; c.super.instance = &s.identityInstance;
      c.radiance = v3(l.radiance); c.radius = l.radius;
      c.pre.position0 = v3(l.position); c.pre.position1 = v3(l.edge1);
      return &c.super;
    }
    if (l.type == OC_LIGHT_HDRI) { // lights/HDRILight.cpp:20-97: the reference builds the importance distribution of the map itself
      s.hdris.emplace_back();
      HDRILight &h = s.hdris.back();
      h.super.isVisible = l.isVisible != 0; h.super.instance = &s.identityInstance;
      h.radianceScale = v3(l.radiance);
      h.light2world = linear3f(v3(l.fx), v3(l.fy), v3(l.fz));
      h.world2light = rcp(h.light2world);
      h.map = makeTexture(s.textures, l.hdriMap);
      h.rcpSize = make_vec2f(1.f / h.map->size.x, 1.f / h.map->size.y);
      s.hdriDists.emplace_back();
      Distribution2D &d = s.hdriDists.back();
      d.size = h.map->size; d.rcpSize = make_vec2f(1.f / d.size.x, 1.f / d.size.y);
      s.hdriCdfs.emplace_back((size_t)d.size.x * d.size.y); d.cdf_x = s.hdriCdfs.back().data();
      s.hdriCdfs.emplace_back((size_t)d.size.y); d.cdf_y = s.hdriCdfs.back().data();
      HDRILight_initDistribution(h.map, &d);
      h.distribution = &d;
      return &h.super;
    }
    s.ambients.emplace_back();
    AmbientLight &a = s.ambients.back();
    a.super.isVisible = l.isVisible != 0; a.super.instance = &s.identityInstance;
    a.radiance = v3(l.radiance);
    return &a.super;
  };
  // render/pathtracer/PathTracerData.cpp:32-82
  for (uint32_t i = 0; i < w->numLights; i++) s.ptLights.push_back(makeLight(w->lights[i]));
  s.ptData.lights = s.ptLights.empty() ? nullptr : s.ptLights.data();
  s.ptData.numLights = w->numLights; s.ptData.numGeoLights = 0; s.ptData.lightsCDF = (float *)w->lightsCDF;
  // render/scivis/SciVisData.cpp:17-105
  for (uint32_t i = 0; i < w->numScivisLights; i++) s.svLights.push_back(makeLight(w->scivisLights[i]));
  s.svData.lights = s.svLights.empty() ? nullptr : s.svLights.data();
  s.svData.numLights = w->numScivisLights; s.svData.numLightsVisibleOnly = w->numScivisLightsVisibleOnly;
  s.svData.aoColorPi = v3(w->aoColorPi);

  // common/World.cpp:70-188
  s.world.instances = s.instancePtrs.empty() ? nullptr : s.instancePtrs.data();
  s.world.embreeSceneHandleGeometries = anyGeometry ? &s.geomScene : nullptr;
#ifdef OSPRAY_ENABLE_VOLUMES
  s.world.embreeSceneHandleVolumes = anyVolume ? &s.volScene : nullptr;
#endif
  s.world.scivisData = &s.svData; s.world.pathtracerData = &s.ptData;
  (void)ren; (void)anyVolume;
}

// PathTracerData::generateGeometryLights (render/pathtracer/PathTracerData.cpp:87-160) with the reference's own
// GeometricModel_gatherEmissivePrimIDs: proxy lights in front of the light list, then the uniform CDF over all of them
void addGeometryLights(Scene &s, Renderer *renderer)
{
  std::vector<Light *> geo;
  for (Instance &inst : s.instances) {
    Group *grp = inst.group;
    for (int32 g = 0; g < grp->numGeometricModels; g++) {
      GeometricModel *model = grp->geometricModels[g];
      if (model->numMaterials == 0) continue;
      bool emissive = false; // GeometricModel::hasEmissiveMaterials
      for (int32 k = 0; k < model->numMaterials; k++) {
        Material *m = model->material ? model->material[k]
                                      : (renderer->material ? renderer->material[std::min<uint32_t>(model->materialID[k], renderer->numMaterials - 1)] : nullptr);
        if (m && m->isEmissive) emissive = true;
      }
      if (!emissive || !model->geom->getAreas) continue;
      s.geoPrimIDs.emplace_back(model->geom->numPrimitives); s.geoDistributions.emplace_back(model->geom->numPrimitives);
      float pdf = 0.f;
      const int n = GeometricModel_gatherEmissivePrimIDs(model, renderer, &inst, s.geoPrimIDs.back().data(), s.geoDistributions.back().data(), pdf);
      if (!n) continue;
      s.geoLights.emplace_back();
      GeometryLight &gl = s.geoLights.back();
      gl.super.instance = &inst; gl.super.isVisible = true;
      gl.model = model; gl.numPrimitives = n; gl.pdf = pdf;
      gl.primIDs = s.geoPrimIDs.back().data(); gl.distribution = s.geoDistributions.back().data();
      geo.push_back(&gl.super);
    }
  }
  if (geo.empty()) return;
  geo.insert(geo.end(), s.ptLights.begin(), s.ptLights.end());
  s.ptLights = geo;
  s.geoCDF.assign(s.ptLights.size(), 1.f);
  Distribution1D_create((int)s.geoCDF.size(), s.geoCDF.data());
  s.ptData.lights = s.ptLights.data(); s.ptData.numLights = (uint32_t)s.ptLights.size();
  s.ptData.numGeoLights = (uint32_t)s.geoLights.size(); s.ptData.lightsCDF = s.geoCDF.data();
}

struct RendererSet
{
  PathTracer pt; SciVis sv; AORenderer ao; DebugRenderer dbg;
  PixelFilter pf; LUTPixelFilter lut; Distribution2D dist;
  std::vector<Material *> materials;
  std::deque<Texture2D> textures;
};

Renderer *buildRenderer(RendererSet &r, Scene &s, const OcRenderer *o)
{
  Renderer *base = o->type == OC_RENDERER_PATHTRACER ? &r.pt.super
      : o->type == OC_RENDERER_SCIVIS ? &r.sv.super
      : o->type == OC_RENDERER_DEBUG  ? &r.dbg.super
                                      : &r.ao.super;
  // render/Renderer.cpp:38-92
  base->spp = o->spp; base->bgColor = make_vec4f(o->bgColor[0], o->bgColor[1], o->bgColor[2], o->bgColor[3]);
  base->maxDepth = o->maxDepth; base->minContribution = o->minContribution;
  for (uint32_t i = 0; i < o->numMaterials; i++) r.materials.push_back(s.allMaterials[o->materials[i]]);
  base->numMaterials = o->numMaterials; base->material = r.materials.empty() ? nullptr : r.materials.data();
  // render/Renderer.cpp:146-178 + pf/PixelFilter.cpp:23-69
  if (o->pfType == OC_PF_LUT) {
    r.dist.size = make_vec2i(o->pfLutW, o->pfLutH); r.dist.rcpSize = make_vec2f(1.f / o->pfLutW, 1.f / o->pfLutH);
    r.dist.cdf_x = (float *)o->pfCdfX; r.dist.cdf_y = (float *)o->pfCdfY;
    r.lut.super.width = o->pfWidth; r.lut.distribution = &r.dist;
    base->pixelFilter = &r.lut.super;
  } else {
    r.pf.type = o->pfType == OC_PF_POINT ? PIXEL_FILTER_TYPE_POINT : PIXEL_FILTER_TYPE_BOX;
    r.pf.width = o->pfWidth;
    base->pixelFilter = &r.pf;
  }
  base->mipBiasFactor = o->mipBiasFactor; // Renderer.cpp:91
  base->backplate = makeTexture(r.textures, o->backplate); base->maxDepthTexture = makeTexture(r.textures, o->maxDepthTexture); // :61-80
  if (o->type == OC_RENDERER_PATHTRACER) { // render/pathtracer/PathTracer.cpp:29-51
    r.pt.rouletteDepth = o->rouletteDepth; r.pt.maxScatteringEvents = o->maxScatteringEvents; r.pt.maxRadiance = o->maxRadiance;
    r.pt.numLightSamples = o->numLightSamples; r.pt.limitIndirectLightSamples = o->limitIndirectLightSamples != 0;
    r.pt.backgroundRefraction = o->backgroundRefraction != 0;
  } else if (o->type == OC_RENDERER_SCIVIS) { // render/scivis/SciVis.cpp:27-36
    r.sv.shadowsEnabled = o->shadowsEnabled != 0; r.sv.visibleLights = o->visibleLights != 0; r.sv.aoSamples = o->aoSamples;
    r.sv.aoRadius = o->aoRadius; r.sv.volumeSamplingRate = o->volumeSamplingRate;
  } else if (o->type == OC_RENDERER_DEBUG) { // render/debug/DebugRenderer.cpp:69-75 (OC_DEBUG_* == ispc::DebugRendererType)
    r.dbg.type = (DebugRendererType)o->debugMethod;
  } else { // render/ao/AORenderer.cpp:26-35
    r.ao.aoSamples = o->aoSamples; r.ao.aoRadius = o->aoRadius; r.ao.aoIntensity = o->aoIntensity;
    r.ao.volumeSamplingRate = o->volumeSamplingRate;
  }
  return base;
}

} // namespace

// Same contract as oracle_render_frame (oracle.cpp): renders one frame of `fb` (frameID already advanced) into the
// HOST buffers of `fb`.  Returns 0, or -1 when the image size is not a multiple of the 8x8 render task (the scalar
// render task of the reference does not clip partial tasks, PathTracer.ispc:232-236).
REF_API int ospref_render_frame(const OcWorld *w, const OcCamera *cam, const OcRenderer *ren, const OcFrameBuffer *fb, int numThreads)
{
  if (fb->sizeX % 8 || fb->sizeY % 8) return -1;
  auto s = std::make_unique<Scene>();
  buildScene(*s, w, ren);
  auto rs = std::make_unique<RendererSet>();
  Renderer *renderer = buildRenderer(*rs, *s, ren);
  if (ren->type == OC_RENDERER_PATHTRACER && ren->geometryLights) addGeometryLights(*s, renderer);

  // camera/OrthographicCamera.cpp:25-49 (committed values)
  OrthographicCamera oc;
  oc.super.direction = v3(cam->direction); oc.super.nearClip = cam->nearClip;
  oc.super.subImage = box2f(make_vec2f(cam->subImage[0], cam->subImage[1]), make_vec2f(cam->subImage[2], cam->subImage[3]));
  oc.super.shutter = box1f(0.5f, 0.5f);
  oc.dir = v3(cam->dir_00); oc.org = v3(cam->org); oc.du_size = v3(cam->du_size); oc.dv_up = v3(cam->dv_up); oc.height = cam->imgPlaneSizeY;
  // camera/PerspectiveCamera.cpp:26-117 (committed values)
  PerspectiveCamera pc;
  pc.super.direction = v3(cam->direction); pc.super.nearClip = cam->nearClip;
  pc.super.subImage = box2f(make_vec2f(cam->subImage[0], cam->subImage[1]), make_vec2f(cam->subImage[2], cam->subImage[3]));
  pc.super.shutter = box1f(0.5f, 0.5f);
  pc.super.needLensSample = cam->needLensSample != 0; pc.super.needTimeSample = cam->needTimeSample != 0;
  pc.org = v3(cam->org); pc.dir_00 = v3(cam->dir_00); pc.du_size = v3(cam->du_size); pc.dv_up = v3(cam->dv_up);
  pc.imgPlaneSize = make_vec2f(cam->imgPlaneSizeY * cam->aspect, cam->imgPlaneSizeY);
  pc.scaledAperture = cam->scaledAperture; pc.aspect = cam->aspect;
  pc.stereoMode = (OSPStereoMode)cam->stereoMode; pc.ipd_offset = v3(cam->ipdOffset); // PerspectiveCamera.cpp:61,85-100

  // fb/LocalFB.cpp:49-180, fb/FrameBuffer.cpp:31-137
  LocalFB lfb;
  lfb.super.size = make_vec2i(fb->sizeX, fb->sizeY);
  lfb.super.rcpSize = make_vec2f(1.f / fb->sizeX, 1.f / fb->sizeY);
  lfb.super.renderTaskSize = make_vec2i(8, 8);
  lfb.super.frameID = fb->frameID; lfb.super.targetFrames = fb->targetFrames; lfb.super.channels = fb->channels;
  lfb.super.projectedDepth = fb->projectedDepth != 0; lfb.super.accumulateVariance = fb->accumulateVariance != 0;
  lfb.colorBuffer = (vec4f *)fb->color; lfb.varianceBuffer = (vec4f *)fb->variance; lfb.depthBuffer = fb->depth;
  lfb.normalBuffer = (vec3f *)fb->normal; lfb.firstNormalBuffer = (vec3f *)fb->firstNormal;
  lfb.positionBuffer = (vec3f *)fb->position; lfb.albedoBuffer = (vec3f *)fb->albedo;
  lfb.numRenderTasks = make_vec2i(fb->sizeX / 8, fb->sizeY / 8);
  lfb.primitiveIDBuffer = fb->primID; lfb.objectIDBuffer = fb->objID; lfb.instanceIDBuffer = fb->instID;

  // task list: every 8x8 task inside a 64x64 tile this rank owns (modules/multiDevice/MultiDevice.cpp:403-416)
  std::vector<uint32_t> taskIDs;
  const int tilesX = (fb->sizeX + 63) / 64;
  const uint32_t numRanks = fb->numRanks ? fb->numRanks : 1;
  for (int ty = 0; ty < lfb.numRenderTasks.y; ty++)
    for (int tx = 0; tx < lfb.numRenderTasks.x; tx++) {
      const uint32_t tile = (uint32_t)((ty / 8) * tilesX + tx / 8);
      const uint32_t task = (uint32_t)(ty * lfb.numRenderTasks.x + tx);
      // LocalFrameBuffer::getRenderTaskIDs (fb/LocalFB.cpp:238-261): adaptive selection of the active tasks
      if (fb->taskError && !(fb->taskError[task] > fb->errorThreshold)) continue;
      if (tile % numRanks == fb->rank) taskIDs.push_back(task);
    }

  sycl::queue q; q.numThreads = numThreads > 0 ? numThreads : (int)std::max(1u, std::thread::hardware_concurrency());
  sycl::event ev;
  const ospray::FeatureFlags ff = allFeatures();
  const vec3ui dims = make_vec3ui(8, 8, (uint32_t)taskIDs.size());
  PanoramicCamera pano; // camera/PanoramicCamera.cpp:29-53 (committed values)
  if (cam->type == OC_CAMERA_PANORAMIC) {
    pano.super.direction = v3(cam->direction); pano.super.nearClip = cam->nearClip;
    pano.super.subImage = box2f(make_vec2f(cam->subImage[0], cam->subImage[1]), make_vec2f(cam->subImage[2], cam->subImage[3]));
    pano.super.shutter = box1f(0.5f, 0.5f);
    pano.org = v3(cam->org);
    pano.frame = linear3f(v3(cam->du_size), v3(cam->dv_up), v3(cam->dir_00));
    pano.stereoMode = (int)cam->stereoMode; pano.ipd_offset = cam->ipdOffset.x;
  }
  Camera *camera = cam->type == OC_CAMERA_PANORAMIC ? &pano.super : cam->type == OC_CAMERA_ORTHOGRAPHIC ? &oc.super : &pc.super;
  uint64_t before[2]; ospref_ray_counters(before, 0);
  if (ren->type == OC_RENDERER_PATHTRACER)
    PathTracer_renderTaskLauncher(&q, &ev, dims, renderer, &lfb.super, camera, &s->world, taskIDs.data(), ff);
  else if (ren->type == OC_RENDERER_SCIVIS)
    SciVis_renderTaskLauncher(&q, &ev, dims, renderer, &lfb.super, camera, &s->world, taskIDs.data(), ff);
  else if (ren->type == OC_RENDERER_DEBUG)
    DebugRenderer_renderTaskLauncher(&q, &ev, dims, renderer, &lfb.super, camera, &s->world, taskIDs.data(), ff);
  else
    AORenderer_renderTaskLauncher(&q, &ev, dims, renderer, &lfb.super, camera, &s->world, taskIDs.data(), ff);
  if (fb->rayCounters) {
    uint64_t after[2]; ospref_ray_counters(after, 0);
    // (the two debug methods that trace nothing hand no primary ray to the scene)
    const bool noRays = ren->type == OC_RENDERER_DEBUG && (ren->debugMethod == OC_DEBUG_TEST_FRAME || ren->debugMethod == OC_DEBUG_RAY_DIR);
    fb->rayCounters[0] += noRays ? 0ull : (unsigned long long)taskIDs.size() * 64ull * (unsigned long long)ren->spp;
    const unsigned long long total = (after[0] - before[0]) + (after[1] - before[1]);
    fb->rayCounters[1] += (after[0] - before[0]) - std::min<unsigned long long>(after[0] - before[0], taskIDs.size() * 64ull * ren->spp);
    fb->rayCounters[2] += after[1] - before[1];
    (void)total;
  }
  return 0;
}

// get4f(TextureParam, dg) of the reference (texture/TextureParam.ispc:10-22 -> Texture_dispatch_get -> Texture2D_get) for n
// (st, pixelFootprint) pairs; `param` and its texture hold HOST pointers
namespace ispc { vec4f get4f(const TextureParam &tp, const DifferentialGeometry &dg); }
REF_API void ospref_texture_sample(const OcTextureParam *param, const float *st, const float *footprint, uint32_t n, float *out)
{
  std::deque<Texture2D> store;
  const TextureParam tp = makeTextureParam(store, *param);
  for (uint32_t i = 0; i < n; i++) {
    DifferentialGeometry dg;
    dg.st = make_vec2f(st[2 * i], st[2 * i + 1]);
    dg.pixelFootprint = footprint[i];
    dg.lP = make_vec3f(0.f);
    const vec4f v = get4f(tp, dg);
    out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
  }
}

// ---- known-answer vectors straight from the reference's sampler sources (math/random.ih, sobol.ih, halton.ih,
// render/util.ih); same layout as oracle_sampler_kat ---------------------------------------------------------
REF_API void ospref_sampler_kat(int kind, uint32_t seed, uint32_t first, uint32_t count, float *out)
{
  for (uint32_t k = 0; k < count; k++) {
    const uint32_t i = first + k;
    if (kind == 0) {
      float *o = out + 12 * (size_t)k;
      LDSampler ld;
      LDSampler_init(ld, seed, i, i, 16, 16);
      vec3ui ss;
      const vec2f p = LDSampler_getNext5Samples(ld, ss);
      const vec2f l = LDSampler_finalizeDim23(ld, ss);
      o[0] = p.x; o[1] = p.y; o[2] = l.x; o[3] = l.y; o[4] = LDSampler_finalizeDim4(ld, ss);
      LDSampler_nextGroup(ld);
      float s1;
      const vec2f ls = LDSampler_get3LightSamples(ld, 0, false, s1);
      o[5] = ls.x; o[6] = ls.y; o[7] = s1;
      float s2; uint32 u3;
      const vec2f b = LDSampler_getNext4Samples(ld, s2, u3);
      o[8] = b.x; o[9] = b.y; o[10] = s2; o[11] = LDSampler_finalizeDim3(ld, u3);
    } else if (kind == 1) {
      float *o = out + 4 * (size_t)k;
      o[0] = Halton_sample2(i); o[1] = CranleyPattersonRotation(Halton_sample3(i), 1.f / 6.f);
      o[2] = PermHalton_sample5(i); o[3] = PermHalton_sample7(i);
    } else if (kind == 2) {
      float *o = out + 4 * (size_t)k;
      RandomSampler rs;
      RandomSampler_init(&rs, seed, i);
      for (int j = 0; j < 4; j++) o[j] = RandomSampler_getFloat(&rs);
    } else if (kind == 3) {
      uint32 v0 = seed, v1 = i;
      tea8(v0, v1);
      out[2 * (size_t)k] = to_float_unorm(v0); out[2 * (size_t)k + 1] = to_float_unorm(v1);
    } else {
      const uint32 v = OwenScramble4(interleaveZOrder16(i, seed), 0x49c5f37eu);
      memcpy(out + k, &v, 4);
    }
  }
}

// default (Gaussian) pixel filter LUT built by the reference's LUTPixelFilter_buildLUT (pf/LUTPixelFilter.ispc:120-142)
// (which also runs Distribution1D_create per row and Distribution2D_init, math/Distribution2D.ispc)
REF_API void ospref_gauss_pixelfilter_lut(float width, int lut, float *cdfX, float *cdfY)
{
  LUTPixelFilter f; Distribution2D d;
  d.size = make_vec2i(lut, lut); d.rcpSize = make_vec2f(1.f / lut, 1.f / lut); d.cdf_x = cdfX; d.cdf_y = cdfY;
  f.super.width = width; f.distribution = &d; f.lutFilterType = LUT_PIXEL_FILTER_TYPE_GAUSSIAN;
  LUTPixelFilter_buildLUT(&f); // rows via Distribution1D_create, then Distribution2D_init builds the marginal
}

// frame ops through the reference's own kernels and launchers (fb/frame_ops/*.ispc, DEFINE_FRAMEOP_KERNEL_LAUNCHER)
REF_API void ospref_fb_color_convert(const OcVec4f *src, void *out, uint32_t n, uint32_t format)
{
  LiveColorConversion l;
  l.super = FrameBufferView(make_vec2ui(n, 1), (const vec4f *)src);
  l.targetColorFormat = (OSPFrameBufferFormat)format; l.convBuffer = out;
  sycl::queue q; sycl::event ev;
  ColorConversion_kernelLauncher(&q, &ev, make_vec2ui(n, 1), &l.super);
}
REF_API void ospref_fb_tonemap(const OcVec4f *src, OcVec4f *out, uint32_t n, const OcToneMapper *tm)
{
  LiveToneMapper l;
  l.super = FrameBufferView(make_vec2ui(n, 1), (const vec4f *)src);
  l.super.colorBufferOutput = (vec4f *)out;
  l.exposure = tm->exposure; l.a = tm->a; l.b = tm->b; l.c = tm->c; l.d = tm->d; l.acesColor = tm->acesColor != 0;
  sycl::queue q; sycl::event ev;
  ToneMapper_kernelLauncher(&q, &ev, make_vec2ui(n, 1), &l.super);
}
REF_API void ospref_fb_variance(const OcVec4f *color, const OcVec4f *variance, int sizeX, int sizeY, int taskSize, float *taskError)
{
  const vec2ui tasks = make_vec2ui((sizeX + taskSize - 1) / taskSize, (sizeY + taskSize - 1) / taskSize);
  LiveVariance l;
  l.super = FrameBufferView(make_vec2ui(sizeX, sizeY), (const vec4f *)color);
  l.super.viewDims = tasks; // fb/frame_ops/Variance.cpp: the view is the render-task grid
  l.rtSize = make_vec2ui(taskSize, taskSize); l.varianceBuffer = (const vec4f *)variance; l.taskVarianceBuffer = taskError;
  sycl::queue q; sycl::event ev;
  Variance_kernelLauncher(&q, &ev, tasks, &l.super);
}
// LocalFB_accumulateSample (fb/LocalFB.ispc:14-78) applied to an externally produced sample image
REF_API void ospref_fb_accumulate(const OcFrameBuffer *fb, const OcVec4f *sc, const float *sz)
{
  LocalFB lfb;
  lfb.super.size = make_vec2i(fb->sizeX, fb->sizeY); lfb.super.frameID = fb->frameID;
  lfb.super.accumulateVariance = fb->accumulateVariance != 0 && fb->variance;
  lfb.colorBuffer = (vec4f *)fb->color; lfb.varianceBuffer = (vec4f *)fb->variance; lfb.depthBuffer = sz ? fb->depth : nullptr;
  RenderTaskDesc desc;
  for (int y = 0; y < fb->sizeY; y++)
    for (int x = 0; x < fb->sizeX; x++) {
      const size_t i = (size_t)y * fb->sizeX + x;
      ScreenSample s = make_ScreenSample_zero();
      s.sampleID.x = x; s.sampleID.y = y;
      s.rgb = make_vec3f(sc[i].x, sc[i].y, sc[i].z); s.alpha = sc[i].w; s.z = sz ? sz[i] : inf;
      LocalFB_accumulateSample(&lfb.super, s, desc);
    }
}
REF_API int ospref_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

extern "C" uint32_t ospref_albedo_tables(float *out, uint32_t capacity)
{
  // precomputeMicrofacetAlbedoTables (render/bsdfs/MicrofacetAlbedoTables.ispc:162-172) in the order of ospcuda.h's block:
  // albedo dir / avg, dielectric dir / avg, 1/eta dir / avg, reflection dir / avg, sheen
  const size_t a = MICROFACET_ALBEDO_TABLE_SIZE, d = MICROFACET_DIELECTRIC_ALBEDO_TABLE_SIZE, sh = MICROFACET_SHEEN_ALBEDO_TABLE_SIZE;
  const uint32_t n = (uint32_t)(a * a + a + 3 * (d * d * d + d * d) + sh * sh);
  if (out && capacity >= n) std::memcpy(out, refAlbedoTables()->albedo_dir, n * sizeof(float));
  return n;
}

