// Frame kernels: pathtracer, scivis and ao renderers with the framebuffer accumulation fused
// into the epilogue. One thread per pixel, one 8x8 render task per CTA-half, tiles (64x64)
// assigned to ranks round-robin like modules/multiDevice/MultiDevice.cpp:403-416.
//
// reference: render/pathtracer/PathTracer.ispc:75-249, PathSampler.ispc:109-415,
// NextEventEstimation.ispc:25-160, TransparentShadow.ispc:18-85, VirtualLight.ispc:38-107,
// PathTracerUtil.ih:30-46; render/RendererRenderTaskFn.inl:5-88; render/scivis/SciVis.ispc:24-281,
// scivis/surfaces.ispc:17-128, scivis/surfaces.ih:49-92, scivis/lightAlpha.ispc:19-122,
// scivis/volumes.ispc:302-439; render/ao/AORenderer.ispc:23-149, ao/surfaces.ispc:18-74,
// ao/volumes.ispc, render/util.ispc:12-54; render/ScreenSample.ih:12-92; fb/LocalFB.ispc:14-78.
#include <algorithm>
#include <utility>
#include <vector>
#include "runtime.h"
#include <cfloat>
#include "volume.cuh"

namespace oc {

struct ScreenSample
{
  vec3f rgb; float alpha; float z;
  vec3f position, firstNormal, normal, albedo;
  uint32_t primID, geomID, instID;
};
__device__ __forceinline__ void ScreenSample_zero(ScreenSample &s)
{
  s.rgb = make_vec3f(0.f); s.alpha = 0.f; s.z = INFINITY;
  s.position = s.firstNormal = s.normal = s.albedo = make_vec3f(0.f);
  s.primID = s.geomID = s.instID = OC_INVALID_ID;
}
__device__ __forceinline__ void ScreenSample_accumulate(ScreenSample &a, const ScreenSample &s)
{
  a.rgb = a.rgb + s.rgb; a.alpha = a.alpha + s.alpha;
  a.position = a.position + s.position; a.firstNormal = a.firstNormal + s.firstNormal;
  a.normal = a.normal + s.normal; a.albedo = a.albedo + s.albedo;
  if (fabsf(s.z) < fabsf(a.z)) { a.z = s.z; a.primID = s.primID; a.geomID = s.geomID; a.instID = s.instID; }
}
__device__ __forceinline__ void ScreenSample_normalize(ScreenSample &s, float rspp)
{
  s.rgb = s.rgb * rspp; s.alpha *= rspp; s.position = s.position * rspp;
  s.firstNormal = s.firstNormal * rspp; s.normal = s.normal * rspp; s.albedo = s.albedo * rspp;
}

struct Counters { uint32_t primary, secondary, shadow; };

struct Ctx
{
  const OcWorld *world;
  const OcRenderer &ren;
  const OcCamera &cam;
  const OcFrameBuffer &fb;
  Stack stack;
  Counters cnt;
  RayCone cone;      // frames with textures: ray cone of the sample being rendered (ScreenSample::rayCone)
  vec2f pixelCenter; // ScreenSample::pos: where the backplate is looked up
  __device__ Ctx(const OcWorld *w, const OcRenderer &r, const OcCamera &c, const OcFrameBuffer &f, Stack s)
      : world(w), ren(r), cam(c), fb(f), stack(s) { cnt.primary = cnt.secondary = cnt.shadow = 0; }
};

__device__ __forceinline__ vec4f getBackground(const OcRenderer &r)
{
  return make_vec4f(r.bgColor[0], r.bgColor[1], r.bgColor[2], r.bgColor[3]);
}

// Renderer_getBackground / Renderer_getMaxDepth with renderer textures (render/Renderer.ispc:18-39; Texture2D.ih:20-47:
// a bare Texture2D is sampled at the finest footprint, the backplate coordinate is clamped to the texel centres)
__device__ __forceinline__ vec4f getBackgroundTex(const OcRenderer &r, vec2f screenPos)
{
  if (!r.backplate) return getBackground(r);
  const float hx = (1.f / (float)r.backplate->sizeX) * 0.5f, hy = (1.f / (float)r.backplate->sizeY) * 0.5f;
  const vec2f p = make_vec2f(fmaxf(hx, fminf(screenPos.x, 1.0f - hx)), fmaxf(hy, fminf(screenPos.y, 1.0f - hy)));
  return Texture2D_get(r.backplate, p, FLT_MIN);
}
__device__ __forceinline__ float rendererMaxDepth(const OcRenderer &r, vec2f screenPos)
{
  if (!r.maxDepthTexture) return INFINITY;
  const float d = Texture2D_get(r.maxDepthTexture, screenPos, FLT_MIN).x;
  return 0.f > d ? 0.f : d; // max(0.f, d)
}

// ===================================== path tracer =========================================
struct PathVertexLite { vec3f P, Ns; float pdf_w; uint32_t numLightSamples; };

__device__ __forceinline__ float misHeuristic(float pdf1, float pdf2)
{
  const float p = sqr(pdf1) * (1.f / (sqr(pdf1) + sqr(pdf2)));
  return pdf1 > 1e17f ? 1.0f : p;
}

template <bool TEX = false>
__device__ __forceinline__ vec3f evaluateVirtualLights(Ctx &c, bool straightPath, vec3f throughput, const PathVertexLite &last,
    bool environment, const Ray &ray)
{
  vec3f L = make_vec3f(0.f);
  const uint32_t numLights = c.world->numLights, numGeoLights = c.world->numGeoLights;
  if (numLights <= numGeoLights) return L; // no virtual lights (geometry lights come first and are hit, not evaluated)
  // (the last vertex usually IS the ray's origin: sqrtf(0) would take the square root's slow path; +0 either way)
  const vec3f toOrg = last.P - ray.org;
  const float toOrg2 = dot(toOrg, toOrg);
  const float lower = toOrg2 == 0.f ? 0.f : sqrtf(toOrg2);
  const float upper = environment ? INFINITY : length(last.P - (ray.org + ray.t * ray.dir));
  const float selectionPDF = last.numLightSamples / (float)numLights;
  for (uint32_t i = numGeoLights; i < numLights; i++) {
    const OcLight &l = c.world->lights[i];
    if (!straightPath || l.isVisible) {
      const LightEval le = (TEX && l.type == OC_LIGHT_HDRI) ? HDRILight_eval(l, ray.dir, upper)
          : (TEX && l.type == OC_LIGHT_CYLINDER) ? CylinderLight_eval(l, last.P, ray.dir, lower, upper) : Light_eval(l, last.P, last.Ns, ray.dir, lower, upper);
      if (reduce_max(le.radiance) > 0.0f) {
        float misWeight = 1.0f;
        if (selectionPDF) misWeight = misHeuristic(last.pdf_w, selectionPDF * le.pdf);
        L = L + throughput * le.radiance * misWeight;
      }
    }
  }
  return L;
}

// ============================== scivis / ao shared pieces ==================================
// Everything the scivis / ao / debug frame kernel calls, as static members of one class template: TEX = the frame references
// textures (material maps, backplate, depth limit), so hits carry texture coordinates, both tangents and the ray cone
// (camera -> c.cone -> postIntersect footprint; shadow cones per scivis/surfaces.ispc:64-67, SciVis.ispc:262-264). Frames
// without textures run the TEX = false instantiation, whose code is what it was before textures existed.
template <bool TEX>
struct Mega {
struct SciVisBSDF { vec3f diffuse, specular, transmission, albedo; float shininess, opacity; };

// scivis/surfaces.ih:49-92
static __device__ __forceinline__ SciVisBSDF evalMaterial(const DifferentialGeometry &dg)
{
  SciVisBSDF r;
  r.diffuse = make_vec3f(dg.color);
  r.specular = make_vec3f(0.f);
  r.shininess = 10.f;
  r.transmission = make_vec3f(1.f - dg.color.w);
  r.opacity = dg.color.w;
  const OcMaterial *m = dg.material;
  if (m && m->type == OC_MATERIAL_OBJ) {
    const bool tex = TEX && m->hasTextures;
    float d = m->d;
    if (tex && m->mapD.tex) d = d * texParamGet4f(m->mapD, dg.st, dg.pixelFootprint).x;
    d = d * dg.color.w;
    r.diffuse = r.diffuse * toV(m->Kd);
    if (tex && m->mapKd.tex) {
      const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint);
      r.diffuse = r.diffuse * make_vec3f(t.x, t.y, t.z);
      d *= t.w;
    }
    r.diffuse = r.diffuse * d;
    r.specular = d * toV(m->Ks);
    if (tex && m->mapKs.tex) { const vec4f t = texParamGet4f(m->mapKs, dg.st, dg.pixelFootprint); r.specular = r.specular * make_vec3f(t.x, t.y, t.z); }
    r.shininess = m->Ns;
    if (tex && m->mapNs.tex) r.shininess = r.shininess * texParamGet4f(m->mapNs, dg.st, dg.pixelFootprint).x;
    r.transmission = toV(m->Tf) * d + make_vec3f(1.f - d);
    r.opacity = d;
  }
  r.albedo = r.diffuse + r.specular;
  r.shininess *= 4.0f;
  r.diffuse = r.diffuse * OC_ONE_OVER_PI;
  r.specular = r.specular * ((r.shininess + 2.f) * OC_ONE_OVER_FOUR_PI * (1.f / (2.f - expf(r.shininess * -OC_LNTWO_OVER_TWO))));
  return r;
}

// single-volume marcher, AO flavour (ao/volumes.ispc:244-330): returns (rgb, alpha)
static __device__ __noinline__ vec4f integrateVolumeAO(Ctx &c, const VolumeHit &vh, const Ray &ray, RandomSampler &rs, float samplingRate)
{
  const OcVolume &v = *vh.vol;
  vec3f org = ray.org, dir = ray.dir;
  if (!vh.inst->isIdentity) {
    const affine3f rx = toA(vh.inst->rcp_xfm);
    org = xfmPoint(rx, ray.org); dir = xfmVector(rx, ray.dir);
  }
  vec3f color = make_vec3f(0.f);
  float alpha = 0.f;
  VolumeIterator it;
  volumeIteratorInit(it, v, org, dir, vh.t0, vh.t1);
  const float jitter = RandomSampler_getFloat(rs);
  float distance = INFINITY, iuDistance = jitter, iuLength = 0.f;
  float ivLower = 0.f, ivUpper = INFINITY, samplingStep = 0.f;
  while (alpha < .99f) {
    float emptySpace = 0.f;
    while (iuDistance > iuLength) {
      const float prevUpper = ivUpper;
      if (volumeIteratorNext(it, v, ivLower, ivUpper)) {
        emptySpace += fmaxf(ivLower - prevUpper, 0.f);
        iuDistance -= iuLength;
        samplingStep = it.nominalDeltaT / samplingRate;
        iuLength = (ivUpper - ivLower) / samplingStep;
        distance = (distance == INFINITY) ? ivLower : distance;
      } else {
        return make_vec4f(color, alpha);
      }
    }
    const float newDistance = ivLower + iuDistance * samplingStep;
    const vec3f p = org + newDistance * dir;
    const float sampleVal = volumeSample(v, p);
    if (!isnan(sampleVal)) {
      vec4f sc = transferFunctionGet(v.tf, sampleVal);
      const float dt = newDistance - distance - emptySpace;
      sc.w = 1.f - expf(-sc.w * dt * v.densityScale);
      color = color + ((1.f - alpha) * sc.w * make_vec3f(sc));
      alpha = alpha + ((1.f - alpha) * sc.w);
    }
    iuDistance += 1.f;
    distance = newDistance;
  }
  return make_vec4f(color, alpha);
}

struct SSI { vec3f shadedColor, transmission, albedo; float opacity; };

// lightAlpha (scivis/lightAlpha.ispc:19-122)
static __device__ __noinline__ vec3f lightAlpha(Ctx &c, Ray &ray, RayCone shadowCone, RandomSampler &rs, vec3f weight, float quality)
{
  vec3f alpha = make_vec3f(1.f);
  const float org_t_max = ray.t;
  const float org_width = shadowCone.width;
  while (true) {
    c.cnt.shadow++;
    traceWorldNI<false>(c.world, ray, c.stack);
    if (c.world->hasVolumes) {
      VolumeHit vh;
      if (traceVolumeRay(c.world, ray, vh)) {
        const vec4f vc = integrateVolumeAO(c, vh, ray, rs, c.ren.volumeSamplingRate * quality);
        alpha = alpha * make_vec3f(1.f - vc.w);
      }
    }
    if (hadHit(ray)) {
      DifferentialGeometry dg;
      postIntersectNI<TEX>(c.world, c.ren, dg, ray, DG_COLOR | DG_TEXCOORD, TEX ? &shadowCone : nullptr);
      const SciVisBSDF bsdf = evalMaterial(dg);
      alpha = alpha * bsdf.transmission;
      setRay(ray, ray.org, ray.dir, ray.t + dg.epsilon, org_t_max);
      shadowCone.width = org_width;
    } else
      break;
    if (luminance(alpha * weight) < c.ren.minContribution) {
      alpha = make_vec3f(0.f);
      break;
    }
  }
  return alpha;
}

// SciVis_computeAO (scivis/SciVis.ispc:224-281)
static __device__ __noinline__ vec3f SciVis_computeAO(Ctx &c, RandomSampler &rs, const DifferentialGeometry &dg, int sampleCnt, float aoRadius,
    uint32_t px, uint32_t py, int sampleIDz, float coneWidth)
{
  RayCone shadowCone;
  shadowCone.width = coneWidth; shadowCone.dwdt = 0.0f; // constant for AO (SciVis.ispc:262-264)
  const int accumID = sampleIDz * sampleCnt;
  uint32_t v0 = px, v1 = py;
  tea8(v0, v1);
  const vec2f rot = make_vec2f(to_float_unorm(v0), to_float_unorm(v1));
  vec3f hits = make_vec3f(0.f);
  const linear3f localToWorld = frame(dg.Ns);
  for (int i = 0; i < sampleCnt; i++) {
    const vec2f halton = make_vec2f(PermHalton_sample5(accumID + i), PermHalton_sample7(accumID + i));
    const vec2f r = make_vec2f(CranleyPattersonRotation(halton.x, rot.x), CranleyPattersonRotation(halton.y, rot.y));
    const vec3f ao_dir = localToWorld * cosineSampleHemisphere(r);
    if (dot(ao_dir, dg.Ns) < 0.01f) {
      hits = hits + make_vec3f(1.f);
      continue;
    }
    Ray ao_ray;
    setRay(ao_ray, dg.P, ao_dir, dg.epsilon, aoRadius);
    hits = hits + (make_vec3f(1.f) - lightAlpha(c, ao_ray, shadowCone, rs, make_vec3f(1.f), 0.1f));
  }
  return make_vec3f(1.0f) - (hits / (float)sampleCnt);
}

// scivis/surfaces.ispc:17-128
static __device__ __noinline__ SSI SciVis_computeShading(Ctx &c, const DifferentialGeometry &dg, RandomSampler &rs, vec3f inDir, uint32_t px,
    uint32_t py, int sampleIDz, float coneWidth)
{
  SSI ret;
  const SciVisBSDF bsdf = evalMaterial(dg);
  ret.albedo = bsdf.albedo;
  vec3f color = make_vec3f(0.f);
  const OcWorld *w = c.world;
  for (uint32_t i = w->numScivisLightsVisibleOnly; i < w->numScivisLights; i++) {
    const LightSample light = (TEX && w->scivisLights[i].type == OC_LIGHT_CYLINDER) ? CylinderLight_sample(w->scivisLights[i], dg, make_vec2f(0.f, 0.f))
                                                                                   : Light_sample(w->scivisLights[i], dg, make_vec2f(0.f, 0.f));
    if (reduce_max(light.weight) > 0.f) {
      const float cosNL = dot(light.dir, dg.Ns);
      if (cosNL < 0.0f) continue;
      const vec3f H = normalize(light.dir - inDir);
      const float cosNH = dot(H, dg.Ns);
      const vec3f specular = bsdf.specular * powf(cosNH, bsdf.shininess);
      const vec3f diffuse = bsdf.diffuse * cosNL;
      const vec3f light_contrib = (diffuse + specular) * light.weight;
      if (c.ren.shadowsEnabled) {
        if (luminance(light_contrib) > c.ren.minContribution) {
          vec3f P = dg.P;
          if (dot(dg.Ng, light.dir) < 0.0f) P = P - (2.f * dg.epsilon) * dg.Ng;
          Ray shadowRay;
          setRay(shadowRay, P, light.dir, 0.0f, light.dist);
          RayCone shadowCone; // connects the vertex's cone to the point on the light (surfaces.ispc:64-67)
          shadowCone.width = coneWidth; shadowCone.dwdt = -coneWidth * (1.f / light.dist);
          const vec3f la = lightAlpha(c, shadowRay, shadowCone, rs, light_contrib, 0.25f);
          color = color + la * light_contrib;
        }
      } else {
        color = color + light_contrib;
      }
    }
  }
  vec3f ao = make_vec3f(1.f);
  const vec3f aoColorPi = toV(w->aoColorPi);
  if (c.ren.aoSamples > 0 && luminance(aoColorPi) > c.ren.minContribution)
    ao = SciVis_computeAO(c, rs, dg, c.ren.aoSamples, c.ren.aoRadius, px, py, sampleIDz, coneWidth);
  color = color + bsdf.diffuse * ao * aoColorPi;
  ret.shadedColor = color;
  ret.transmission = bsdf.transmission;
  ret.opacity = bsdf.opacity;
  return ret;
}

// gradient shading of one volume sample (scivis/volumes.ispc:391-419), out of line: the marching loop of volumes without
// gradient shading (the default) keeps its registers
static __device__ __noinline__ vec4f shadeVolumeSampleSciVis(Ctx &c, const VolumeHit &vh, vec3f p, vec4f sc, float dt, RandomSampler &rs, vec3f rayDir,
    uint32_t px, uint32_t py, int sampleIDz)
{
  const OcVolume &v = *vh.vol;
  vec3f ns = volumeGradient(v, p);
  if (dot(ns, ns) > 1e-6f) {
    ns = neg(ns);
    DifferentialGeometry dg;
    dg.color = sc;
    dg.material = nullptr;
    dg.epsilon = dt / 2.f;
    const vec3f nw = vh.inst->isIdentity ? ns : xfmVector(transposed(toA(vh.inst->rcp_xfm).l), ns);
    dg.Ns = dg.Ng = normalize(nw);
    dg.P = p;
    dg.dPds = make_vec3f(1.f, 0.f, 0.f);
    dg.primID = 0; dg.objID = dg.instID = OC_INVALID_ID;
    const SSI shading = SciVis_computeShading(c, dg, rs, rayDir, px, py, sampleIDz, 0.f); // no coneWidth used (volumes.ispc:414)
    const vec4f shadedColor = make_vec4f(shading.shadedColor, 1.f - luminance(shading.transmission));
    sc = lerp4(v.gradientShadingScale, sc, shadedColor);
  }
  return sc;
}

// single-volume marcher, SciVis flavour with gradient shading (scivis/volumes.ispc:302-439):
// returns (rgb, transmission)
// volT / volHit follow the reference CPU form (sampleAllVolumes, scivis/volumes.ispc:160-169): the
// first sample in front of the surface hit that is not fully transparent defines depth and IDs.
struct VolumeMarch { vec4f color; float volT; bool volHit; }; // handed back by value: references would pin volT / volHit to local memory
static __device__ __noinline__ VolumeMarch integrateVolumeSciVis(Ctx &c, const VolumeHit &vh, const Ray &ray, RandomSampler &rs, float samplingRate,
    uint32_t px, uint32_t py, int sampleIDz, float volT)
{
  bool volHit = false;
  const OcVolume &v = *vh.vol;
  vec3f org = ray.org, dir = ray.dir;
  if (!vh.inst->isIdentity) {
    const affine3f rx = toA(vh.inst->rcp_xfm);
    org = xfmPoint(rx, ray.org); dir = xfmVector(rx, ray.dir);
  }
  vec3f color = make_vec3f(0.f);
  float transmission = 1.f;
  VolumeIterator it;
  volumeIteratorInit(it, v, org, dir, vh.t0, vh.t1);
  const float jitter = RandomSampler_getFloat(rs);
  float distance = INFINITY, iuDistance = jitter, iuLength = 0.f;
  float ivLower = 0.f, ivUpper = INFINITY, samplingStep = 0.f;
  while (transmission > c.ren.minContribution) {
    float emptySpace = 0.f;
    while (iuDistance > iuLength) {
      const float prevUpper = ivUpper;
      if (volumeIteratorNext(it, v, ivLower, ivUpper)) {
        emptySpace += fmaxf(ivLower - prevUpper, 0.f);
        iuDistance -= iuLength;
        samplingStep = it.nominalDeltaT / samplingRate;
        iuLength = (ivUpper - ivLower) / samplingStep;
        distance = (distance == INFINITY) ? ivLower : distance;
      } else {
        return VolumeMarch{make_vec4f(color, transmission), volT, volHit};
      }
    }
    const float newDistance = ivLower + iuDistance * samplingStep;
    const vec3f p = org + newDistance * dir;
    const float sampleVal = volumeSample(v, p);
    if (!isnan(sampleVal)) {
      vec4f sc = transferFunctionGet(v.tf, sampleVal);
      const float dt = newDistance - distance - emptySpace;
      if (v.gradientShadingScale > 0.f && sc.w > 0.0f) sc = shadeVolumeSampleSciVis(c, vh, p, sc, dt, rs, ray.dir, px, py, sampleIDz);
      sc.w = expf(-sc.w * dt * v.densityScale);
      if (newDistance < volT && sc.w < 1.0f - c.ren.minContribution) { volT = newDistance; volHit = true; }
      color = color + transmission * (1.f - sc.w) * make_vec3f(sc);
      transmission *= sc.w;
    }
    iuDistance += 1.f;
    distance = newDistance;
  }
  return VolumeMarch{make_vec4f(color, transmission), volT, volHit};
}

// visible lights (scivis/surfaces.ispc:130-150)
static __device__ vec3f evaluateLightsSciVis(Ctx &c, const Ray &ray)
{
  vec3f color = make_vec3f(0.f);
  const OcWorld *w = c.world;
  for (uint32_t i = 0; i < w->numScivisLights; i++) {
    const OcLight &l = w->scivisLights[i];
    if (!l.isVisible) continue;
    const LightEval le = (TEX && l.type == OC_LIGHT_HDRI) ? HDRILight_eval(l, ray.dir, ray.t)
        : (TEX && l.type == OC_LIGHT_CYLINDER) ? CylinderLight_eval(l, ray.org, ray.dir, ray.t0, ray.t) : Light_eval(l, ray.org, make_vec3f(0.f), ray.dir, ray.t0, ray.t);
    color = color + le.radiance;
  }
  return color;
}

// SciVis_renderSample (scivis/SciVis.ispc:24-222)
static __device__ void SciVis_renderSample(Ctx &c, Ray &ray, ScreenSample &sample, uint32_t px, uint32_t py, int sampleIDz)
{
  const OcRenderer &ren = c.ren;
  RandomSampler rs;
  RandomSampler_init(rs, px + (uint32_t)c.fb.sizeX * py, (uint32_t)sampleIDz);
  bool firstHit = true, allHitsFullyTransmissive = true;
  const float originalRayTFar = ray.t;
  vec3f outColor = make_vec3f(0.f), outTransmission = make_vec3f(1.f);
  bool primary = true;
  while (true) {
    if (primary) c.cnt.primary++; else c.cnt.secondary++;
    primary = false;
    traceWorldNI<false>(c.world, ray, c.stack);
    float volumeT = ray.t;
    if (c.world->hasVolumes) {
      VolumeHit vh;
      if (traceVolumeRay(c.world, ray, vh)) {
        const VolumeMarch vm = integrateVolumeSciVis(c, vh, ray, rs, ren.volumeSamplingRate, px, py, sampleIDz, volumeT);
        const vec4f volumeColor = vm.color;
        const bool volHit = vm.volHit;
        volumeT = vm.volT;
        outColor = outColor + outTransmission * make_vec3f(volumeColor);
        outTransmission = outTransmission * volumeColor.w;
        if (volHit && sample.geomID == OC_INVALID_ID) {
          sample.primID = 0;
          sample.instID = (vh.inst->userID == OC_INVALID_ID) ? vh.instID : vh.inst->userID;
          sample.geomID = (vh.vol->userID == OC_INVALID_ID) ? vh.geomID : vh.vol->userID;
        }
      }
    }
    if (ren.visibleLights)
      outColor = outColor + outTransmission * evaluateLightsSciVis(c, ray);
    if (firstHit || allHitsFullyTransmissive) {
      sample.z = fminf(ray.t, volumeT);
      if (c.fb.projectedDepth) sample.z *= dot(ray.dir, toV(c.cam.direction));
    }
    if (hadHit(ray)) {
      DifferentialGeometry dg;
      postIntersectNI<TEX>(c.world, ren, dg, ray, DG_NG | DG_NS | DG_NORMALIZE | DG_FACEFORWARD | DG_COLOR | DG_TANGENTS | DG_TEXCOORD, TEX ? &c.cone : nullptr);
      const SSI sh = SciVis_computeShading(c, dg, rs, ray.dir, px, py, sampleIDz, c.cone.width);
      if (firstHit || allHitsFullyTransmissive) {
        sample.position = dg.P; sample.firstNormal = dg.Ns; sample.normal = dg.Ns; sample.albedo = sh.albedo;
        if (sample.geomID == OC_INVALID_ID) { sample.instID = dg.instID; sample.geomID = dg.objID; sample.primID = dg.primID; }
        firstHit = false;
      }
      allHitsFullyTransmissive = allHitsFullyTransmissive && sh.opacity < ren.minContribution;
      outColor = outColor + outTransmission * sh.shadedColor;
      outTransmission = outTransmission * sh.transmission;
      if (luminance(outTransmission) < ren.minContribution) {
        outTransmission = make_vec3f(0.f);
        break;
      }
      setRay(ray, ray.org, ray.dir, ray.t + dg.epsilon, originalRayTFar);
    } else {
      const vec4f bg = TEX ? getBackgroundTex(ren, c.pixelCenter) : getBackground(ren);
      outColor = outColor + outTransmission * make_vec3f(bg);
      outTransmission = outTransmission * (1.f - bg.w);
      break;
    }
  }
  sample.rgb = outColor;
  sample.alpha = 1.f - luminance(outTransmission);
}

// computeAO (render/util.ispc:12-54)
static __device__ __noinline__ float computeAO(Ctx &c, const DifferentialGeometry &dg, int sampleCnt, float aoRadius, uint32_t px, uint32_t py,
    int sampleIDz)
{
  const int accumID = sampleIDz * sampleCnt;
  uint32_t v0 = px, v1 = py;
  tea8(v0, v1);
  const vec2f rot = make_vec2f(to_float_unorm(v0), to_float_unorm(v1));
  int hits = 0;
  const linear3f localToWorld = frame(dg.Ns);
  for (int i = 0; i < sampleCnt; i++) {
    const vec2f halton = make_vec2f(PermHalton_sample5(accumID + i), PermHalton_sample7(accumID + i));
    const vec2f r = make_vec2f(CranleyPattersonRotation(halton.x, rot.x), CranleyPattersonRotation(halton.y, rot.y));
    const vec3f ao_dir = localToWorld * cosineSampleHemisphere(r);
    if (dot(ao_dir, dg.Ns) < 0.01f) { hits++; continue; }
    Ray ao_ray;
    setRay(ao_ray, dg.P, ao_dir, dg.epsilon, aoRadius);
    c.cnt.shadow++;
    if (traceWorldNI<true>(c.world, ao_ray, c.stack)) hits++;
  }
  return 1.0f - (hits / (float)sampleCnt);
}

// AORenderer_renderSample (render/ao/AORenderer.ispc:23-149)
static __device__ void AO_renderSample(Ctx &c, Ray &ray, ScreenSample &sample, uint32_t px, uint32_t py, int sampleIDz)
{
  const OcRenderer &ren = c.ren;
  RandomSampler rs;
  RandomSampler_init(rs, px + (uint32_t)c.fb.sizeX * py, (uint32_t)sampleIDz);
  bool firstHit = true;
  const float originalRayTFar = ray.t;
  const vec3f primaryDir = ray.dir;
  vec4f outputColor = make_vec4f(0.f, 0.f, 0.f, 0.f);
  while (outputColor.w < 0.99f) {
    if (firstHit) c.cnt.primary++; else c.cnt.secondary++;
    traceWorldNI<false>(c.world, ray, c.stack);
    if (c.world->hasVolumes) {
      VolumeHit vh;
      if (traceVolumeRay(c.world, ray, vh)) {
        const vec4f vc = integrateVolumeAO(c, vh, ray, rs, ren.volumeSamplingRate);
        outputColor = outputColor + (1.f - outputColor.w) * vc;
      }
    }
    vec4f blendedColor;
    const bool rayHadHit = hadHit(ray);
    if (rayHadHit) {
      DifferentialGeometry dg;
      postIntersectNI<TEX>(c.world, ren, dg, ray, DG_NG | DG_NS | DG_NORMALIZE | DG_FACEFORWARD | DG_COLOR | DG_TEXCOORD | DG_TANGENTS, TEX ? &c.cone : nullptr);
      // getSurfaceColor (ao/surfaces.ispc:18-42)
      vec3f surfaceColor = make_vec3f(1.f);
      float opacity = 1.f;
      if (dg.material && dg.material->type == OC_MATERIAL_OBJ) {
        const OcMaterial *m = dg.material;
        surfaceColor = toV(m->Kd);
        if (TEX && m->mapKd.tex) {
          const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint);
          surfaceColor = surfaceColor * make_vec3f(t.x, t.y, t.z);
          opacity *= t.w;
        }
        opacity *= m->d * ((TEX && m->mapD.tex) ? texParamGet4f(m->mapD, dg.st, dg.pixelFootprint).x : 1.f);
      }
      const vec3f color = surfaceColor * make_vec3f(dg.color);
      opacity = opacity * dg.color.w;
      float intensity = fabsf(dot(dg.Ns, primaryDir)) * opacity;
      if (ren.aoSamples > 0 && ren.aoIntensity > 0.01f) {
        const float ao = computeAO(c, dg, ren.aoSamples, ren.aoRadius, px, py, sampleIDz);
        intensity *= (1.0f - ren.aoIntensity) + (ao * ren.aoIntensity);
      }
      blendedColor = make_vec4f(color * intensity, opacity);
      if (firstHit) {
        sample.z = ray.t;
        if (c.fb.projectedDepth) sample.z *= dot(ray.dir, toV(c.cam.direction));
        sample.albedo = color; sample.position = dg.P; sample.firstNormal = dg.Ns; sample.normal = dg.Ns;
        sample.instID = dg.instID; sample.geomID = dg.objID; sample.primID = dg.primID;
      }
      setRay(ray, ray.org, ray.dir, ray.t + dg.epsilon, originalRayTFar);
    } else {
      blendedColor = TEX ? getBackgroundTex(ren, c.pixelCenter) : getBackground(ren);
      if (firstHit) {
        sample.z = ray.t;
        sample.albedo = make_vec3f(blendedColor);
        sample.normal = primaryDir;
      }
    }
    outputColor = outputColor + (1.f - outputColor.w) * blendedColor;
    firstHit = false;
    if (!rayHadHit) break;
  }
  sample.rgb = make_vec3f(outputColor);
  sample.alpha = outputColor.w;
}

// ======================================= debug renderer ======================================
// render/debug/DebugRenderer.ispc:27-420 (methods of DebugRendererShared.h:12-30; texCoord, dPdt and volume are refused at
// commit: no texture coordinates / second tangent / debug volume marcher on this device)
static __device__ __forceinline__ vec3f makeRandomColor(uint32_t i) // rkcommon/utility/random.ih (un-vendored; restated)
{
  const uint32_t mx = 13 * 17 * 43, my = 11 * 29, mz = 7 * 23 * 63;
  const uint32_t g = i * (3 * 5 * 127) + 12312314;
  if (i == 0xFFFFFFFFu) return make_vec3f(0.f);
  return make_vec3f((g % mx) * (1.f / (mx - 1)), (g % my) * (1.f / (my - 1)), (g % mz) * (1.f / (mz - 1)));
}
static __device__ __forceinline__ float debugEyeLight(const Ray &ray) { return 0.2f + 0.8f * fabsf(dot(normalize(ray.Ng), ray.dir)); }
static __device__ __forceinline__ vec3f debugTangentColor(vec3f t)
{
  vec3f c = normalize(t);
  if (c.x < 0.f) c.x = c.x * -0.3f;
  if (c.y < 0.f) c.y = c.y * -0.3f;
  if (c.z < 0.f) c.z = c.z * -0.3f;
  return c;
}

static __device__ void Debug_renderSample(Ctx &c, Ray &ray, ScreenSample &sample, uint32_t px, uint32_t py, int sampleIDz)
{
  const OcRenderer &ren = c.ren;
  const uint32_t method = ren.debugMethod;
  if (method == OC_DEBUG_TEST_FRAME) { // no ray at all (DebugRenderer.ispc:50-62)
    sample.rgb = make_vec3f((px % 256) / 255.f, (py % 256) / 255.f, ((px + py + (uint32_t)sampleIDz) % 256) / 255.f);
    sample.alpha = 1.f;
    sample.z = 1.f;
  } else if (method == OC_DEBUG_RAY_DIR) {
    sample.rgb = vabs(ray.dir);
    sample.alpha = 1.f;
    sample.z = 1.f;
  } else {
    // hitBackground (DebugRenderer.ispc:27-38)
    c.cnt.primary++;
    traceWorldNI<false>(c.world, ray, c.stack);
    sample.z = ray.t;
    sample.alpha = 1.f;
    sample.rgb = make_vec3f(TEX ? getBackgroundTex(ren, c.pixelCenter) : getBackground(ren));
    if (hadHit(ray)) {
      DifferentialGeometry dg;
      switch (method) {
      case OC_DEBUG_TEX_COORD: // DebugRenderer.ispc:112-122
        postIntersect<true>(c.world, ren, dg, ray, DG_TEXCOORD, &c.cone);
        sample.rgb = vabs(make_vec3f(dg.st.x, dg.st.y, 0.0f));
        break;
      case OC_DEBUG_DPDT: // :142-160
        postIntersect<true>(c.world, ren, dg, ray, DG_TANGENTS, &c.cone);
        sample.rgb = debugTangentColor(dg.dPdt);
        break;
      case OC_DEBUG_EYE_LIGHT: sample.rgb = make_vec3f(debugEyeLight(ray)); break;
      case OC_DEBUG_NG:
        postIntersect(c.world, ren, dg, ray, DG_NORMALIZE | DG_NG);
        sample.rgb = vabs(dg.Ng);
        break;
      case OC_DEBUG_NS:
        postIntersect(c.world, ren, dg, ray, DG_NORMALIZE | DG_NS);
        sample.rgb = vabs(dg.Ns);
        break;
      case OC_DEBUG_COLOR:
        postIntersect(c.world, ren, dg, ray, DG_COLOR | DG_NS);
        sample.rgb = make_vec3f(dg.color) * fabsf(dot(normalize(ray.dir), normalize(dg.Ns)));
        break;
      case OC_DEBUG_DPDS:
        postIntersect(c.world, ren, dg, ray, DG_TANGENTS);
        sample.rgb = debugTangentColor(dg.dPds);
        break;
      case OC_DEBUG_PRIM_ID: sample.rgb = debugEyeLight(ray) * makeRandomColor(ray.primID); break;
      case OC_DEBUG_GEOM_ID: sample.rgb = debugEyeLight(ray) * makeRandomColor(ray.geomID); break;
      case OC_DEBUG_INST_ID: sample.rgb = debugEyeLight(ray) * makeRandomColor(ray.instID); break;
      case OC_DEBUG_BACKFACING_NG:
        sample.rgb = make_vec3f(debugEyeLight(ray));
        if (dot(ray.Ng, ray.dir) > 0.f) sample.rgb.y = 0.f;
        break;
      case OC_DEBUG_BACKFACING_NS: {
        postIntersect(c.world, ren, dg, ray, DG_NORMALIZE | DG_NS);
        const float cs = dot(dg.Ns, ray.dir);
        sample.rgb = make_vec3f(.2f + .8f * fabsf(cs));
        if (cs > 0.f) sample.rgb.y = 0.f;
        break;
      }
      default: break;
      }
    }
  }
  if (c.fb.projectedDepth) sample.z *= dot(ray.dir, toV(c.cam.direction));
}

// Renderer_default_renderTask pixel body (render/RendererRenderTaskFn.inl:30-86)
template <int RENDERER>
static __device__ void defaultRenderPixel(Ctx &c, uint32_t x, uint32_t y, ScreenSample &screenSample)
{
  const OcRenderer &ren = c.ren;
  const OcFrameBuffer &fb = c.fb;
  const int spp = ren.spp;
  const int startSampleID = fb.frameID * spp + 1;
  ScreenSample_zero(screenSample);
  // the reference declares `sample` outside the spp loop and never re-zeroes it
  // (RendererRenderTaskFn.inl:32-33): AOVs / IDs of an earlier sample persist through a miss
  ScreenSample sample;
  ScreenSample_zero(sample);
  const vec2f rcpSize = make_vec2f(1.f / (float)fb.sizeX, 1.f / (float)fb.sizeY);
  c.pixelCenter = make_vec2f(((float)x + 0.5f) * rcpSize.x, ((float)y + 0.5f) * rcpSize.y);
  c.cone.width = c.cone.dwdt = 0.f;
  const float tMax = TEX ? rendererMaxDepth(ren, c.pixelCenter) : INFINITY; // depth limit from map_maxDepth (:41-43)
  for (int s = 0; s < spp; s++) {
    const float pixel_du = Halton_sample2(startSampleID + s);
    const float pixel_dv = CranleyPattersonRotation(Halton_sample3(startSampleID + s), 1.f / 6.f);
    const vec2f pfSample = pixelFilterSample(ren, make_vec2f(pixel_du, pixel_dv));
    const vec2f pixel_center = make_vec2f(((float)x + 0.5f) * rcpSize.x, ((float)y + 0.5f) * rcpSize.y);
    const vec2f screen = pixel_center + pfSample * rcpSize;
    Ray ray;
    cameraInitRay(c.cam, ray, screen, make_vec2f(0.f, 0.f), pixel_center);
    if (TEX || RENDERER == OC_RENDERER_DEBUG) { // the sample's ray cone in pixel units (:64-71)
      c.cone = cameraInitRayCone(c.cam, screen, pixel_center);
      c.cone.dwdt *= rcpSize.y; c.cone.width *= rcpSize.y;
    }
    if (TEX) ray.t = fminf(ray.t, tMax);
    if (RENDERER == OC_RENDERER_SCIVIS) SciVis_renderSample(c, ray, sample, x, y, startSampleID + s);
    else if (RENDERER == OC_RENDERER_DEBUG) Debug_renderSample(c, ray, sample, x, y, startSampleID + s);
    else AO_renderSample(c, ray, sample, x, y, startSampleID + s);
    ScreenSample_accumulate(screenSample, sample);
  }
  ScreenSample_normalize(screenSample, 1.f / (float)spp);
}
}; // struct Mega

// ================================ framebuffer accumulation ==================================
// LocalFB_accumulateSample (fb/LocalFB.ispc:14-78), FRAMEBUFFER_ACCUMULATE_VALUE (FrameBuffer.ih:17-22)
__device__ __forceinline__ float accum1(float value, float nv, float f) { return (f > 0.8f) ? nv : lerpf(f, value, nv); }
__device__ __forceinline__ void accum3(OcVec3f *buf, size_t i, vec3f nv, float f)
{
  if (f > 0.8f) { buf[i].x = nv.x; buf[i].y = nv.y; buf[i].z = nv.z; return; }
  const OcVec3f o = buf[i];
  buf[i].x = lerpf(f, o.x, nv.x); buf[i].y = lerpf(f, o.y, nv.y); buf[i].z = lerpf(f, o.z, nv.z);
}
__device__ __forceinline__ void accum4(OcVec4f *buf, size_t i, vec4f nv, float f)
{
  float4 *p = reinterpret_cast<float4 *>(buf) + i;
  float4 o;
  if (f > 0.8f) o = make_float4(nv.x, nv.y, nv.z, nv.w);
  else {
    o = *p;
    o.x = lerpf(f, o.x, nv.x); o.y = lerpf(f, o.y, nv.y); o.z = lerpf(f, o.z, nv.z); o.w = lerpf(f, o.w, nv.w);
  }
  *p = o;
}

__device__ __forceinline__ void accumulateSample(const OcFrameBuffer &fb, uint32_t x, uint32_t y, const ScreenSample &s)
{
  const size_t pixelID = x + (size_t)fb.sizeX * y;
  const int accumID = fb.frameID;
  const float accScale = 1.f / (float)(accumID + 1);
  const vec4f colorSample = make_vec4f(s.rgb, s.alpha);
  if (fb.color) accum4(fb.color, pixelID, colorSample, accScale);
  if (fb.accumulateVariance) accum4(fb.variance, pixelID, colorSample, 1.f / (float)(accumID / 2 + 1));
  bool updateIDBufs = accumID == 0;
  if (fb.depth) {
    if (accumID > 0 && fabsf(s.z) < fabsf(fb.depth[pixelID])) updateIDBufs = true;
    if (updateIDBufs) fb.depth[pixelID] = s.z;
  }
  if (updateIDBufs) {
    if (fb.primID) fb.primID[pixelID] = s.primID;
    if (fb.objID) fb.objID[pixelID] = s.geomID;
    if (fb.instID) fb.instID[pixelID] = s.instID;
  }
  if (fb.normal) accum3(fb.normal, pixelID, s.normal, accScale);
  if (fb.firstNormal) accum3(fb.firstNormal, pixelID, s.firstNormal, accScale);
  if (fb.albedo) accum3(fb.albedo, pixelID, s.albedo, accScale);
  if (fb.position) accum3(fb.position, pixelID, s.position, accScale);
}

// frame control words (OcFrameBuffer.control): [0] cancel flag, [1] work units done, [2] work units of the frame
__device__ __forceinline__ bool frameCancelled(const OcFrameBuffer &fb)
{
  return fb.control != nullptr && *reinterpret_cast<const volatile uint32_t *>(fb.control) != 0u;
}

// ======================================= kernels ===========================================
struct FrameArgs
{
  const OcWorld *world;
  OcCamera cam;
  OcRenderer ren;
  OcFrameBuffer fb;
  uint32_t tilesX, tilesY;
  uint32_t ptFirstBounceLightSamples, ptIndirectLightSamples; // wavefront pathtracer (PathTracer.ispc:96-104)
};

} // namespace oc
#include "pathtracer_wf.cuh"
namespace oc {

#ifndef OC_RENDER_MIN_BLOCKS
#define OC_RENDER_MIN_BLOCKS 10 // 96 registers: C2 7.9 -> 7.5 ms, C3 143 -> 137 ms vs the unconstrained 128 (tools/ab_workloads.sh)
#endif
template <int RENDERER, bool TEX>
__global__ void __launch_bounds__(64, OC_RENDER_MIN_BLOCKS) renderFrameKernel(const __grid_constant__ FrameArgs a)
{
  // block -> (owned tile, 8x8 task inside the 64x64 tile); thread -> pixel inside the task
  const uint32_t ownedTile = blockIdx.x >> 6, taskInTile = blockIdx.x & 63u;
  const uint32_t tileID = a.fb.rank + ownedTile * a.fb.numRanks;
  const uint32_t tx = tileID % a.tilesX, ty = tileID / a.tilesX;
  const uint32_t x = tx * 64 + (taskInTile & 7u) * 8 + (threadIdx.x & 7u);
  const uint32_t y = ty * 64 + (taskInTile >> 3) * 8 + (threadIdx.x >> 3);
  __shared__ uint32_t s_stack[OC_STACK_SHARED * 64];
  Stack stack;
  stack.base = s_stack + threadIdx.x;
  stack.stride = 64;
  stack.ovf = nullptr; // traceHitNI keeps the overflow entries in its own frame
  Ctx c(a.world, a.ren, a.cam, a.fb, stack);
  if (frameCancelled(a.fb)) return; // ospCancel: render tasks that have not started yet do not start
  if (a.fb.control && blockIdx.x == 0 && threadIdx.x == 0) a.fb.control[2] = gridDim.x;
  // adaptive accumulation: a converged render task is skipped as a whole (one CTA = one task; LocalFB.cpp:245-257)
  const bool inImage = x < (uint32_t)a.fb.sizeX && y < (uint32_t)a.fb.sizeY;
  if (inImage && (!a.fb.taskError || a.fb.taskError[(y >> 3) * (((uint32_t)a.fb.sizeX + 7u) >> 3) + (x >> 3)] > a.fb.errorThreshold)) {
    ScreenSample ss;
    Mega<TEX>::template defaultRenderPixel<RENDERER>(c, x, y, ss);
    accumulateSample(a.fb, x, y, ss);
  }
  if (a.fb.rayCounters) {
    uint32_t p = c.cnt.primary, s = c.cnt.secondary, h = c.cnt.shadow;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      p += __shfl_xor_sync(0xffffffffu, p, o);
      s += __shfl_xor_sync(0xffffffffu, s, o);
      h += __shfl_xor_sync(0xffffffffu, h, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(a.fb.rayCounters + 0, (unsigned long long)p);
      atomicAdd(a.fb.rayCounters + 1, (unsigned long long)s);
      atomicAdd(a.fb.rayCounters + 2, (unsigned long long)h);
    }
  }
  __syncthreads();
  if (a.fb.control && threadIdx.x == 0) atomicAdd(a.fb.control + 1, 1u); // one render-task CTA finished (ospGetProgress)
}

// batch traversal for tests / ospPick-like queries
__global__ void traceKernel(const OcWorld *world, OcRayHit *rays, uint32_t n, int occlusion)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ uint32_t s_stack[OC_STACK_SHARED * OC_MAX_BLOCK];
  if (i >= n) return;
  uint32_t ovf[OC_STACK_OVF];
  Stack stack;
  stack.base = s_stack + threadIdx.x;
  stack.stride = blockDim.x;
  stack.ovf = ovf;
  OcRayHit rh = rays[i];
  Ray r;
  setRay(r, make_vec3f(rh.org[0], rh.org[1], rh.org[2]), make_vec3f(rh.dir[0], rh.dir[1], rh.dir[2]), rh.t0, rh.t);
  r.u = r.v = 0.f; r.Ng = make_vec3f(0.f);
  if (occlusion) {
    const Hit h = traceHit<1>(world, r.org, r.dir, r.t0, r.t, stack);
    if (h.inst != OC_INVALID_ID) { r.t = -INFINITY; r.instID = h.inst; resolveHitIds(world, h.rec, h.inst, r.primID, r.geomID); }
  } else {
    const Hit h = traceHit<0>(world, r.org, r.dir, r.t0, r.t, stack);
    resolveHit(world, h, r);
  }
  rh.t = r.t; rh.u = r.u; rh.v = r.v; rh.Ng[0] = r.Ng.x; rh.Ng[1] = r.Ng.y; rh.Ng[2] = r.Ng.z;
  rh.primID = r.primID; rh.geomID = r.geomID; rh.instID = r.instID;
  rays[i] = rh;
}

} // namespace oc

using namespace oc;

// ---- wavefront path tracer: scratch + launch sequence ----------------------------------------------------------
namespace {
// optional per-launch-class timing (ospcuda_profile_enable): CUDA event pairs around every wavefront launch
struct PtProfile
{
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4]; // 0 TRACE, 1 SHADE, 2 GEN, 3 TAIL
  std::vector<cudaEvent_t> pool;
  size_t poolNext = 0;
  cudaEvent_t get()
  {
    if (poolNext == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[poolNext++];
  }
} g_prof;

struct ProfScope
{
  int cls; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t st;
  ProfScope(int c, cudaStream_t s) : cls(c), st(s) { if (g_prof.enabled) { a = g_prof.get(); b = g_prof.get(); cudaEventRecord(a, st); } }
  ~ProfScope() { if (a) { cudaEventRecord(b, st); g_prof.ev[cls].emplace_back(a, b); } }
};

struct PtHost
{
  void *block = nullptr;
  size_t blockBytes = 0;
  int device = -1;
  int traceBlocks = 0, shadeBlocks[3] = {0, 0, 0}, genBlocks = 0, tailBlocks = 0, aoShadeBlocks = 0, traceVariant = 1, shadeVariant = 4; // persistent grids (resident CTAs per SM x SM count)
};
PtHost g_ptPerDevice[32]; // one scratch block per CUDA device of the process (in-process multidevice)
PtHost &ptHost()
{
  int dev = 0;
  cudaGetDevice(&dev);
  return g_ptPerDevice[dev & 31];
}
#define g_pt (ptHost())

size_t alignUp(size_t v) { return (v + 255) & ~size_t(255); }

// carve the SoA arrays for `paths` slots out of one allocation (grown on demand, kept across frames)
int ptCarve(oc::PtScratch &s, uint32_t paths, uint32_t nls, bool needAcc, uint32_t numCounters, bool hasVol)
{
  const size_t P = paths;
  size_t bytes = 0;
  auto take = [&](size_t n) { const size_t o = bytes; bytes += alignUp(n); return o; };
  size_t off[40];
  int k = 0;
  for (int i = 0; i < 8; i++) off[k++] = take(P * 16);       // rayO rayD hit0 hit1 thr con lastP lastN
  off[k++] = take(P * 4);                                    // hitInst
  off[k++] = take(P * 8);                                    // smp
  off[k++] = take(P * 4);                                    // flags
  for (int i = 0; i < 5; i++) off[k++] = take(P * 16);       // sZ sPos sFN sNrm sAlb
  for (int i = 0; i < 3; i++) off[k++] = take(P * 16 * nls); // shA shB shC
  off[k++] = take(P * 8 * nls);                              // shHit
  for (int i = 0; i < 6; i++) off[k++] = take(needAcc ? P * 16 : 0);
  off[k++] = take(hasVol ? P * 16 : 0);                      // rnd
  off[k++] = take(hasVol ? P * 4 * nls : 0);                 // shT0
  off[k++] = take(P * 8);                                    // cone
  for (int i = 0; i < 2; i++) off[k++] = take(P * 16);       // med shMed
  for (int i = 0; i < 2; i++) off[k++] = take(P * 4);        // pathQ
  for (int i = 0; i < 2; i++) off[k++] = take(P * 4 * nls);  // shadowQ
  off[k++] = take(numCounters * 4);
  int dev = 0;
  cudaGetDevice(&dev);
  PtHost &h = g_pt;
  if (bytes > h.blockBytes) {
    if (h.block) { OC_TRY(cudaDeviceSynchronize()); cudaFree(h.block); }
    h.block = nullptr; h.blockBytes = 0;
    OC_TRY(cudaMalloc(&h.block, bytes));
    h.blockBytes = bytes; h.device = dev;
  }
  char *b = (char *)h.block;
  k = 0;
  s.rayO = (float4 *)(b + off[k++]); s.rayD = (float4 *)(b + off[k++]); s.hit0 = (float4 *)(b + off[k++]); s.hit1 = (float4 *)(b + off[k++]);
  s.thr = (float4 *)(b + off[k++]); s.con = (float4 *)(b + off[k++]); s.lastP = (float4 *)(b + off[k++]); s.lastN = (float4 *)(b + off[k++]);
  s.hitInst = (uint32_t *)(b + off[k++]); s.smp = (uint2 *)(b + off[k++]); s.flags = (uint32_t *)(b + off[k++]);
  s.sZ = (float4 *)(b + off[k++]); s.sPos = (float4 *)(b + off[k++]); s.sFN = (float4 *)(b + off[k++]); s.sNrm = (float4 *)(b + off[k++]);
  s.sAlb = (float4 *)(b + off[k++]);
  s.shA = (float4 *)(b + off[k++]); s.shB = (float4 *)(b + off[k++]); s.shC = (float4 *)(b + off[k++]);
  s.shHit = (uint2 *)(b + off[k++]);
  s.accC = (float4 *)(b + off[k++]); s.accZ = (float4 *)(b + off[k++]); s.accPos = (float4 *)(b + off[k++]); s.accFN = (float4 *)(b + off[k++]);
  s.accNrm = (float4 *)(b + off[k++]); s.accAlb = (float4 *)(b + off[k++]);
  s.rnd = (uint4 *)(b + off[k++]); s.shT0 = (float *)(b + off[k++]);
  s.cone = (float2 *)(b + off[k++]);
  s.med = (float4 *)(b + off[k++]); s.shMed = (float4 *)(b + off[k++]);
  s.pathQ[0] = (uint32_t *)(b + off[k++]); s.pathQ[1] = (uint32_t *)(b + off[k++]);
  s.shadowQ[0] = (uint32_t *)(b + off[k++]); s.shadowQ[1] = (uint32_t *)(b + off[k++]);
  s.counters = (uint32_t *)(b + off[k++]);
  s.nls = nls;
  return 0;
}

size_t ptBytesPerPath(uint32_t nls, bool needAcc, bool hasVol)
{
  return 8 * 16 + 16 + 5 * 16 + (48 + 8 + 8) * (size_t)nls + (needAcc ? 96 : 0) + 8 + 8 + 32 + (hasVol ? 16 + 4 * (size_t)nls : 0);
}

// Launch sequence of one sample pass: GEN, then TRACE(d) / SHADE(d) for d = 0 .. maxDepth. Every stage d owns four
// counters (path / continuation / shadow queue sizes, TRACE's fetch cursor), zeroed once per pass; SHADE(d) appends to the
// queues of stage d+1 (ping-pong on the parity of d), TRACE(d+1) and SHADE(d+1) consume them. Launches whose queues turn
// out empty exit at once.
int renderPathTracerWavefront(const oc::FrameArgs &fa, uint32_t ownedTiles, bool ao = false)
{
  using namespace oc;
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  PtArgs k;
  k.a = fa;
  // light sample counts (PathTracer.ispc:96-104); world->numLights lives in device memory: mirror it via the args
  k.nFirst = fa.ptFirstBounceLightSamples;
  k.nIndirect = fa.ptIndirectLightSamples;
  // ao: one vertex per surface the sample passes through; the loop runs until every sample is opaque or has left the
  // scene (the queue size is read back after each bounce from the second on), 250 bounds the counter space
  const uint32_t maxDepth = ao ? 250u : (fa.ren.maxDepth < 250u ? fa.ren.maxDepth : 250u);
  if (!ao) k.a.ren.maxDepth = maxDepth;
  const uint32_t nls = ao ? (uint32_t)std::max(1, fa.ren.aoSamples) : (k.nFirst > 0 ? k.nFirst : 1);
  k.shadowT0InB = ao ? 1u : 0u;
  const bool needAcc = fa.ren.spp > 1;
  const uint32_t numCounters = 4 * (maxDepth + 2);
  // batch the owned tiles so that the path state stays under ~1.5 GB (many light samples / 4K frames)
  const size_t budget = size_t(1536) << 20;
  const bool hasVol = !ao && fa.ren.worldHasVolumes != 0;
  k.vol = hasVol ? 1u : 0u;
  const bool hasTex = !ao && fa.ren.hasTextures != 0; // the kernels for worlds with volumes are compiled with texture support too
  uint32_t tilesPerBatch = (uint32_t)std::max<size_t>(1, budget / (ptBytesPerPath(nls, needAcc, hasVol) * 4096));
  // (smaller batches were tried for L2 residency of the path state, alternating between two streams so that one batch's
  // thinned-out end overlaps the next batch's dense bounces: SHADE -17 %, but every batch pays its own tail and the
  // persistent grids leave no room for real overlap -- C1 2.54 -> 2.76 ms, C5s 35 -> 54 ms, C3 93 -> 120 ms; profiles/)
  static const int batchEnv = getenv("OSPCUDA_PT_BATCH_TILES") ? atoi(getenv("OSPCUDA_PT_BATCH_TILES")) : 0;
  if (batchEnv > 0 && tilesPerBatch > (uint32_t)batchEnv) tilesPerBatch = (uint32_t)batchEnv;
  if (tilesPerBatch > ownedTiles) tilesPerBatch = ownedTiles;
  if ((uint64_t)tilesPerBatch * 4096u * nls >= (1ull << 31))
    return oc_set_error(cudaErrorInvalidValue, "pathtracer: lightSamples x paths per batch exceeds the shadow-ray id space");
  if (int rc = ptCarve(k.s, tilesPerBatch * 4096u, nls, needAcc, numCounters, hasVol)) return rc;
  if (!g_pt.traceBlocks) {
    int dev = 0, sms = 148, a = 1, b = 1, bv = 1, g = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // A/B knobs: OSPCUDA_PT_REFILL = idle lanes that trigger a refill (4 / 8 / 16), OSPCUDA_PT_TRACE_BLOCKS = resident CTAs per SM
    // the register budget is set for (5: no spills, 6: 80 registers)
    const int refill = getenv("OSPCUDA_PT_REFILL") ? atoi(getenv("OSPCUDA_PT_REFILL")) : 16;
    const int minb = getenv("OSPCUDA_PT_TRACE_BLOCKS") ? atoi(getenv("OSPCUDA_PT_TRACE_BLOCKS")) : 5;
    if (minb >= 6) g_pt.traceVariant = refill <= 16 ? 4 : 5;
    else g_pt.traceVariant = refill <= 8 ? 0 : refill <= 16 ? 1 : refill <= 24 ? 2 : 3;
    if (minb >= 6) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, ptTraceKernel<16, 6>, PT_TRACE_BLOCK, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, ptTraceKernel<16, 5>, PT_TRACE_BLOCK, 0);
    const int smb = getenv("OSPCUDA_PT_SHADE_BLOCKS") ? atoi(getenv("OSPCUDA_PT_SHADE_BLOCKS")) : 4;
    g_pt.shadeVariant = smb <= 4 ? 4 : smb == 5 ? 5 : 6;
    int bt = 1;
    if (g_pt.shadeVariant == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, ptShadeKernel<false, 4, false>, PT_TRACE_BLOCK, 0);
    else if (g_pt.shadeVariant == 5) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, ptShadeKernel<false, 5, false>, PT_TRACE_BLOCK, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, ptShadeKernel<false, 6, false>, PT_TRACE_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bv, ptShadeKernel<true, 4, true>, PT_TRACE_BLOCK, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt, ptShadeKernel<false, 4, true>, PT_TRACE_BLOCK, 0);
    g_pt.shadeBlocks[2] = sms * (bt > 0 ? bt : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g, ptGenKernel<false, false>, 128, 0);
    g_pt.traceBlocks = sms * (a > 0 ? a : 1);
    g_pt.shadeBlocks[0] = sms * (b > 0 ? b : 1);
    g_pt.shadeBlocks[1] = sms * (bv > 0 ? bv : 1);
    g_pt.genBlocks = sms * (g > 0 ? g : 1);
    int t = 1, as = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&as, aoShadeKernel, 128, 0);
    g_pt.aoShadeBlocks = sms * (as > 0 ? as : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t, ptTailKernel<false, false>, PT_TRACE_BLOCK, 0);
    g_pt.tailBlocks = sms * (t > 0 ? t : 1);
  }
  // depth from which the tail kernel takes over. A TRACE + SHADE launch pair pays for itself while the queue holds many
  // paths per resident lane; below that a launch costs its latency floor whatever it holds, and carrying the paths to their
  // end lane by lane is cheaper. With the whole 1024^2 frame on one GPU that point is two bounces after Russian roulette
  // starts to cull; a sub-device's share of a multi-GPU frame (or a small image) gets there at the roulette depth itself
  // (tools/gpu_r2s.sh, one sub-device's tiles of C1: 1/2 1.48 -> 1.44 ms, 1/4 0.92 -> 0.86 ms, 1/8 0.63 -> 0.55 ms).
  // OSPCUDA_PT_TAIL_FROM overrides, 0 = never.
  static const int tailEnv = getenv("OSPCUDA_PT_TAIL_FROM") ? atoi(getenv("OSPCUDA_PT_TAIL_FROM")) : -1;
  const uint32_t pathsPerBatch = tilesPerBatch * 4096u;
  const uint32_t tailAuto = (uint32_t)std::max(1, fa.ren.rouletteDepth + (pathsPerBatch >= 768u * 1024u ? 2 : 0));
  const uint32_t tailFrom = tailEnv == 0 ? 0xffffffffu : tailEnv > 0 ? (uint32_t)tailEnv : tailAuto;
  static const bool foldEnv = !(getenv("OSPCUDA_PT_FOLD_IN_TRACE") && atoi(getenv("OSPCUDA_PT_FOLD_IN_TRACE")) == 0);
  // (worlds whose occluders may be transparent keep the fold in SHADE: the occluder's material chain and the closest-hit
  // loop are better off there -- C4 5.0 vs 5.6 ms)
  static const bool sortShade = !(getenv("OSPCUDA_PT_SORT_SHADE") && atoi(getenv("OSPCUDA_PT_SORT_SHADE")) == 0);
  const bool foldInTrace = !ao && !hasVol && nls == 1 && foldEnv && fa.ren.worldShadowOpaque != 0;
  const uint32_t progressStages = ao ? 2u : std::min(maxDepth + 1u, tailFrom == 0xffffffffu ? maxDepth + 1u : tailFrom + 1u);
  k.progressTotal = ((ownedTiles + tilesPerBatch - 1) / tilesPerBatch) * (uint32_t)fa.ren.spp * progressStages;
  k.progressUnit = 0;
  auto C = [&](uint32_t stage, int which) { return k.s.counters + 4 * stage + which; }; // 0 path, 1 cont, 2 shadow, 3 cursor
  for (uint32_t tile0 = 0; tile0 < ownedTiles; tile0 += tilesPerBatch) {
    const uint32_t nt = std::min(tilesPerBatch, ownedTiles - tile0);
    k.firstTile = tile0;
    k.numPaths = nt * 4096u;
    for (int pass = 0; pass < fa.ren.spp; pass++) {
      k.pass = pass;
      OC_TRY(cudaMemsetAsync(k.s.counters, 0, numCounters * sizeof(uint32_t), st));
      k.depth = 0;
      k.pathIn = k.shadowIn = nullptr; k.cntPathIn = k.cntShadowIn = nullptr;
      k.shadowOut = nullptr; k.cntShadowOut = nullptr; k.fetchCursor = nullptr;
      k.pathOut = k.s.pathQ[0]; k.cntPathOut = C(0, 0);
      {
        ProfScope ps(2, st);
        const int gb = (int)std::min<uint32_t>((uint32_t)g_pt.genBlocks, (k.numPaths + 127u) / 128u);
        if (ao) aoGenKernel<<<gb, 128, 0, st>>>(k);
        else if (hasVol) ptGenKernel<true, true><<<gb, 128, 0, st>>>(k);
        else if (hasTex) ptGenKernel<false, true><<<gb, 128, 0, st>>>(k);
        else ptGenKernel<false, false><<<gb, 128, 0, st>>>(k);
      }
      oc_count_launch();
      for (uint32_t d = 0; d <= maxDepth; d++) {
        k.depth = d;
        // ospGetProgress: stages (bounces) finished out of those a frame can take; the tail / an early end skips the rest
        k.progressUnit = ((tile0 / tilesPerBatch) * (uint32_t)fa.ren.spp + (uint32_t)pass) * progressStages + std::min(d, progressStages - 1u) + 1u;
        const uint32_t q = d & 1u;
        if (!ao && d > 0 && d >= tailFrom) { // one launch carries every remaining path to its end
          k.pathIn = k.s.pathQ[q]; k.cntPathIn = C(d, 0);
          k.fetchCursor = C(d, 3);
          {
            ProfScope ps(3, st);
            if (hasVol) ptTailKernel<true, true><<<g_pt.tailBlocks, PT_TRACE_BLOCK, 0, st>>>(k);
            else if (hasTex) ptTailKernel<false, true><<<g_pt.tailBlocks, PT_TRACE_BLOCK, 0, st>>>(k);
            else ptTailKernel<false, false><<<g_pt.tailBlocks, PT_TRACE_BLOCK, 0, st>>>(k);
          }
          oc_count_launch();
          break;
        }
        // TRACE(d): the shadow rays emitted at vertex d-1, then the continuation rays of the paths in the path queue
        k.pathIn = k.s.pathQ[q]; k.cntPathIn = C(d, 0);
        k.shadowIn = d == 0 ? nullptr : k.s.shadowQ[q]; k.cntShadowIn = d == 0 ? nullptr : C(d, 2);
        k.fetchCursor = C(d, 3);
        k.foldInTrace = foldInTrace ? 1u : 0u;
        k.sortShade = sortShade ? 1u : 0u;
        {
          ProfScope ps(0, st);
          switch (g_pt.traceVariant) {
          case 0: ptTraceKernel<8, 5><<<g_pt.traceBlocks, PT_TRACE_BLOCK, 0, st>>>(k); break;
          case 1: ptTraceKernel<16, 5><<<g_pt.traceBlocks, PT_TRACE_BLOCK, 0, st>>>(k); break;
          case 2: ptTraceKernel<24, 5><<<g_pt.traceBlocks, PT_TRACE_BLOCK, 0, st>>>(k); break;
          case 3: ptTraceKernel<32, 5><<<g_pt.traceBlocks, PT_TRACE_BLOCK, 0, st>>>(k); break;
          case 4: ptTraceKernel<16, 6><<<g_pt.traceBlocks, PT_TRACE_BLOCK, 0, st>>>(k); break;
          default: ptTraceKernel<24, 6><<<g_pt.traceBlocks, PT_TRACE_BLOCK, 0, st>>>(k); break;
          }
        }
        oc_count_launch();
        // SHADE(d): same path queue in, queues of stage d+1 out
        k.pathOut = k.s.pathQ[q ^ 1u]; k.cntPathOut = C(d + 1, 0);
        k.shadowOut = k.s.shadowQ[q ^ 1u]; k.cntShadowOut = C(d + 1, 2);
        {
          ProfScope ps(1, st);
          if (ao) aoShadeKernel<<<g_pt.aoShadeBlocks, 128, 0, st>>>(k);
          else if (hasVol) ptShadeKernel<true, 4, true><<<g_pt.shadeBlocks[1], PT_TRACE_BLOCK, 0, st>>>(k);
          else if (hasTex) ptShadeKernel<false, 4, true><<<g_pt.shadeBlocks[2], PT_TRACE_BLOCK, 0, st>>>(k);
          else if (g_pt.shadeVariant == 4) ptShadeKernel<false, 4, false><<<g_pt.shadeBlocks[0], PT_TRACE_BLOCK, 0, st>>>(k);
          else if (g_pt.shadeVariant == 5) ptShadeKernel<false, 5, false><<<g_pt.shadeBlocks[0], PT_TRACE_BLOCK, 0, st>>>(k);
          else ptShadeKernel<false, 6, false><<<g_pt.shadeBlocks[0], PT_TRACE_BLOCK, 0, st>>>(k);
        }
        oc_count_launch();
        if (ao && d >= 1) { // anything left behind transparent surfaces?
          uint32_t left = 0;
          OC_TRY(cudaMemcpyAsync(&left, C(d + 1, 0), sizeof(left), cudaMemcpyDeviceToHost, st));
          OC_TRY(cudaStreamSynchronize(st));
          if (left == 0) break;
        }
      }
    }
  }
  return oc_check(cudaGetLastError(), ao ? "wavefront ao renderer" : "wavefront pathtracer");
}
} // namespace

extern "C" int ospcuda_profile_enable(int on)
{
  g_prof.enabled = on != 0;
  for (auto &v : g_prof.ev) v.clear();
  g_prof.poolNext = 0;
  return 0;
}

extern "C" int ospcuda_profile_read(float ms[4], uint32_t launches[4])
{
  OC_TRY(cudaStreamSynchronize((cudaStream_t)ospcuda_stream()));
  for (int c = 0; c < 4; c++) {
    ms[c] = 0.f;
    launches[c] = (uint32_t)g_prof.ev[c].size();
    for (auto &e : g_prof.ev[c]) {
      float t = 0.f;
      cudaEventElapsedTime(&t, e.first, e.second);
      ms[c] += t;
    }
    g_prof.ev[c].clear();
  }
  g_prof.poolNext = 0;
  return 0;
}

extern "C" int ospcuda_render_frame(const OcWorld *d_world, const OcCamera *camera, const OcRenderer *renderer,
    const OcFrameBuffer *fb)
{
  FrameArgs a;
  a.world = d_world; a.cam = *camera; a.ren = *renderer; a.fb = *fb;
  if (a.fb.numRanks == 0) { a.fb.numRanks = 1; a.fb.rank = 0; }
  a.tilesX = (fb->sizeX + 63) / 64; a.tilesY = (fb->sizeY + 63) / 64;
  const uint32_t totalTiles = a.tilesX * a.tilesY;
  const uint32_t owned = a.fb.rank < totalTiles ? (totalTiles - a.fb.rank + a.fb.numRanks - 1) / a.fb.numRanks : 0;
  if (owned == 0) return 0;
  const dim3 grid(owned * 64), block(64);
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  switch (renderer->type) {
  case OC_RENDERER_PATHTRACER: {
    const uint32_t numLights = renderer->worldNumLights;
    a.ptFirstBounceLightSamples = (renderer->numLightSamples >= 0 && numLights > 0) ? (uint32_t)renderer->numLightSamples
                                                                                     : std::min(1024u, numLights);
    a.ptIndirectLightSamples = (renderer->limitIndirectLightSamples && a.ptFirstBounceLightSamples >= 1) ? 1u : a.ptFirstBounceLightSamples;
    return renderPathTracerWavefront(a, owned);
  }
  case OC_RENDERER_SCIVIS:
    if (renderer->hasTextures) renderFrameKernel<OC_RENDERER_SCIVIS, true><<<grid, block, 0, st>>>(a);
    else renderFrameKernel<OC_RENDERER_SCIVIS, false><<<grid, block, 0, st>>>(a);
    break;
  case OC_RENDERER_AO: {
    // worlds without volumes: the ao renderer runs as a wavefront over the ray-queue TRACE kernel (pathtracer_wf.cuh);
    // with volumes the megakernel marches them (OSPCUDA_AO_MEGAKERNEL=1 forces it, for A/B)
    static const bool forceMega = getenv("OSPCUDA_AO_MEGAKERNEL") && atoi(getenv("OSPCUDA_AO_MEGAKERNEL")) > 0;
    // (frames that reference textures take the frame kernel as well: it carries the ray cones)
    if (!renderer->worldHasVolumes && !forceMega && !renderer->hasTextures) return renderPathTracerWavefront(a, owned, true);
    if (renderer->hasTextures) renderFrameKernel<OC_RENDERER_AO, true><<<grid, block, 0, st>>>(a);
    else renderFrameKernel<OC_RENDERER_AO, false><<<grid, block, 0, st>>>(a);
    break;
  }
  case OC_RENDERER_DEBUG:
    if (renderer->hasTextures) renderFrameKernel<OC_RENDERER_DEBUG, true><<<grid, block, 0, st>>>(a);
    else renderFrameKernel<OC_RENDERER_DEBUG, false><<<grid, block, 0, st>>>(a);
    break;
  default: return oc_set_error(cudaErrorInvalidValue, "ospcuda_render_frame: unknown renderer type");
  }
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_render_frame");
}

namespace oc {
// precomputeMicrofacetAlbedoTables (MicrofacetAlbedoTables.ispc:34-160): one thread per directional table entry integrates
// its 32 x 32 stratified samples (MicrofacetAlbedo.ih:131-190); the averages follow by the trapezoidal rule over cosThetaO
__global__ void albedoTablesDirKernel(float *t)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 32; // sqrt(1024) samples per dimension
  int which, idx;
  if (i < OC_MFA_SIZE * OC_MFA_SIZE) { which = 0; idx = (int)i; }
  else if (i < OC_MFA_SIZE * OC_MFA_SIZE + 3u * OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE) {
    const uint32_t j = i - OC_MFA_SIZE * OC_MFA_SIZE;
    which = 1 + (int)(j / (OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE)); idx = (int)(j % (OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE));
  } else if (i < OC_MFA_SIZE * OC_MFA_SIZE + 3u * OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE + OC_MFS_SIZE * OC_MFS_SIZE) {
    which = 4; idx = (int)(i - (OC_MFA_SIZE * OC_MFA_SIZE + 3u * OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE)); // sheen (:140-160)
  } else return;
  float cosThetaO, roughness, eta = 1.f;
  float *dst;
  if (which == 0) {
    cosThetaO = (float)(idx % OC_MFA_SIZE) / (OC_MFA_SIZE - 1); roughness = (float)(idx / OC_MFA_SIZE) / (OC_MFA_SIZE - 1);
    dst = t + OC_MFT_ALBEDO_DIR + idx;
  } else if (which == 4) {
    cosThetaO = (float)(idx % OC_MFS_SIZE) / (OC_MFS_SIZE - 1); roughness = (float)(idx / OC_MFS_SIZE) / (OC_MFS_SIZE - 1);
    dst = t + OC_MFT_SHEEN_DIR + idx;
  } else {
    const int ic = idx % OC_MFD_SIZE, ie = (idx / OC_MFD_SIZE) % OC_MFD_SIZE, ir = idx / (OC_MFD_SIZE * OC_MFD_SIZE);
    cosThetaO = (float)ic / (OC_MFD_SIZE - 1); roughness = (float)ir / (OC_MFD_SIZE - 1);
    const float f = (float)ie / (OC_MFD_SIZE - 1);
    // dielectric / reflection tables: eta in [1/3, 1]; rcp-eta table: eta in [1, 3]
    const float minEta = which == 2 ? 1.f : 1.f / 3.f, maxEta = which == 2 ? 3.f : 1.f;
    eta = (1.f - f) * minEta + f * maxEta;
    dst = t + (which == 1 ? OC_MFT_DIEL_DIR : which == 2 ? OC_MFT_DIELRCP_DIR : OC_MFT_DIELREFL_DIR) + idx;
  }
  const GGX g = makeGGX(roughness, 0.f);
  const float sheenAlpha = fmaxf(sqr(roughness), 0.001f); // roughnessToAlpha
  float sum = 0.f;
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      const vec2f s = make_vec2f(fminf(((float)a + 0.5f) / n, 1.f - 1e-6f), fminf(((float)b + 0.5f) / n, 1.f - 1e-6f));
      sum += which == 0 ? mfAlbedoSample(cosThetaO, g, s) : which == 4 ? mfSheenAlbedoSample(cosThetaO, sheenAlpha, s)
                                                          : mfDielAlbedoSample(cosThetaO, eta, g, s, which == 3);
    }
  *dst = fminf(sum / (n * n), 1.f);
}
__global__ void albedoTablesAvgKernel(float *t)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const float *dir; float *avg; int size;
  if (i < OC_MFA_SIZE) { dir = t + OC_MFT_ALBEDO_DIR + i * OC_MFA_SIZE; avg = t + OC_MFT_ALBEDO_AVG + i; size = OC_MFA_SIZE; }
  else if (i < OC_MFA_SIZE + 3u * OC_MFD_SIZE * OC_MFD_SIZE) {
    const uint32_t j = i - OC_MFA_SIZE, w = j / (OC_MFD_SIZE * OC_MFD_SIZE), r = j % (OC_MFD_SIZE * OC_MFD_SIZE);
    dir = t + (w == 0 ? OC_MFT_DIEL_DIR : w == 1 ? OC_MFT_DIELRCP_DIR : OC_MFT_DIELREFL_DIR) + r * OC_MFD_SIZE;
    avg = t + (w == 0 ? OC_MFT_DIEL_AVG : w == 1 ? OC_MFT_DIELRCP_AVG : OC_MFT_DIELREFL_AVG) + r; size = OC_MFD_SIZE;
  } else return;
  const int n = 1024 + 1; // MicrofacetAlbedoTable_integrateAvg (:10-32)
  float sum = 0.f;
  for (int k = 0; k < n; k++) {
    const float cosThetaO = (float)k / (n - 1);
    sum += interp1D(cosThetaO * (size - 1), dir, size) * cosThetaO * ((k == 0 || k == n - 1) ? 0.5f : 1.f);
  }
  *avg = fminf(2.f * (sum / (n - 1)), 1.f);
}
// HDRILight_calcRowImportance + Distribution1D_create per row (lights/HDRILight.ispc:152-186, math/Distribution1D.ispc:14-42):
// one thread per row of the map
__device__ __forceinline__ float hdriCreateCdf(int size, float *cdf)
{
  float sum = cdf[0];
  for (int i = 1; i < size; i++) cdf[i] = sum += cdf[i];
  const float rcpSum = 1.0f / sum;
  for (int i = 0; i < size; i++) {
    const float c = cdf[i];
    cdf[i] = c >= sum ? 0x1.000002p+0f : c * rcpSum;
  }
  return sum;
}
__global__ void hdriRowImportanceKernel(const OcTexture2D *map, int sizeX, int sizeY, float *cdfX, float *cdfY)
{
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= sizeY) return;
  const float rcpX = 1.f / (float)sizeX, rcpY = 1.f / (float)sizeY;
  const float fy = y * rcpY;
  const float sinTheta = (float)sin((double)fy * 3.14159265358979323846);
  float *row = cdfX + (size_t)y * sizeX;
  for (int x = 0; x < sizeX; x++) {
    const vec4f col = Texture2D_get(map, make_vec2f(x * rcpX, fy), FLT_MIN);
    row[x] = sinTheta * luminance(make_vec3f(col.x, col.y, col.z));
  }
  cdfY[y] = hdriCreateCdf(sizeX, row);
}
__global__ void hdriColumnCdfKernel(int sizeY, float *cdfY) { hdriCreateCdf(sizeY, cdfY); }
__global__ void textureSampleKernel(OcTextureParam tp, const float *st, const float *fp, uint32_t n, float4 *out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const vec4f v = texParamGet4f(tp, make_vec2f(st[2 * i], st[2 * i + 1]), fp[i]);
  out[i] = make_float4(v.x, v.y, v.z, v.w);
}
} // namespace oc

extern "C" int ospcuda_microfacet_albedo_tables(const float **d_tables, uint32_t *numFloats)
{
  static float *tables[32] = {nullptr}; // per CUDA device of the process
  int dev = 0;
  cudaGetDevice(&dev);
  float *&t = tables[dev & 31];
  if (!t) {
    OC_TRY(cudaMalloc(&t, OC_MFT_TOTAL * sizeof(float)));
    cudaStream_t st = (cudaStream_t)ospcuda_stream();
    const uint32_t nDir = OC_MFA_SIZE * OC_MFA_SIZE + 3u * OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE + OC_MFS_SIZE * OC_MFS_SIZE;
    oc::albedoTablesDirKernel<<<(nDir + 127) / 128, 128, 0, st>>>(t);
    const uint32_t nAvg = OC_MFA_SIZE + 3u * OC_MFD_SIZE * OC_MFD_SIZE;
    oc::albedoTablesAvgKernel<<<(nAvg + 127) / 128, 128, 0, st>>>(t);
    oc_count_launch(); oc_count_launch();
    OC_TRY(cudaStreamSynchronize(st));
  }
  if (d_tables) *d_tables = t;
  if (numFloats) *numFloats = OC_MFT_TOTAL;
  return oc_check(cudaGetLastError(), "ospcuda_microfacet_albedo_tables");
}

extern "C" int ospcuda_hdri_distribution(const OcTexture2D *d_map, int32_t sizeX, int32_t sizeY, float *d_cdfX, float *d_cdfY)
{
  if (!d_map || sizeX <= 0 || sizeY <= 0) return oc_set_error(cudaErrorInvalidValue, "ospcuda_hdri_distribution: no map");
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  oc::hdriRowImportanceKernel<<<(sizeY + 63) / 64, 64, 0, st>>>(d_map, sizeX, sizeY, d_cdfX, d_cdfY);
  oc::hdriColumnCdfKernel<<<1, 1, 0, st>>>(sizeY, d_cdfY);
  oc_count_launch(); oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_hdri_distribution");
}

extern "C" int ospcuda_texture_sample(const OcTextureParam *param, const float *d_st, const float *d_footprint, uint32_t n, float *d_out)
{
  if (n == 0) return 0;
  if (!param || !param->tex) return oc_set_error(cudaErrorInvalidValue, "ospcuda_texture_sample: no texture");
  oc::textureSampleKernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)ospcuda_stream()>>>(*param, d_st, d_footprint, n, (float4 *)d_out);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_texture_sample");
}

extern "C" int ospcuda_trace(const OcWorld *d_world, OcRayHit *d_rays, uint32_t numRays, int occlusionOnly)
{
  if (numRays == 0) return 0;
  traceKernel<<<(numRays + 127) / 128, 128, 0, (cudaStream_t)ospcuda_stream()>>>(d_world, d_rays, numRays, occlusionOnly);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_trace");
}
