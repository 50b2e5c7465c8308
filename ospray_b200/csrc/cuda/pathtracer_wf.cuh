// Wavefront path tracer: the pathtracer renderer as GEN / TRACE / SHADE kernels over compacted path and ray queues.
//
// Same algorithm, same sampler call order and the same per-path arithmetic as the reference (render/pathtracer/
// PathTracer.ispc:75-249, PathSampler.ispc:109-415, NextEventEstimation.ispc:25-160, TransparentShadow.ispc:18-85,
// VirtualLight.ispc:38-107); only the schedule differs:
//
//   GEN                 one thread per pixel sample: sampler init, pixel filter, camera ray; the path joins the queues
//   TRACE(d)            a persistent-thread RAY QUEUE kernel: every lane walks the two-level BVH for one ray -- a shadow ray
//                       (any hit) emitted at vertex d-1 or the continuation ray (closest hit) -- and, when its ray is done,
//                       fetches the next one from the queue once enough lanes of its warp are idle (warp-aggregated atomic
//                       on a fetch cursor): warps stay full although rays take very different numbers of steps. The kernel
//                       holds nothing but the walk: ray, compact hit (t, u, v, record, instance), stack.
//   SHADE(d)            per path: fold the shadow-ray results of vertex d-1 into the contribution (transparent occluders take
//                       the reference's closest-hit loop, out of line), then everything between two traces of samplePath():
//                       hit -> geometric normal / ids from the leaf record, free-path sampling in volumes, background,
//                       virtual lights, postIntersect, BSDF, NEE light sampling (shadow rays are EMITTED into the next ray
//                       queue), BSDF sampling, roulette; finished paths accumulate into the framebuffer where they end
//   ...                 until maxPathLength
//
// Path state lives in HBM as float4 SoA arrays (128-bit coalesced accesses).
#pragma once

#include <cooperative_groups.h>

namespace oc {
namespace cg = cooperative_groups;

enum : uint32_t {
  PT_DEPTH_MASK = 0xffu, PT_SCAT_SHIFT = 8, PT_SCAT_MASK = 0xff00u,
  PT_STRAIGHT = 1u << 16, PT_SPECTRANS = 1u << 17, PT_FIRSTBOUNCE = 1u << 18, PT_DONE = 1u << 19, PT_HAS_SHADOW = 1u << 20
};
#define PT_DEAD 0xffffffffu       // flags of a path that TRACE finished (its queue entry remains and is skipped)
#define PT_SHADOW_RAY 0x80000000u // ray-queue ids: bit 31 set = shadow entry [slot * nls + i], clear = continuation ray of path `slot`
#define PT_Q_DONE 0x40000000u     // path-queue entries: the path has ended and only waits for its shadow rays (no continuation ray)

struct PtScratch
{
  float4 *rayO;   // org.xyz, t (far)
  float4 *rayD;   // dir.xyz, t0
  float4 *hit0;   // t, u, v, primID
  float4 *hit1;   // Ng.xyz (instance space, unnormalised), geomID
  uint32_t *hitInst;
  float4 *thr;    // throughput.xyz, sample alpha
  float4 *con;    // contribution.xyz, lastVertex.pdf_w
  float4 *lastP;  // lastVertex.P.xyz, lastVertex.numLightSamples
  float4 *lastN;  // lastVertex.Ns.xyz
  uint2 *smp;     // LDSampler revIndex, scramble
  uint32_t *flags;
  float4 *sZ;     // this sample: z, primID, geomID, instID
  float4 *sPos, *sFN, *sNrm, *sAlb; // this sample's AOVs (only when the framebuffer has the channel)
  float4 *shA, *shB, *shC;          // shadow entries [slot * nls + i]: org.xyz,tmax | dir.xyz,mis | contrib.xyz,valid
  uint2 *shHit;                     // what TRACE found for shadow entry [slot * nls + i]: occluder (record, instance); instance ~0 = none
  float4 *accC, *accZ, *accPos, *accFN, *accNrm, *accAlb; // ScreenSample sums across the spp passes (spp > 1)
  float2 *cone;   // frames with textures: the path's ray cone (width at the last vertex, dwdt) (common/RayCone.ih)
  float4 *med;    // extended materials: the medium the path is in (attenuation, ior) (PathStructs.ih:78 currentMedium)
  float4 *shMed;  //   and the one its pending shadow rays started in (NextEventEstimation.ispc:123)
  uint4 *rnd;     // worlds with volumes: pathState.randomSampler (PCG32 state lo/hi, stream)
  float *shT0;    // worlds with volumes: t0 of shadow entry [slot * nls + i] after the transparent-shadow loop
  uint32_t *pathQ[2], *shadowQ[2];
  uint32_t *counters;
  uint32_t nls;
};

struct PtArgs
{
  FrameArgs a;
  PtScratch s;
  uint32_t numPaths;   // slots in this batch
  uint32_t firstTile;  // first owned tile of the batch
  uint32_t nFirst, nIndirect;
  int pass;            // sample index within the pixel (0 .. spp-1)
  uint32_t depth;
  // queues of this launch: SHADE reads pathIn and appends to pathOut / shadowOut; TRACE reads shadowIn, then the
  // continuation rays of the paths in pathIn (entries without PT_Q_DONE)
  const uint32_t *pathIn, *shadowIn;
  const uint32_t *cntPathIn, *cntShadowIn; // cntShadowIn == nullptr: no shadow rays (depth 0)
  uint32_t *pathOut, *shadowOut;
  uint32_t *cntPathOut, *cntShadowOut;
  uint32_t *fetchCursor;  // TRACE's work-fetch cursor (zeroed with the counters every pass)
  uint32_t vol;           // the world holds volumes (host copy of OcWorld.hasVolumes): free-path sampling + transmittance
  uint32_t progressUnit, progressTotal; // ospGetProgress: what the SHADE / TAIL launch of this stage reports when it starts
  uint32_t shadowT0InB;   // any-hit rays start at shB[e].w instead of 0 (ao renderer: the ambient-occlusion rays start at dg.epsilon)
  // pathtracer with ONE light sample per vertex, a world without volumes whose surfaces are all opaque occluders: the TRACE
  // lane that answers a shadow ray folds the answer into its path and finishes a path that only waited for it (its entry in
  // the path queue stays and is skipped by the next SHADE through the PT_DEAD flags word)
  uint32_t foldInTrace;
  uint32_t sortShade;   // extended-material frames: SHADE sorts each CTA's queue entries by material class (OSPCUDA_PT_SORT_SHADE)
};

__device__ __forceinline__ void ptAppend(uint32_t *q, uint32_t *cnt, uint32_t p)
{
  cg::coalesced_group g = cg::coalesced_threads();
  uint32_t base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(cnt, g.size());
  base = g.shfl(base, 0);
  q[base + g.thread_rank()] = p;
}

__device__ __forceinline__ void ptSlotToPixel(const PtArgs &k, uint32_t slot, uint32_t &x, uint32_t &y)
{
  const uint32_t ownedTile = k.firstTile + (slot >> 12), taskInTile = (slot >> 6) & 63u, pin = slot & 63u;
  const uint32_t tileID = k.a.fb.rank + ownedTile * k.a.fb.numRanks;
  const uint32_t tx = tileID % k.a.tilesX, ty = tileID / k.a.tilesX;
  x = tx * 64 + (taskInTile & 7u) * 8 + (pin & 7u);
  y = ty * 64 + (taskInTile >> 3) * 8 + (pin >> 3);
}

__device__ __forceinline__ void ptSamplerParams(const PtArgs &k, uint32_t ix, uint32_t iy, uint32_t &seed, uint32_t &sampleOffs,
    uint32_t &sampleOffsLight, uint32_t &bits, uint32_t &bitsLight)
{
  // PathTracer.ispc:126-147
  const OcFrameBuffer &fb = k.a.fb;
  seed = 0x49c5f37eu;
  if (fb.targetFrames) {
    const uint32_t targetSpp = (uint32_t)fb.targetFrames * k.a.ren.spp;
    sampleOffs = OwenScramble4(interleaveZOrder16(ix, iy), seed) * targetSpp;
    sampleOffsLight = sampleOffs * k.nFirst;
    bits = 16 + 32 - __clz(targetSpp - 1);
    bitsLight = 48 - __clz(targetSpp * k.nFirst - 1);
  } else {
    seed ^= (uint32_t)fb.sizeX * iy + ix;
    sampleOffs = 0; sampleOffsLight = 0; bits = 16; bitsLight = 16;
  }
}

// a path ended: clamp (pathtracer), fold into the pixel's ScreenSample (spp > 1) and accumulate into the framebuffer
template <bool PT = true>
__device__ __forceinline__ void ptFinishPath(const PtArgs &k, uint32_t p, vec3f contribution, float alpha)
{
  const PtScratch &s = k.s;
  const OcFrameBuffer &fb = k.a.fb;
  uint32_t x, y;
  ptSlotToPixel(k, p, x, y);
  ScreenSample sm;
  sm.rgb = contribution;
  sm.alpha = alpha;
  if (PT) {
    if (isnan(contribution.x) || isnan(contribution.y) || isnan(contribution.z)) { sm.rgb = make_vec3f(0.f); sm.alpha = 1.0f; }
    sm.rgb = vmin(sm.rgb, make_vec3f(k.a.ren.maxRadiance));
  }
  const float4 z = s.sZ[p];
  sm.z = z.x; sm.primID = __float_as_uint(z.y); sm.geomID = __float_as_uint(z.z); sm.instID = __float_as_uint(z.w);
  sm.position = sm.firstNormal = sm.normal = sm.albedo = make_vec3f(0.f);
  if (fb.position) { const float4 v = s.sPos[p]; sm.position = make_vec3f(v.x, v.y, v.z); }
  if (fb.firstNormal) { const float4 v = s.sFN[p]; sm.firstNormal = make_vec3f(v.x, v.y, v.z); }
  if (fb.normal) { const float4 v = s.sNrm[p]; sm.normal = make_vec3f(v.x, v.y, v.z); }
  if (fb.albedo) { const float4 v = s.sAlb[p]; sm.albedo = make_vec3f(v.x, v.y, v.z); }
  const int spp = k.a.ren.spp;
  if (spp > 1) {
    ScreenSample acc;
    if (k.pass == 0) ScreenSample_zero(acc);
    else {
      const float4 c = s.accC[p], az = s.accZ[p];
      acc.rgb = make_vec3f(c.x, c.y, c.z); acc.alpha = c.w;
      acc.z = az.x; acc.primID = __float_as_uint(az.y); acc.geomID = __float_as_uint(az.z); acc.instID = __float_as_uint(az.w);
      acc.position = acc.firstNormal = acc.normal = acc.albedo = make_vec3f(0.f);
      if (fb.position) { const float4 v = s.accPos[p]; acc.position = make_vec3f(v.x, v.y, v.z); }
      if (fb.firstNormal) { const float4 v = s.accFN[p]; acc.firstNormal = make_vec3f(v.x, v.y, v.z); }
      if (fb.normal) { const float4 v = s.accNrm[p]; acc.normal = make_vec3f(v.x, v.y, v.z); }
      if (fb.albedo) { const float4 v = s.accAlb[p]; acc.albedo = make_vec3f(v.x, v.y, v.z); }
    }
    ScreenSample_accumulate(acc, sm);
    if (k.pass < spp - 1) {
      s.accC[p] = make_float4(acc.rgb.x, acc.rgb.y, acc.rgb.z, acc.alpha);
      s.accZ[p] = make_float4(acc.z, __uint_as_float(acc.primID), __uint_as_float(acc.geomID), __uint_as_float(acc.instID));
      if (fb.position) s.accPos[p] = make_float4(acc.position.x, acc.position.y, acc.position.z, 0.f);
      if (fb.firstNormal) s.accFN[p] = make_float4(acc.firstNormal.x, acc.firstNormal.y, acc.firstNormal.z, 0.f);
      if (fb.normal) s.accNrm[p] = make_float4(acc.normal.x, acc.normal.y, acc.normal.z, 0.f);
      if (fb.albedo) s.accAlb[p] = make_float4(acc.albedo.x, acc.albedo.y, acc.albedo.z, 0.f);
      return;
    }
    sm = acc;
    ScreenSample_normalize(sm, 1.f / (float)spp);
  }
  accumulateSample(fb, x, y, sm);
}

__device__ __forceinline__ void ptCountRays(const OcFrameBuffer &fb, uint32_t a, uint32_t b, uint32_t c)
{
  if (!fb.rayCounters) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (a) atomicAdd(fb.rayCounters + 0, (unsigned long long)a);
    if (b) atomicAdd(fb.rayCounters + 1, (unsigned long long)b);
    if (c) atomicAdd(fb.rayCounters + 2, (unsigned long long)c);
  }
}

#define PT_TRACE_BLOCK 128
__device__ __forceinline__ void ptPrefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- volumes under the pathtracer ------------------------------------------------------------------------------------
// pathtracer/volumes/VolumeSampler.ispc:13-282 in its single-volume-interval form (what the reference compiles for its GPU
// target) and pathtracer/volumes/HenyeyGreenstein.ih:11-74. The structured-volume stand-in of Open VKL reports ONE interval
// with the value range (-inf, inf) to this renderer, i.e. delta tracking runs against the global majorant
// densityScale * max opacity (oracle.cpp deltaTracking is the same statement). Everything volume-related sits behind
// out-of-line functions that load and store the path's RandomSampler themselves: worlds without volumes (k.vol == 0)
// never touch them and the register allocation of the TRACE / SHADE kernels is unchanged.
__device__ __forceinline__ RandomSampler ptLoadRs(const PtScratch &s, uint32_t p)
{
  const uint4 v = s.rnd[p];
  RandomSampler rs;
  rs.state = ((uint64_t)v.y << 32) | v.x; rs.stream = v.z;
  return rs;
}
__device__ __forceinline__ void ptStoreRs(const PtScratch &s, uint32_t p, const RandomSampler &rs)
{
  s.rnd[p] = make_uint4((uint32_t)rs.state, (uint32_t)(rs.state >> 32), rs.stream, 0u);
}
// LinearTransferFunction_getMaxOpacity over (-inf, inf) = table entry [0][31] of precomputeMaxOpacityRanges
// (LinearTransferFunction.cpp:110-160) = the largest opacity
__device__ __forceinline__ float ptTfMaxOpacityAll(const OcTransferFunction &tf)
{
  float m = __ldg(tf.opacity);
  for (uint32_t i = 1; i < tf.numOpacities; i++) m = fmaxf(m, __ldg(tf.opacity + i));
  return m;
}
// delta_tracking (VolumeSampler.ispc:13-103): returns (albedo.xyz, t) with t = INFINITY when the ray leaves unscattered
__device__ __forceinline__ float4 ptDeltaTracking(const OcVolume &v, float lower, float upper, RandomSampler &rs, vec3f o, vec3f w)
{
  const float4 none = make_float4(0.f, 0.f, 0.f, INFINITY);
  if (!(lower < upper)) return none;
  float t = lower;
  const float mu_max = v.densityScale * ptTfMaxOpacityAll(v.tf);
  if (mu_max <= 0.f) return none;
  while (true) {
    float xi = RandomSampler_getFloat(rs);
    const float dt = -logf(1.f - xi) / mu_max;
    t += dt;
    if (t > upper) break;
    xi = RandomSampler_getFloat(rs);
    const float sample = volumeSample(v, o + t * w);
    if (isnan(sample)) continue;
    const vec4f color = transferFunctionGet(v.tf, sample);
    const float mu_t = v.densityScale * color.w;
    if (xi < mu_t / mu_max) return make_float4(color.x, color.y, color.z, t);
  }
  return none;
}
// the clipped interval both volume queries track over (VolumeSampler.ispc:139-171 / :222-254); false: nothing to track
__device__ __forceinline__ bool ptVolumeInterval(const OcWorld *world, const Ray &ray, float firstLower, float rayLower, float rayUpper,
    VolumeHit &vh, vec3f &org, vec3f &dir, float &lower, float &upper)
{
  if (!traceVolumeRay(world, ray, vh)) return false;
  lower = fmaxf(firstLower, vh.t0);
  upper = hadHit(ray) ? fminf(vh.t1, ray.t) : vh.t1;
  if (lower > upper) return false;
  org = ray.org; dir = ray.dir;
  if (!vh.inst->isIdentity) {
    const affine3f rx = toA(vh.inst->rcp_xfm);
    org = xfmPoint(rx, ray.org); dir = xfmVector(rx, ray.dir);
  }
  if (upper < rayLower || rayUpper < lower) return false;
  lower = fmaxf(rayLower, lower); upper = fminf(rayUpper, upper);
  return true;
}
struct PtFreePath { float4 albedoDist; uint32_t geomID, instID; };
// volumeSampleFreePath (VolumeSampler.ispc:105-196) for the ray TRACE just resolved against the geometry
static __device__ __noinline__ PtFreePath ptVolumeFreePath(const PtArgs &k, uint32_t p, vec3f rorg, vec3f rdir, float t0, float t,
    uint32_t hitGeomID, float rayLower, float rayUpper)
{
  PtFreePath r;
  r.albedoDist = make_float4(0.f, 0.f, 0.f, INFINITY); r.geomID = r.instID = OC_INVALID_ID;
  Ray ray;
  setRay(ray, rorg, rdir, t0, t);
  ray.geomID = hitGeomID; // hadHit() of the geometry query
  VolumeHit vh;
  vec3f org, dir;
  float lower, upper;
  if (!ptVolumeInterval(k.a.world, ray, 0.f, rayLower, rayUpper, vh, org, dir, lower, upper)) return r;
  RandomSampler rs = ptLoadRs(k.s, p);
  r.albedoDist = ptDeltaTracking(*vh.vol, lower, upper, rs, org, dir);
  ptStoreRs(k.s, p, rs);
  r.geomID = vh.geomID; r.instID = vh.instID;
  return r;
}
// volumeTransmittance (VolumeSampler.ispc:198-279) with the RandomSampler handed in
__device__ __forceinline__ float ptVolumeTransmittance(const OcWorld *world, const Ray &ray, float rayLower, float rayUpper, RandomSampler &rs)
{
  VolumeHit vh;
  vec3f org, dir;
  float lower, upper;
  if (!ptVolumeInterval(world, ray, ray.t0, rayLower, rayUpper, vh, org, dir, lower, upper)) return 1.f;
  const float4 r = ptDeltaTracking(*vh.vol, lower, upper, rs, org, dir);
  return (r.w < upper) ? 0.f : 1.f;
}
// nextEventEstimation's volume part for one unoccluded shadow ray (NextEventEstimation.ispc:126-151): T and an independent
// T_mis; returns the factor that replaces the surface-only MIS weight, T * misHeuristic(ls.pdf, fe.pdf * T_mis)
__device__ __forceinline__ float ptMis(float pdf1, float pdf2)
{
  const float q = sqr(pdf1) * (1.f / (sqr(pdf1) + sqr(pdf2)));
  return pdf1 > 1e17f ? 1.0f : q;
}
static __device__ __noinline__ float ptVolumeShadowWeight(const PtArgs &k, uint32_t p, vec3f org, vec3f dir, float t0, float tmax,
    float lsPdf, float fePdf)
{
  Ray sr;
  setRay(sr, org, dir, t0, tmax); // no hit: the state transparentShadow leaves behind when the light is reached
  RandomSampler rs = ptLoadRs(k.s, p);
  float w = 0.f;
  const float T = ptVolumeTransmittance(k.a.world, sr, 0.f, tmax, rs);
  if (T > 0.f) {
    const float T_mis = ptVolumeTransmittance(k.a.world, sr, 0.f, tmax, rs);
    w = T * ptMis(lsPdf, fePdf * T_mis);
  }
  ptStoreRs(k.s, p, rs);
  return w;
}
__device__ __forceinline__ float hgPhase(float cosTheta, float g) // HenyeyGreenstein.ih:11-15
{
  const float denom = 1.f + g * g - 2.f * g * cosTheta;
  return (1.f / (4.f * (float)M_PI)) * (1.f - g * g) * (1.f / (denom * sqrtf(denom)));
}
__device__ __forceinline__ bool hgIsSmooth(float g) { return !(fabsf(g) > 1.f - 1e-3f); }
__device__ __forceinline__ EvalRes HenyeyGreenstein_eval(float g, vec3f wo, vec3f wi) // :24-35
{
  EvalRes r;
  r.pdf = hgPhase(dot(neg(wo), wi), g);
  r.value = make_vec3f(r.pdf);
  return r;
}
static __device__ __noinline__ SampleRes HenyeyGreenstein_sample(float g, vec3f wo, vec2f s) // :37-72
{
  SampleRes res;
  res.weight = make_vec3f(1.f);
  float cosTheta;
  if (!hgIsSmooth(g)) {
    res.wi = g > 0 ? neg(wo) : wo;
    res.pdf = INFINITY;
    res.type = g > 0 ? SC_SPECULAR_TRANSMISSION : SC_SPECULAR_REFLECTION;
    return res;
  } else if (fabsf(g) < 1e-3f) {
    cosTheta = 1.f - 2.f * s.x;
  } else {
    const float sqrTerm = (1.f - g * g) * (1.f / (1.f - g + 2.f * g * s.x));
    cosTheta = (1.f + g * g - sqrTerm * sqrTerm) * (1.f / (2.f * g));
  }
  res.type = SC_SMOOTH;
  const float sinTheta = sqrtf(fmaxf(0.f, 1.f - cosTheta * cosTheta));
  const float phi = (float)(2.0 * 3.14159265358979323846 * (double)s.y); // M_PI is a double in the reference's scalar form
  res.wi = frame(neg(wo)) * cartesian(phi, sinTheta, cosTheta);
  res.pdf = hgPhase(cosTheta, g);
  return res;
}

// Does the occluder an any-hit query found (leaf record `rec` of instance `instID`) certainly take the shadow-ray
// contribution to exactly 0, as the reference's closest-hit loop would (TransparentShadow.ispc:18-85: lightContrib *=
// transparency per hit, result 0 as soon as reduce_max(lightContrib) <= minContribution)?  OBJ_getTransparency
// (OBJ.ispc:93-111) is T = Tf*d + (1-d) with d = material.d * dg.color.w.
//  * no colour alpha, d == 1, Tf == 0: T == 0 exactly.
//  * vec4 vertex colours whose alphas are ALL 1.0f (cornell_box): the interpolated alpha is 1 up to a few ulp
//    (barycentric / bilinear weights sum to 1 +- 4 ulp), so |T| <= 1e-6 and the product is <= maxContrib * 1e-6, which the
//    reference zeroes whenever that is <= minContribution. Checked here with the actual contribution.
// Anything else (per-primitive model colours, real alpha, thin glass, ...) answers "no" and takes the reference loop.
__device__ __forceinline__ bool ptOccluderIsOpaque(const OcWorld *world, const OcRenderer &ren, uint32_t rec, uint32_t instID, float maxContrib)
{
  const uint32_t so = world->shadowOpaque; // the host looked at every model and material of the world already
  if (so == 2u) return true;
  if (so == 1u) return maxContrib * 1e-6f <= ren.minContribution;
  uint32_t primID, geomID;
  resolveHitIds(world, rec, instID, primID, geomID);
  const OcInstance *inst = world->instances + instID;
  const OcModel *m = inst->group->models + geomID;
  if (m->color.addr) return false;
  const uint32_t gf = m->geom->flags;
  if (gf & OC_GEOMFLAG_HAS_ALPHA) {
    if (!(gf & OC_GEOMFLAG_ALPHA_ALL_ONE) || !(maxContrib * 1e-6f <= ren.minContribution)) return false;
  }
  const OcMaterial *mat = getMaterial(world, ren, m, primID);
  if (mat && (mat->type == OC_MATERIAL_THINGLASS || mat->cutout)) return false; // map_d / the alpha of map_kd may cut the surface out
  if (mat && mat->type == OC_MATERIAL_LUMINOUS) return mat->transparency == 0.f; // Luminous_getTransparency
  if (mat && mat->type == OC_MATERIAL_GLASS) return false;
  if (mat && mat->type == OC_MATERIAL_PRINCIPLED) {
    const OcPrincipled &q = mat->principled;
    if (q.maps) return false;
    // Principled_getTransparency's early exit (Principled.ispc:452-454)
    if (1.f - clampf(q.opacity, 0.f, 1.f) <= PR_EPS && clampf(q.transmission, 0.f, 1.f) <= PR_EPS) return true;
    // solid ROUGH glass lets no shadow ray through either (:477-483: T stays 0 unless roughness < EPS or eta == 1), whichever
    // side the ray comes from; with opacity 1 the cut-out term is exactly 0 too
    if (!q.thin && q.opacity >= 1.f && !(q.roughness < PR_EPS)) {
      float ior = q.ior;
      if (ior < 1.f) ior = 1.f / ior;
      const float etaIn = clampf(q.outsideIor * (1.f / ior), 1.f / 3.f, 3.f), etaOut = clampf(ior * (1.f / q.outsideIor), 1.f / 3.f, 3.f);
      if (fabsf(etaIn - 1.f) > PR_EPS && fabsf(etaOut - 1.f) > PR_EPS) return true;
    }
    return false;
  }
  if (!mat || mat->type != OC_MATERIAL_OBJ) return true; // no / foreign material: transparency 0 (shading.cuh)
  return mat->d == 1.f && mat->Tf.x == 0.f && mat->Tf.y == 0.f && mat->Tf.z == 0.f;
}

// The reference's closest-hit transparent-shadow loop (TransparentShadow.ispc:18-85), taken only when the any-hit query of
// TRACE found a non-opaque occluder (or the contribution is not finite, in which case no query was made). Out of line:
// opaque scenes never take it. Returns the attenuated contribution in xyz and the number of additional shadow segments
// traced in w.
template <bool TEX>
static __device__ __noinline__ float4 ptTransparentShadowSlow(const PtArgs &k, size_t e, float cx, float cy, float cz, uint32_t *stackBase,
    uint32_t stackStride)
{
  uint32_t shadowCount = 0;
  vec3f lightContrib = make_vec3f(cx, cy, cz);
  const float4 a = k.s.shA[e], b = k.s.shB[e];
  Ray sr;
  setRay(sr, make_vec3f(a.x, a.y, a.z), make_vec3f(b.x, b.y, b.z), 0.f, a.w);
  const OcRenderer &ren = k.a.ren;
  uint32_t maxDepth = ren.maxDepth;
  const float tOriginal = sr.t;
  bool counted = isfinite(lightContrib.x) && isfinite(lightContrib.y) && isfinite(lightContrib.z); // the any-hit query was this segment
  // the shadow cone (NextEventEstimation.ispc:103-106) starts from `sample.rayCone.width` of a ScreenSample that
  // PathTraceIntegrator_Li zero-initialises WITHOUT its rayCone (ScreenSample.ih:40-57, PathTracer.ispc:43): the reference
  // reads an indeterminate value there; this device defines it as 0 (finest MIP level along shadow rays)
  RayCone shadowCone;
  shadowCone.width = 0.f; shadowCone.dwdt = 0.f;
  Medium medium = mediumVacuum();
  if (TEX) { const float4 m4 = k.s.shMed[e / k.s.nls]; medium.attenuation = make_vec3f(m4.x, m4.y, m4.z); medium.ior = m4.w; }
  while (1) {
    if (!counted) shadowCount++;
    counted = false;
    const Hit h = traceHitNI<0>(k.a.world, sr.org, sr.dir, sr.t0, sr.t, stackBase, stackStride);
    resolveHit(k.a.world, h, sr);
    if (!hadHit(sr)) break;
    DifferentialGeometry dg;
    postIntersect<TEX>(k.a.world, ren, dg, sr, DG_NS | DG_NG | DG_FACEFORWARD | DG_NORMALIZE | DG_TEXCOORD | DG_COLOR, TEX ? &shadowCone : nullptr);
    lightContrib = lightContrib * materialTransparency<TEX>(dg, sr.dir, medium);
    shadowCone.width = 0.f;
    if (TEX && (medium.attenuation.x != 0.f || medium.attenuation.y != 0.f || medium.attenuation.z != 0.f)) // Beer's law (:66-69)
      lightContrib = lightContrib * expf3(medium.attenuation * (sr.t - sr.t0));
    if (reduce_max(lightContrib) <= ren.minContribution || --maxDepth <= 0) { lightContrib = make_vec3f(0.f); break; }
    if (TEX) materialSelectNextMedium(dg, medium);
    setRay(sr, sr.org, sr.dir, sr.t + dg.epsilon, tOriginal);
  }
  if (k.vol) k.s.shT0[e] = sr.t0; // where the ray stands after the transparent surfaces it crossed (volumeTransmittance starts there)
  return make_float4(lightContrib.x, lightContrib.y, lightContrib.z, __uint_as_float(shadowCount));
}

// evaluateVirtualLights in a world with volumes (VirtualLight.ispc:76-101): the MIS weight of a light the path runs into
// carries a transmittance estimate along the segment since the last vertex
static __device__ __noinline__ vec3f ptVirtualLightsVol(const PtArgs &k, uint32_t p, bool straightPath, vec3f throughput, vec3f lastP,
    vec3f lastNs, float lastPdf, uint32_t lastNumLightSamples, bool environment, vec3f rorg, vec3f rdir, float rt)
{
  const OcWorld *world = k.a.world;
  vec3f L = make_vec3f(0.f);
  const uint32_t numLights = world->numLights, numGeoLights = world->numGeoLights;
  if (numLights <= numGeoLights) return L;
  const float lower = length(lastP - rorg);
  const float upper = environment ? INFINITY : length(lastP - (rorg + rt * rdir));
  const float selectionPDF = lastNumLightSamples / (float)numLights;
  RandomSampler rs = ptLoadRs(k.s, p);
  for (uint32_t i = numGeoLights; i < numLights; i++) {
    const OcLight &l = world->lights[i];
    if (!straightPath || l.isVisible) {
      const LightEval le = l.type == OC_LIGHT_HDRI ? HDRILight_eval(l, rdir, upper)
          : l.type == OC_LIGHT_CYLINDER ? CylinderLight_eval(l, lastP, rdir, lower, upper) : Light_eval(l, lastP, lastNs, rdir, lower, upper);
      if (reduce_max(le.radiance) > 0.0f) {
        float misWeight = 1.0f;
        if (selectionPDF) {
          Ray shadowRay;
          setRay(shadowRay, lastP, rdir, lower, upper);
          const float T = ptVolumeTransmittance(world, shadowRay, lower, upper, rs);
          misWeight = ptMis(lastPdf * T, selectionPDF * le.pdf);
        }
        L = L + throughput * le.radiance * misWeight;
      }
    }
  }
  ptStoreRs(k.s, p, rs);
  return L;
}

// ---- GEN --------------------------------------------------------------------------------------------------------------
// PathTracer_renderPixel, one sample (PathTracer.ispc:149-193) up to the camera ray + samplePath()'s prologue
// (PathSampler.ispc:116-133). Pixels outside the image and pixels of converged render tasks (adaptive accumulation,
// LocalFB.cpp:245-257) start no path.
template <bool VOL, bool TEX>
__global__ void __launch_bounds__(128) ptGenKernel(const __grid_constant__ PtArgs k)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcFrameBuffer &fb = k.a.fb;
  uint32_t primary = 0;
  if (frameCancelled(fb)) return;
  if (fb.control && blockIdx.x == 0 && threadIdx.x == 0) fb.control[2] = k.progressTotal;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < k.numPaths; p += gridDim.x * blockDim.x) {
    uint32_t ix, iy;
    ptSlotToPixel(k, p, ix, iy);
    if (ix >= (uint32_t)fb.sizeX || iy >= (uint32_t)fb.sizeY) continue;
    if (fb.taskError && !(fb.taskError[(iy >> 3) * (((uint32_t)fb.sizeX + 7u) >> 3) + (ix >> 3)] > fb.errorThreshold)) continue;
    uint32_t seed, sampleOffs, sampleOffsLight, bits, bitsLight;
    ptSamplerParams(k, ix, iy, seed, sampleOffs, sampleOffsLight, bits, bitsLight);
    const uint32_t sampleID = (uint32_t)fb.frameID * ren.spp + k.pass;
    LDSampler ld;
    LDSampler_init(ld, seed, sampleOffs + sampleID, sampleOffsLight + sampleID * k.nFirst, bits, bitsLight);
    const vec2f rcpSize = make_vec2f(1.f / (float)fb.sizeX, 1.f / (float)fb.sizeY);
    const vec2f pixel_center = make_vec2f(((float)ix + 0.5f) * rcpSize.x, ((float)iy + 0.5f) * rcpSize.y);
    uint32_t ss234[3];
    const vec2f pixelSample = LDSampler_getNext5Samples(ld, ss234);
    const vec2f pfSample = pixelFilterSample(ren, pixelSample);
    const vec2f screen = pixel_center + pfSample * rcpSize;
    vec2f lens = make_vec2f(0.5f, 0.5f);
    if (k.a.cam.needLensSample) lens = LDSampler_finalizeDim23(ld, ss234);
    Ray ray;
    cameraInitRay(k.a.cam, ray, screen, lens, pixel_center);
    if (TEX) { // ray cone in pixel units (PathTracer.ispc:182-187) and the optional depth limit (:189-192)
      RayCone cone = cameraInitRayCone(k.a.cam, screen, pixel_center);
      cone.dwdt *= rcpSize.y; cone.width *= rcpSize.y;
      s.cone[p] = make_float2(cone.width, cone.dwdt);
      s.med[p] = make_float4(0.f, 0.f, 0.f, 1.f); // make_Medium_vacuum (PathTracer.ispc:56)
      ray.t = fminf(ray.t, rendererMaxDepth(ren, screen));
    }
    s.smp[p] = make_uint2(ld.revIndex, ld.scramble);
    if (VOL) { // pathState.randomSampler (PathTracer.ispc:157)
      RandomSampler rs;
      RandomSampler_init(rs, seed, sampleOffs + sampleID);
      ptStoreRs(s, p, rs);
    }
    s.flags[p] = PT_STRAIGHT | PT_SPECTRANS | PT_FIRSTBOUNCE;
    s.thr[p] = make_float4(1.f, 1.f, 1.f, 1.f);
    s.con[p] = make_float4(0.f, 0.f, 0.f, INFINITY);
    s.lastP[p] = make_float4(ray.org.x, ray.org.y, ray.org.z, __uint_as_float(0u));
    s.lastN[p] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, 0.f);
    s.sZ[p] = make_float4(INFINITY, __uint_as_float(OC_INVALID_ID), __uint_as_float(OC_INVALID_ID), __uint_as_float(OC_INVALID_ID));
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fb.position) s.sPos[p] = zero;
    if (fb.firstNormal) s.sFN[p] = zero;
    if (fb.normal) s.sNrm[p] = zero;
    if (fb.albedo) s.sAlb[p] = zero;
    s.rayO[p] = make_float4(ray.org.x, ray.org.y, ray.org.z, ray.t);
    s.rayD[p] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.t0);
    primary++;
    ptAppend(k.pathOut, k.cntPathOut, p); // depth 0: the path queue is the continuation-ray queue too
  }
  ptCountRays(fb, primary, 0u, 0u);
}

// The answer to the one shadow ray of path p (= shadow entry p: one light sample per vertex), folded where it is found:
// unoccluded -> the sample's contribution joins the path's (same sum, same order as the fold in SHADE: there is one term);
// opaque occluder -> nothing to add; anything else (transparent occluder) is left to SHADE's closest-hit loop. A path that
// had ended and only waited for this answer is finished here and marked PT_DEAD.
static __device__ __noinline__ void ptFoldShadowInTrace(const PtArgs &k, uint32_t p, uint32_t rec, uint32_t inst)
{
  const PtScratch &s = k.s;
  const float4 cc = s.shC[p];
  const vec3f lightContrib = make_vec3f(cc.x, cc.y, cc.z);
  const bool miss = inst == OC_INVALID_ID;
  const uint32_t flags = s.flags[p];
  if (!miss && !ptOccluderIsOpaque(k.a.world, k.a.ren, rec, inst, reduce_max(lightContrib))) {
    s.shHit[p] = make_uint2(rec, inst);
    return; // (the path's entry is in the path queue: the next SHADE walks the reference's closest-hit loop and finishes it)
  }
  float4 con4 = s.con[p];
  if (miss && reduce_max(lightContrib) > 0) {
    const float w = s.shB[p].w;
    // L = lightContrib * misWeight, then contribution + L: two roundings, as in SHADE's fold and in the reference
    con4.x = __fadd_rn(con4.x, __fmul_rn(lightContrib.x, w)); con4.y = __fadd_rn(con4.y, __fmul_rn(lightContrib.y, w));
    con4.z = __fadd_rn(con4.z, __fmul_rn(lightContrib.z, w));
  }
  if (flags & PT_DONE) {
    ptFinishPath(k, p, make_vec3f(con4.x, con4.y, con4.z), s.thr[p].w);
    s.flags[p] = PT_DEAD; // its entry in the path queue is skipped by the next SHADE
    return;
  }
  s.con[p] = con4;
  s.flags[p] = flags & ~PT_HAS_SHADOW;
}

// ---- TRACE ------------------------------------------------------------------------------------------------------------
// Persistent-thread ray queue. The rays of one bounce -- the shadow queue, then the continuation rays of the path queue --
// form one index space [0, n). A warp reserves PT_TRACE_CHUNK consecutive indices at a time (one atomic on the fetch cursor
// per chunk) and hands them to its lanes. The walk runs in warp-synchronous steps: every round the warp counts how many of
// its lanes stand at an inner node and how many at a leaf and takes the step the majority wants -- one node test
// (Traversal::inner) or one leaf (Traversal::leaf) -- side by side; the others wait for their turn. Lanes whose walk has
// ended stay idle until REFILL_IDLE of them have gathered: then they write their results together and take the next rays
// of the chunk, so the warp stays filled while individual rays take anything from 2 to 200 steps. A lane holds nothing but
// the walk: instance-space ray, compact hit, stack position; the world-space ray is re-read from the queue's arrays when
// the walk leaves a transformed instance, the ids / geometric normal of the hit are read from its leaf record at the end.
#ifndef PT_TRACE_CHUNK
#define PT_TRACE_CHUNK 128u
#endif
template <int REFILL_IDLE, int MIN_BLOCKS>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, MIN_BLOCKS) ptTraceKernel(const __grid_constant__ PtArgs k)
{
  __shared__ uint32_t s_stack[OC_STACK_SHARED * PT_TRACE_BLOCK];
  const uint32_t nS = k.cntShadowIn ? *k.cntShadowIn : 0u;
  const uint32_t n = nS + *k.cntPathIn;
  if (n == 0 || frameCancelled(k.a.fb)) return;
  uint32_t ovf[OC_STACK_OVF];
  Stack st;
  st.base = s_stack + threadIdx.x; st.stride = PT_TRACE_BLOCK; st.ovf = ovf;
  const PtScratch &s = k.s;
  const OcWorld *__restrict__ world = k.a.world;
  const uint32_t lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;
  uint32_t chunkNext = 0, chunkEnd = 0; // warp-uniform
  bool exhausted = false;               // warp-uniform: the cursor has run past the queue
  // rays a warp reserves at a time: PT_TRACE_CHUNK when the queue is long; when it holds only a few rays per resident warp
  // (late bounces, small tiles of a multi-GPU frame) smaller chunks spread them over all warps instead of leaving most idle
  const uint32_t warpsInGrid = gridDim.x * (PT_TRACE_BLOCK / 32);
  const uint32_t chunk = n >= 2u * PT_TRACE_CHUNK * warpsInGrid ? PT_TRACE_CHUNK : (n >= 128u * warpsInGrid ? 64u : 32u);
  bool active = false, fin = false;     // fin: the walk has ended, the result is not written yet
  uint32_t id = 0;
  uint32_t nShadow = 0, nCont = 0;
  Traversal<2> tr;
  RayMem src;
  src.po = src.pd = nullptr;
  while (true) {
    uint32_t idle = ~__ballot_sync(0xffffffffu, active);
    if (__popc(idle) >= REFILL_IDLE) {
      // ---- results of the walks that ended since the last time ----
      if (fin) {
        fin = false;
        if (id & PT_SHADOW_RAY) {
          if (k.foldInTrace) ptFoldShadowInTrace(k, id & ~PT_SHADOW_RAY, tr.h.rec, tr.h.inst);
          else s.shHit[id & ~PT_SHADOW_RAY] = make_uint2(tr.h.rec, tr.h.inst);
        } else {
          Ray r;
          const float4 o4 = *src.po, d4 = *src.pd;
          r.org = make_vec3f(o4.x, o4.y, o4.z); r.dir = make_vec3f(d4.x, d4.y, d4.z); r.t0 = d4.w;
          resolveHit(world, tr.h, r);
          s.hit0[id] = make_float4(r.t, r.u, r.v, __uint_as_float(r.primID));
          s.hit1[id] = make_float4(r.Ng.x, r.Ng.y, r.Ng.z, __uint_as_float(r.geomID));
          s.hitInst[id] = r.instID;
        }
      }
      // ---- the idle lanes take the next rays ----
      while (idle && !exhausted) {
        if (chunkNext == chunkEnd) {
          uint32_t base = 0;
          if (lane == 0) base = frameCancelled(k.a.fb) ? 0xffffffffu : atomicAdd(k.fetchCursor, chunk);
          base = __shfl_sync(0xffffffffu, base, 0);
          if (base >= n) { exhausted = true; break; } // queue drained (or the frame was cancelled: ospCancel)
          chunkNext = base;
          chunkEnd = min(base + chunk, n);
        }
        const uint32_t avail = chunkEnd - chunkNext;
        const uint32_t rank = __popc(idle & ltMask);
        const bool take = ((idle >> lane) & 1u) && rank < avail;
        if (take) {
          const uint32_t i = chunkNext + rank;
          const bool occ = i < nS;
          id = occ ? k.shadowIn[i] : k.pathIn[i - nS];
          if (occ || !(id & PT_Q_DONE)) { // a path that has ended has no continuation ray: the lane stays idle until the next refill
            if (occ) {
              const size_t e = id & ~PT_SHADOW_RAY;
              src.po = s.shA + e; src.pd = s.shB + e;
              nShadow++;
              if (k.foldInTrace) { // what the fold reads when this walk ends: asked for now, there when needed
                ptPrefetchL2(s.shC + e); ptPrefetchL2(s.flags + e); ptPrefetchL2(s.con + e);
              }
            } else {
              src.po = s.rayO + id; src.pd = s.rayD + id;
              nCont++;
            }
            const float4 o4 = *src.po, d4 = *src.pd;
            active = tr.init(world, make_vec3f(o4.x, o4.y, o4.z), make_vec3f(d4.x, d4.y, d4.z), (occ && !k.shadowT0InB) ? 0.f : d4.w, o4.w, occ);
            fin = !active; // empty world / empty group: the ray misses, written with the next batch of results
          }
        }
        const uint32_t took = __ballot_sync(0xffffffffu, take);
        chunkNext += __popc(took);
        idle &= ~took;
      }
    }
    const uint32_t nActive = __popc(__ballot_sync(0xffffffffu, active));
    if (nActive == 0) {
      if (__ballot_sync(0xffffffffu, fin)) continue; // rays answered on the spot: write them, take the next
      if (exhausted) break;                           // nothing in flight, nothing left to fetch
      continue;                                       // the entries just taken held no ray (ended paths): take the next
    }
    // ---- one step, the kind most lanes want ----
    const bool wantInner = active && !tr.atLeaf();
    const uint32_t nInner = __popc(__ballot_sync(0xffffffffu, wantInner));
    if (2u * nInner >= nActive) {
      if (wantInner && !tr.inner(world, src, st)) { active = false; fin = true; }
    } else {
      if (active && !wantInner && !tr.leaf(world, src, st)) { active = false; fin = true; }
    }
  }
  ptCountRays(k.a.fb, 0u, k.depth == 0 ? 0u : nCont, nShadow);
}

// The shadow rays of the previous vertex folded into one radiance term, in sample order. TRACE folds the common case itself
// (ptFoldShadowInTrace); this runs for several light samples per vertex, transparent occluders and worlds with volumes.
// (Inlined: out of line it costs the shade step 1 % on C1, tools/ab_c1.sh.)
template <bool VOL, bool TEX>
__device__ __forceinline__ vec3f ptFoldShadows(const PtArgs &k, uint32_t p, uint32_t *stackBase, uint32_t &extraShadow)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcWorld *world = k.a.world;
  vec3f L = make_vec3f(0.f);
  for (uint32_t j = 0; j < s.nls; j++) {
    const size_t e = (size_t)p * s.nls + j;
    const float4 cc = s.shC[e];
    if (cc.w == 0.f) continue;
    vec3f lightContrib = make_vec3f(cc.x, cc.y, cc.z);
    bool resolved = false;
    float shadowT0 = 0.f;
    if (isfinite(lightContrib.x) && isfinite(lightContrib.y) && isfinite(lightContrib.z)) {
      const uint2 occ = s.shHit[e];
      if (occ.y == OC_INVALID_ID) resolved = true; // unoccluded: lightContrib unchanged
      else if (ptOccluderIsOpaque(world, ren, occ.x, occ.y, reduce_max(lightContrib))) { lightContrib = make_vec3f(0.f); resolved = true; }
    }
    if (!resolved) {
      const float4 r = ptTransparentShadowSlow<TEX>(k, e, lightContrib.x, lightContrib.y, lightContrib.z, stackBase, PT_TRACE_BLOCK);
      lightContrib = make_vec3f(r.x, r.y, r.z);
      extraShadow += __float_as_uint(r.w);
      if (VOL) shadowT0 = s.shT0[e];
    }
    if (reduce_max(lightContrib) > 0) {
      const float4 b = s.shB[e];
      if (VOL) { // b.w = ls.pdf, cc.w = fe.pdf: the weight carries the transmittance estimates
        const float4 a = s.shA[e];
        const float w = ptVolumeShadowWeight(k, p, make_vec3f(a.x, a.y, a.z), make_vec3f(b.x, b.y, b.z), shadowT0, a.w, b.w, cc.w == FLT_MIN ? 0.f : cc.w);
        L = L + w * lightContrib;
      } else
        L = L + lightContrib * b.w;
    }
  }
  return L;
}

// ---- SHADE ------------------------------------------------------------------------------------------------------------
// One path, one vertex: returns 0 when the path finished here, 1 when it goes on to the next bounce, 2 when it has ended
// but shadow rays of this vertex are pending (the next SHADE folds them in and finishes it).
template <bool VOL, bool QUEUE, bool TEX>
__device__ __forceinline__ uint32_t ptShadeStep(const PtArgs &k, uint32_t p, uint32_t *stackBase, uint32_t &extraShadow)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcFrameBuffer &fb = k.a.fb;
  const OcWorld *world = k.a.world;
  Stack dummy;
  dummy.base = nullptr; dummy.stride = 0; dummy.ovf = nullptr;
  Ctx c(k.a.world, k.a.ren, k.a.cam, k.a.fb, dummy);

  uint32_t flags = s.flags[p];
  if (QUEUE && flags == PT_DEAD) return 0u;
  float4 thr4 = s.thr[p], con4 = s.con[p];
  if (flags & PT_HAS_SHADOW) {
    // shadow rays of the previous vertex: transparentShadow (TransparentShadow.ispc:18-85) per light sample, summed in
    // sample order like nextEventEstimation's L (NextEventEstimation.ispc:140-156). TRACE made ONE any-hit query per ray:
    // any opaque occluder makes the product of transparencies exactly 0, whatever lies in front of it; only a non-opaque
    // find takes the reference's closest-hit loop.
    const vec3f L = ptFoldShadows<VOL, TEX>(k, p, stackBase, extraShadow);
    con4.x += L.x; con4.y += L.y; con4.z += L.z;
    if (flags & PT_DONE) {
      ptFinishPath(k, p, make_vec3f(con4.x, con4.y, con4.z), thr4.w);
      return 0u;
    }
  }
  flags &= ~(PT_HAS_SHADOW | PT_DONE);
  const uint32_t depth = flags & PT_DEPTH_MASK;
  uint32_t scatteringEvents = (flags & PT_SCAT_MASK) >> PT_SCAT_SHIFT;
  bool straightPath = flags & PT_STRAIGHT, specularTransmissionPath = flags & PT_SPECTRANS, firstBounceLight = flags & PT_FIRSTBOUNCE;
  Ray ray;
  {
    const float4 o = s.rayO[p], d = s.rayD[p], h0 = s.hit0[p], h1 = s.hit1[p];
    ray.org = make_vec3f(o.x, o.y, o.z); ray.dir = make_vec3f(d.x, d.y, d.z); ray.t0 = d.w;
    ray.t = h0.x; ray.u = h0.y; ray.v = h0.z; ray.primID = __float_as_uint(h0.w);
    ray.Ng = make_vec3f(h1.x, h1.y, h1.z); ray.geomID = __float_as_uint(h1.w);
    ray.instID = s.hitInst[p];
    const float rayUpper = o.w; // traceClippingRay without clipping geometry (World.ih:150-163)
    // worlds with volumes: volumeSampleFreePath (PathSampler.ispc:178-194), a scattering event before the surface replaces the hit
    if (VOL) {
      const PtFreePath fp = ptVolumeFreePath(k, p, ray.org, ray.dir, ray.t0, ray.t, ray.geomID, ray.t0, rayUpper);
      if (fp.albedoDist.w < INFINITY) {
        ray.t = fp.albedoDist.w; ray.u = ray.v = 0.f; ray.primID = 0u;
        ray.Ng = make_vec3f(fp.albedoDist.x, fp.albedoDist.y, fp.albedoDist.z); // the albedo travels in Ng
        ray.geomID = fp.geomID; ray.instID = fp.instID;
        flags |= 0x80000000u; // local marker: volume vertex
      }
    }
  }
  const bool volume = VOL && (flags & 0x80000000u) != 0;
  flags &= 0x7fffffffu;
  vec3f throughput = make_vec3f(thr4.x, thr4.y, thr4.z), contribution = make_vec3f(con4.x, con4.y, con4.z);
  float alpha = thr4.w;
  PathVertexLite last;
  {
    const float4 lp = s.lastP[p], ln = s.lastN[p];
    last.P = make_vec3f(lp.x, lp.y, lp.z); last.Ns = make_vec3f(ln.x, ln.y, ln.z);
    last.pdf_w = con4.w; last.numLightSamples = __float_as_uint(lp.w);
  }
  LDSampler ld;
  vec2f pixelCenter = make_vec2f(0.f, 0.f); // pathContext.screen (PathTracer.ispc:194): where the backplate is looked up
  RayCone cone;
  cone.width = cone.dwdt = 0.f;
  Medium medium = mediumVacuum();
  if (TEX) {
    const float2 c2 = s.cone[p]; cone.width = c2.x; cone.dwdt = c2.y;
    const float4 m4 = s.med[p]; medium.attenuation = make_vec3f(m4.x, m4.y, m4.z); medium.ior = m4.w;
  }
  {
    uint32_t ix, iy, seed, so, sol, bits, bitsLight;
    ptSlotToPixel(k, p, ix, iy);
    if (TEX) pixelCenter = make_vec2f(((float)ix + 0.5f) * (1.f / (float)fb.sizeX), ((float)iy + 0.5f) * (1.f / (float)fb.sizeY));
    ptSamplerParams(k, ix, iy, seed, so, sol, bits, bitsLight);
    const uint32_t sampleID = (uint32_t)fb.frameID * ren.spp + k.pass;
    LDSampler_init(ld, seed, so + sampleID, sol + sampleID * k.nFirst, bits, bitsLight);
    const uint2 sm = s.smp[p];
    ld.revIndex = sm.x; ld.scramble = sm.y;
  }

  // ---- one iteration of samplePath()'s loop, after the trace (PathSampler.ispc:196-412) ----
  bool alive = false, hasShadow = false;
  do {
    const bool surface = !volume && hadHit(ray);
    const bool environment = !surface && !volume;
    const vec3f volumeAlbedo = ray.Ng;
    const OcVolume *vmodel = nullptr;
    if (volume) {
      throughput = throughput * volumeAlbedo; // PathSampler.ispc:189-192
      vmodel = world->instances[ray.instID].group->volumes + ray.geomID;
    }
    const vec3f wo = neg(ray.dir);
    if (depth == 0) {
      float4 z = s.sZ[p];
      z.x = fb.projectedDepth ? ray.t * dot(ray.dir, toV(k.a.cam.direction)) : ray.t;
      s.sZ[p] = z;
    }
    if (environment && (ren.backgroundRefraction ? specularTransmissionPath : straightPath)) {
      const vec4f bg = TEX ? getBackgroundTex(ren, pixelCenter) : getBackground(ren);
      alpha -= luminance(throughput) * (1.0f - bg.w);
      contribution = contribution + throughput * make_vec3f(bg);
    }
    if (!volume) { // PathSampler.ispc:220-228
      if (VOL)
        contribution = contribution + ptVirtualLightsVol(k, p, straightPath, throughput, last.P, last.Ns, last.pdf_w, last.numLightSamples,
                                          environment, ray.org, ray.dir, ray.t);
      else
        contribution = contribution + evaluateVirtualLights<TEX>(c, straightPath, throughput, last, environment, ray);
    }
    if (environment) break;
    // (with geometry lights the vertex is still needed: what the path ran into may emit, PathSampler.ispc:247-254)
    if (depth >= ren.maxDepth && !(TEX && world->numGeoLights)) break;

    DifferentialGeometry dg;
    if (volume) { // postIntersect of a volume vertex (PathSampler.ispc:51-67) + VolumetricModel_postIntersect (VolumetricModel.ih:16-33)
      dg.P = ray.org + ray.t * ray.dir;
      dg.Ng = dg.Ns = dg.dPds = make_vec3f(0.f);
      dg.color = make_vec4f(0.f, 0.f, 0.f, 0.f);
      dg.epsilon = 0.f; dg.material = nullptr; dg.primID = 0;
      dg.objID = vmodel->userID == OC_INVALID_ID ? ray.geomID : vmodel->userID;
      dg.instID = ray.instID;
      if (ray.instID != OC_INVALID_ID && world->instances[ray.instID].userID != OC_INVALID_ID)
        dg.instID = world->instances[ray.instID].userID;
    } else
      postIntersect<TEX>(world, ren, dg, ray,
          DG_NS | DG_NG | DG_FACEFORWARD | DG_NORMALIZE | DG_TEXCOORD | DG_COLOR | DG_TANGENTS | DG_MOTIONBLUR, TEX ? &cone : nullptr);
    if (TEX && !volume) { // evaluateGeometryLights (GeometryLight.ispc:22-58): emission of the surface the path ran into
      const vec3f emission = materialEmission(dg);
      if (reduce_max(emission) > 0.f) {
        float misWeight = 1.0f;
        if (world->numGeoLights > 0 && last.numLightSamples > 0) {
          const float lightSelectionPDF = last.numLightSamples / (float)world->numLights;
          const float cosd = dot(dg.Ng, ray.dir);
          const vec3f vdis = last.P - dg.P;
          const float lePdf = dg.areaPDF * dot(vdis, vdis) / fabsf(cosd);
          misWeight = misHeuristic(last.pdf_w, lightSelectionPDF * lePdf);
        }
        contribution = contribution + throughput * emission * misWeight;
      }
    }
    if (depth == 0) {
      float4 z = s.sZ[p];
      z.y = __uint_as_float(dg.primID); z.z = __uint_as_float(dg.objID); z.w = __uint_as_float(dg.instID);
      s.sZ[p] = z;
      if (fb.position) s.sPos[p] = make_float4(dg.P.x, dg.P.y, dg.P.z, 0.f);
      if (fb.firstNormal) s.sFN[p] = make_float4(dg.Ns.x, dg.Ns.y, dg.Ns.z, 0.f);
    }
    if (depth >= ren.maxDepth || scatteringEvents >= (uint32_t)ren.maxScatteringEvents) break;

    ObjBSDF bsdf;
    ExtBSDF xbsdf; // metal / plastic state (extended-material kernels only)
    if (!volume) {
      if (!(TEX && Ext_getBSDF(bsdf, xbsdf, dg, medium, ray.dir))) OBJ_getBSDF<TEX>(bsdf, dg);
      if (!bsdf.valid) break;
    }

    LDSampler_nextGroup(ld);
    uint32_t vertexLightSamples = 0;
    if (volume ? hgIsSmooth(vmodel->anisotropy) : (bsdf.scatteringType & SC_SMOOTH) != 0) { // isSmooth (PathTracerUtil.ih:15-28)
      if (scatteringEvents == 0) { // updateAuxilliaryData (PathTracerUtil.ih:48-60)
        if (fb.albedo) { const vec3f v = throughput * (volume ? volumeAlbedo : bsdf.albedo); s.sAlb[p] = make_float4(v.x, v.y, v.z, 0.f); }
        if (fb.normal) { const vec3f v = throughput * (volume ? dg.Ns : bsdf.frame.vz); s.sNrm[p] = make_float4(v.x, v.y, v.z, 0.f); }
      }
      // nextEventEstimation up to the shadow ray (NextEventEstimation.ispc:36-118): the ray joins the next TRACE's queue
      const uint32_t numLightSamples = firstBounceLight ? k.nFirst : k.nIndirect;
      vertexLightSamples = numLightSamples;
      const uint32_t numLights = world->numLights;
      for (uint32_t j = 0; j < s.nls; j++) {
        const size_t e = (size_t)p * s.nls + j;
        bool valid = false;
        if (j < numLightSamples) {
          float ss;
          const vec2f sl = LDSampler_get3LightSamples(ld, j, numLightSamples > 1, ss);
          const Sample1D sel = Distribution1D_sample((int)numLights, world->lightsCDF, 0, ss);
          const float lightSelectionProb = sel.prob * numLightSamples;
          const OcLight &light = world->lights[sel.idx];
          LightSample ls = (TEX && light.type == OC_LIGHT_GEOMETRY) ? GeometryLight_sample(world, ren, light, dg, sl)
              : (TEX && light.type == OC_LIGHT_HDRI) ? HDRILight_sample(light, sl)
              : (TEX && light.type == OC_LIGHT_CYLINDER) ? CylinderLight_sample(light, dg, sl) : Light_sample(light, dg, sl);
          ls.weight = divZeroAware(ls.weight, lightSelectionProb); // (rejected samples carry a zero weight: see divZeroAware)
          ls.pdf *= lightSelectionProb;
          if (!(reduce_max(ls.weight) <= 0.0f || ls.pdf <= 0.0f)) {
            const EvalRes fe = volume ? HenyeyGreenstein_eval(vmodel->anisotropy, wo, ls.dir)
                : (TEX && bsdf.kind != BSDF_KIND_OBJ) ? Ext_eval(bsdf, xbsdf, wo, ls.dir) : OBJ_eval(bsdf, wo, ls.dir);
            if (!(reduce_max(fe.value) <= 0.0f)) {
              vec3f org = dg.P;
              if (!volume && dot(dg.Ng, ls.dir) < 0.f) org = org - (2.0f * dg.epsilon) * dg.Ng;
              const vec3f unshaded = throughput * fe.value * ls.weight;
              s.shA[e] = make_float4(org.x, org.y, org.z, ls.dist);
              // worlds with volumes: the fold needs the two pdfs apart (the MIS weight carries a transmittance estimate)
              s.shB[e] = make_float4(ls.dir.x, ls.dir.y, ls.dir.z, VOL ? ls.pdf : misHeuristic(ls.pdf, fe.pdf));
              s.shC[e] = make_float4(unshaded.x, unshaded.y, unshaded.z, VOL ? (fe.pdf > 0.f ? fe.pdf : FLT_MIN) : 1.f);
              valid = true;
              hasShadow = true;
              // a contribution that is not finite takes the closest-hit loop in the fold (the reference's arithmetic on it)
              if (QUEUE && isfinite(unshaded.x) && isfinite(unshaded.y) && isfinite(unshaded.z))
                ptAppend(k.shadowOut, k.cntShadowOut, PT_SHADOW_RAY | (uint32_t)e);
            }
          }
        }
        if (!valid) s.shC[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      firstBounceLight = false;
      if (TEX && hasShadow) s.shMed[p] = make_float4(medium.attenuation.x, medium.attenuation.y, medium.attenuation.z, medium.ior);
    }

    float ss;
    uint32_t uu3;
    const vec2f sb = LDSampler_getNext4Samples(ld, ss, uu3);
    SampleRes fs = volume ? HenyeyGreenstein_sample(vmodel->anisotropy, wo, sb)
        : (TEX && bsdf.kind != BSDF_KIND_OBJ) ? Ext_sample(bsdf, xbsdf, wo, sb, ss) : OBJ_sample(bsdf, wo, sb, ss);
    if (!volume) {
      const bool sameSide = dot(fs.wi, dg.Ng) >= 0.f;
      const bool reflection = fs.type & SC_REFLECTION;
      if (sameSide != reflection) fs.weight = make_vec3f(0.0f);
    }
    if (reduce_max(fs.weight) <= 0.0f || fs.pdf <= 0.0f) break;
    throughput = throughput * fs.weight;
    if (depth >= (uint32_t)ren.rouletteDepth) {
      const float rr = LDSampler_finalizeDim3(ld, uu3);
      const float contProb = fminf(reduce_max(throughput), 0.95f);
      if (rr > contProb) break;
      throughput = throughput * (1.f / contProb);
    }
    if (TEX && reduce_min(medium.attenuation) < 0.f) throughput = throughput * expf3(medium.attenuation * ray.t); // Beer's law (:365-368)
    vec3f ray_org = dg.P;
    if (!volume && (fs.type & SC_TRANSMISSION)) {
      ray_org = ray_org - (2.0f * dg.epsilon) * dg.Ng;
      if (TEX) materialSelectNextMedium(dg, medium); // the path crosses an interface (:371-383)
    }
    if ((fs.type & ~SC_SPECULAR_TRANSMISSION) || volume) {
      last.P = dg.P; last.Ns = dg.Ns; last.pdf_w = fs.pdf; last.numLightSamples = vertexLightSamples;
    }
    if (!eq3(ray.dir, fs.wi)) straightPath = false;
    if (!(fs.type & SC_SPECULAR_TRANSMISSION)) specularTransmissionPath = false;
    s.rayO[p] = make_float4(ray_org.x, ray_org.y, ray_org.z, INFINITY);
    s.rayD[p] = make_float4(fs.wi.x, fs.wi.y, fs.wi.z, 0.f);
    if (fs.type & SC_SMOOTH) scatteringEvents++;
    alive = reduce_max(throughput) > ren.minContribution;
  } while (false);

  if (!alive && !hasShadow) {
    ptFinishPath(k, p, contribution, alpha);
    return 0u;
  }
  s.thr[p] = make_float4(throughput.x, throughput.y, throughput.z, alpha);
  s.con[p] = make_float4(contribution.x, contribution.y, contribution.z, last.pdf_w);
  if (alive) {
    s.lastP[p] = make_float4(last.P.x, last.P.y, last.P.z, __uint_as_float(last.numLightSamples));
    s.lastN[p] = make_float4(last.Ns.x, last.Ns.y, last.Ns.z, 0.f);
    s.smp[p] = make_uint2(ld.revIndex, ld.scramble);
    if (TEX) {
      s.cone[p] = make_float2(cone.width, cone.dwdt);
      s.med[p] = make_float4(medium.attenuation.x, medium.attenuation.y, medium.attenuation.z, medium.ior);
    }
  }
  s.flags[p] = ((depth + 1) & PT_DEPTH_MASK) | (scatteringEvents << PT_SCAT_SHIFT) | (straightPath ? PT_STRAIGHT : 0u)
      | (specularTransmissionPath ? PT_SPECTRANS : 0u) | (firstBounceLight ? PT_FIRSTBOUNCE : 0u) | (hasShadow ? PT_HAS_SHADOW : 0u)
      | (alive ? 0u : PT_DONE);
  return alive ? 1u : 2u;
}

// Does the vertex this path is about to shade carry one of the extended materials (metal ... carPaint)?  Read from the hit the
// way postIntersect will (instance -> model -> material); ended paths and misses answer no.
__device__ __forceinline__ bool ptPathShadesExtendedMaterial(const PtArgs &k, uint32_t entry)
{
  if (entry & PT_Q_DONE) return false;
  const uint32_t p = entry & ~PT_Q_DONE;
  const uint32_t geomID = __float_as_uint(k.s.hit1[p].w);
  if (geomID == OC_INVALID_ID) return false;
  const uint32_t instID = k.s.hitInst[p], primID = __float_as_uint(k.s.hit0[p].w);
  const OcInstance *inst = k.a.world->instances + instID;
  if (geomID >= inst->group->numModels) return false;
  const OcMaterial *mat = getMaterial(k.a.world, k.a.ren, inst->group->models + geomID, primID);
  return mat && mat->type >= OC_MATERIAL_METAL;
}

// MIN_BLOCKS: resident CTAs per SM the register budget is set for (4: 128 registers, 5: 96, 6: 80)
template <bool VOL, int MIN_BLOCKS, bool TEX>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, MIN_BLOCKS) ptShadeKernel(const __grid_constant__ PtArgs k)
{
  __shared__ uint32_t s_stack[OC_STACK_SHARED * PT_TRACE_BLOCK]; // the out-of-line transparent-shadow loop walks the hierarchy
  if (TEX) { // (one answer per CTA: its threads meet at barriers below)
    __shared__ int s_cancelled;
    if (threadIdx.x == 0) s_cancelled = frameCancelled(k.a.fb) ? 1 : 0;
    __syncthreads();
    if (s_cancelled) return;
  } else if (frameCancelled(k.a.fb)) return;
  if (k.a.fb.control && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(k.a.fb.control + 1, k.progressUnit);
  const uint32_t n = *k.cntPathIn;
  uint32_t extraShadow = 0;
  if (TEX) {
    // Extended-material frames: after the first bounce the paths of a warp sit on unrelated surfaces, and a warp that holds
    // ONE principled vertex among 31 OBJ ones runs both material codes one after the other. Every CTA therefore sorts its 128
    // queue entries by material class first -- extended materials to the front lanes, the rest to the back -- so that all but
    // one of its warps shade one class. (Which lane shades which path has no effect on the results: a path owns its pixel.)
    __shared__ uint32_t s_perm[PT_TRACE_BLOCK];
    __shared__ uint32_t s_cnt[2];
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
      if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0u;
      __syncthreads();
      const uint32_t i = base + threadIdx.x;
      if (i < n) {
        const uint32_t entry = k.pathIn[i];
        if (k.sortShade && ptPathShadesExtendedMaterial(k, entry)) s_perm[atomicAdd(&s_cnt[0], 1u)] = entry;
        else s_perm[blockDim.x - 1u - atomicAdd(&s_cnt[1], 1u)] = entry;
      }
      __syncthreads();
      const bool mine = threadIdx.x < s_cnt[0] || threadIdx.x >= blockDim.x - s_cnt[1];
      const uint32_t entry = mine ? s_perm[threadIdx.x] : 0u;
      __syncthreads(); // s_perm / s_cnt are rewritten by the next chunk
      if (mine) {
        const uint32_t p = entry & ~PT_Q_DONE;
        const uint32_t r = ptShadeStep<VOL, true, TEX>(k, p, s_stack + threadIdx.x, extraShadow);
        if (r) ptAppend(k.pathOut, k.cntPathOut, r == 2u ? (p | PT_Q_DONE) : p);
      }
    }
  } else
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const uint32_t p = k.pathIn[i] & ~PT_Q_DONE;
      const uint32_t r = ptShadeStep<VOL, true, TEX>(k, p, s_stack + threadIdx.x, extraShadow);
      if (r) ptAppend(k.pathOut, k.cntPathOut, r == 2u ? (p | PT_Q_DONE) : p);
    }
  ptCountRays(k.a.fb, 0u, 0u, extraShadow);
}

// ---- TAIL -------------------------------------------------------------------------------------------------------------
// Launched ONCE, at the depth where Russian roulette has thinned the frame out (PathTracer.cpp:33 rouletteDepth): from there
// on a TRACE + SHADE launch pair per bounce costs two launch + drain latencies (~45 us) for a queue that holds less and
// less (C1: 14 more bounces for 3 % of the frame's rays, tools/gpu_launchlist.sh). Paths are independent, so here every lane
// carries ITS path to the end -- the shadow rays of the previous vertex, the continuation ray, the shade step, again -- and
// fetches the next path of the queue when its own is finished. Latency-bound by construction (the longest path's bounces
// are a serial chain); what it saves is the per-bounce launch pair.
template <bool VOL, bool TEX>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, 3) ptTailKernel(const __grid_constant__ PtArgs k)
{
  __shared__ uint32_t s_stack[OC_STACK_SHARED * PT_TRACE_BLOCK];
  const uint32_t n = *k.cntPathIn;
  if (n == 0 || frameCancelled(k.a.fb)) return;
  if (k.a.fb.control && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(k.a.fb.control + 1, k.progressUnit);
  const PtScratch &s = k.s;
  const OcWorld *world = k.a.world;
  uint32_t *stackBase = s_stack + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t nShadow = 0, nCont = 0, extraShadow = 0;
  bool alive = false;
  uint32_t p = 0, state = 0; // state of the path held: 1 goes on, 2 ended with shadow rays pending
  while (true) {
    __syncwarp();
    const uint32_t need = __ballot_sync(0xffffffffu, !alive);
    if (need) {
      uint32_t base = 0;
      if (lane == 0) base = frameCancelled(k.a.fb) ? 0xffffffffu : atomicAdd(k.fetchCursor, (uint32_t)__popc(need));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (!alive) {
        const uint32_t i = base + __popc(need & ((1u << lane) - 1u));
        if (base < n && i < n) { // base = ~0: the frame was cancelled
          const uint32_t e = k.pathIn[i];
          p = e & ~PT_Q_DONE; state = (e & PT_Q_DONE) ? 2u : 1u; alive = true;
        }
      }
      if (__ballot_sync(0xffffffffu, alive) == 0) break; // queue exhausted and every lane's path finished
    }
    if (alive) {
      if (s.flags[p] & PT_HAS_SHADOW)
        for (uint32_t j = 0; j < s.nls; j++) {
          const size_t e = (size_t)p * s.nls + j;
          const float4 c = s.shC[e];
          if (c.w == 0.f || !(isfinite(c.x) && isfinite(c.y) && isfinite(c.z))) continue;
          const float4 a = s.shA[e], b = s.shB[e];
          const Hit h = traceHitNI<1>(world, make_vec3f(a.x, a.y, a.z), make_vec3f(b.x, b.y, b.z), 0.f, a.w, stackBase, PT_TRACE_BLOCK);
          s.shHit[e] = make_uint2(h.rec, h.inst);
          nShadow++;
        }
      if (state == 1u) {
        const float4 o = s.rayO[p], d = s.rayD[p];
        Ray r;
        r.org = make_vec3f(o.x, o.y, o.z); r.dir = make_vec3f(d.x, d.y, d.z); r.t0 = d.w;
        const Hit h = traceHitNI<0>(world, r.org, r.dir, d.w, o.w, stackBase, PT_TRACE_BLOCK);
        resolveHit(world, h, r);
        s.hit0[p] = make_float4(r.t, r.u, r.v, __uint_as_float(r.primID));
        s.hit1[p] = make_float4(r.Ng.x, r.Ng.y, r.Ng.z, __uint_as_float(r.geomID));
        s.hitInst[p] = r.instID;
        nCont++;
      }
      state = ptShadeStep<VOL, false, TEX>(k, p, stackBase, extraShadow);
      alive = state != 0u;
    }
  }
  ptCountRays(k.a.fb, 0u, nCont, nShadow + extraShadow);
  if (k.a.fb.rayCounters) { // [3]: rays traced by this kernel (the others go through TRACE)
    uint32_t t = nCont + nShadow + extraShadow;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(k.a.fb.rayCounters + 3, (unsigned long long)t);
  }
}

// ===================================== ao renderer as a wavefront =======================================================
// AORenderer_renderSample (render/ao/AORenderer.ispc:23-149) + computeAO (render/util.ispc:12-54) over the same queues and
// the same TRACE kernel as the pathtracer, for worlds without volumes (those keep the megakernel, which marches them):
//   GEN        Renderer_default_renderTask's pixel sample (RendererRenderTaskFn.inl:30-60) -> camera ray
//   TRACE(d)   the ambient-occlusion rays of vertex d-1 (any hit) and the primary / continuation ray (closest hit)
//   SHADE(d)   fold the occlusion answers into vertex d-1's colour and blend it; then vertex d: postIntersect, surface
//              colour, the AO directions (rays emitted into the next queue), continuation behind a transparent surface
// Path state reuses the pathtracer's arrays: thr = outputColor, con = (surface colour, opacity) of the vertex waiting for
// its occlusion answers, lastP.x = its intensity before AO, lastN = (primary direction, original ray far), shA / shB =
// AO rays (org, radius | dir, epsilon), shHit = their answers (rays never traced because the direction grazes the surface
// are entered as occluded on the spot, util.ispc:31-35).
__device__ __forceinline__ vec4f aoBlendOver(vec4f out, vec4f c) // outputColor + (1 - outputColor.w) * blendedColor
{
  const float f = 1.f - out.w;
  return make_vec4f(out.x + f * c.x, out.y + f * c.y, out.z + f * c.z, __fmaf_rn(f, c.w, out.w));
}

__global__ void __launch_bounds__(128) aoGenKernel(const __grid_constant__ PtArgs k)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcFrameBuffer &fb = k.a.fb;
  uint32_t primary = 0;
  if (frameCancelled(fb)) return;
  if (fb.control && blockIdx.x == 0 && threadIdx.x == 0) fb.control[2] = k.progressTotal;
  const int sampleID = fb.frameID * ren.spp + 1 + k.pass; // startSampleID + s
  const float pixel_du = Halton_sample2(sampleID);
  const float pixel_dv = CranleyPattersonRotation(Halton_sample3(sampleID), 1.f / 6.f);
  const vec2f pfSample = pixelFilterSample(ren, make_vec2f(pixel_du, pixel_dv));
  const vec2f rcpSize = make_vec2f(1.f / (float)fb.sizeX, 1.f / (float)fb.sizeY);
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < k.numPaths; p += gridDim.x * blockDim.x) {
    uint32_t ix, iy;
    ptSlotToPixel(k, p, ix, iy);
    if (ix >= (uint32_t)fb.sizeX || iy >= (uint32_t)fb.sizeY) continue;
    if (fb.taskError && !(fb.taskError[(iy >> 3) * (((uint32_t)fb.sizeX + 7u) >> 3) + (ix >> 3)] > fb.errorThreshold)) continue;
    const vec2f pixel_center = make_vec2f(((float)ix + 0.5f) * rcpSize.x, ((float)iy + 0.5f) * rcpSize.y);
    Ray ray;
    cameraInitRay(k.a.cam, ray, pixel_center + pfSample * rcpSize, make_vec2f(0.f, 0.f), pixel_center);
    s.flags[p] = 0u;
    s.thr[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    s.lastN[p] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.t);
    if (k.pass == 0) { // the reference's `sample` lives across the spp loop and is never re-zeroed (RendererRenderTaskFn.inl:32-33)
      s.sZ[p] = make_float4(INFINITY, __uint_as_float(OC_INVALID_ID), __uint_as_float(OC_INVALID_ID), __uint_as_float(OC_INVALID_ID));
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fb.position) s.sPos[p] = zero;
      if (fb.firstNormal) s.sFN[p] = zero;
      if (fb.normal) s.sNrm[p] = zero;
      if (fb.albedo) s.sAlb[p] = zero;
    }
    s.rayO[p] = make_float4(ray.org.x, ray.org.y, ray.org.z, ray.t);
    s.rayD[p] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.t0);
    primary++;
    ptAppend(k.pathOut, k.cntPathOut, p);
  }
  ptCountRays(fb, primary, 0u, 0u);
}

// returns 0 finished, 1 goes on (continuation ray queued), 2 ended with occlusion rays pending
__device__ __forceinline__ uint32_t aoShadeStep(const PtArgs &k, uint32_t p, uint32_t ix, uint32_t iy)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcFrameBuffer &fb = k.a.fb;
  const OcWorld *world = k.a.world;
  const uint32_t flags = s.flags[p];
  const float4 out4 = s.thr[p];
  vec4f outputColor = make_vec4f(out4.x, out4.y, out4.z, out4.w);
  const uint32_t aoSamples = ren.aoSamples > 0 ? (uint32_t)ren.aoSamples : 0u;
  if (flags & PT_HAS_SHADOW) { // the occlusion answers of the previous vertex: computeAO's count, then the blend
    const float4 c = s.con[p];
    float intensity = s.lastP[p].x;
    int hits = 0;
    for (uint32_t j = 0; j < aoSamples; j++)
      if (s.shHit[(size_t)p * s.nls + j].y != OC_INVALID_ID) hits++;
    const float ao = 1.0f - (hits / (float)aoSamples);
    intensity *= (1.0f - ren.aoIntensity) + (ao * ren.aoIntensity);
    outputColor = aoBlendOver(outputColor, make_vec4f(make_vec3f(c.x, c.y, c.z) * intensity, c.w));
    if (flags & PT_DONE) {
      ptFinishPath<false>(k, p, make_vec3f(outputColor), outputColor.w);
      return 0u;
    }
  }
  const uint32_t depth = flags & PT_DEPTH_MASK;
  const bool firstHit = depth == 0;
  const float4 pd = s.lastN[p];
  const vec3f primaryDir = make_vec3f(pd.x, pd.y, pd.z);
  Ray ray;
  {
    const float4 o = s.rayO[p], d = s.rayD[p], h0 = s.hit0[p], h1 = s.hit1[p];
    ray.org = make_vec3f(o.x, o.y, o.z); ray.dir = make_vec3f(d.x, d.y, d.z); ray.t0 = d.w;
    ray.t = h0.x; ray.u = h0.y; ray.v = h0.z; ray.primID = __float_as_uint(h0.w);
    ray.Ng = make_vec3f(h1.x, h1.y, h1.z); ray.geomID = __float_as_uint(h1.w);
    ray.instID = s.hitInst[p];
  }
  if (!hadHit(ray)) {
    const vec4f bg = getBackground(ren);
    if (firstHit) {
      float4 z = s.sZ[p];
      z.x = ray.t;
      s.sZ[p] = z;
      if (fb.albedo) s.sAlb[p] = make_float4(bg.x, bg.y, bg.z, 0.f);
      if (fb.normal) s.sNrm[p] = make_float4(primaryDir.x, primaryDir.y, primaryDir.z, 0.f);
    }
    outputColor = aoBlendOver(outputColor, bg);
    ptFinishPath<false>(k, p, make_vec3f(outputColor), outputColor.w);
    return 0u;
  }
  DifferentialGeometry dg;
  postIntersect(world, ren, dg, ray, DG_NG | DG_NS | DG_NORMALIZE | DG_FACEFORWARD | DG_COLOR | DG_TEXCOORD | DG_TANGENTS);
  // getSurfaceColor (ao/surfaces.ispc:18-42)
  vec3f surfaceColor = make_vec3f(1.f);
  float opacity = 1.f;
  if (dg.material && dg.material->type == OC_MATERIAL_OBJ) {
    surfaceColor = toV(dg.material->Kd);
    opacity *= dg.material->d;
  }
  const vec3f color = surfaceColor * make_vec3f(dg.color);
  opacity = opacity * dg.color.w;
  const float intensity = fabsf(dot(dg.Ns, primaryDir)) * opacity;
  if (firstHit) {
    float z = ray.t;
    if (fb.projectedDepth) z *= dot(ray.dir, toV(k.a.cam.direction));
    s.sZ[p] = make_float4(z, __uint_as_float(dg.primID), __uint_as_float(dg.objID), __uint_as_float(dg.instID));
    if (fb.albedo) s.sAlb[p] = make_float4(color.x, color.y, color.z, 0.f);
    if (fb.position) s.sPos[p] = make_float4(dg.P.x, dg.P.y, dg.P.z, 0.f);
    if (fb.firstNormal) s.sFN[p] = make_float4(dg.Ns.x, dg.Ns.y, dg.Ns.z, 0.f);
    if (fb.normal) s.sNrm[p] = make_float4(dg.Ns.x, dg.Ns.y, dg.Ns.z, 0.f);
  }
  const bool needAO = aoSamples > 0 && ren.aoIntensity > 0.01f;
  // does the sample go on behind this surface? the alpha after the blend does not depend on the occlusion answers
  const bool alive = aoBlendOver(outputColor, make_vec4f(0.f, 0.f, 0.f, opacity)).w < 0.99f;
  if (needAO) {
    // computeAO (render/util.ispc:12-54) up to the rays
    const int sampleIDz = fb.frameID * ren.spp + 1 + k.pass;
    const int accumID = sampleIDz * (int)aoSamples;
    uint32_t v0 = ix, v1 = iy;
    tea8(v0, v1);
    const vec2f rot = make_vec2f(to_float_unorm(v0), to_float_unorm(v1));
    const linear3f localToWorld = frame(dg.Ns);
    for (uint32_t j = 0; j < aoSamples; j++) {
      const size_t e = (size_t)p * s.nls + j;
      const vec2f halton = make_vec2f(PermHalton_sample5(accumID + (int)j), PermHalton_sample7(accumID + (int)j));
      const vec2f r = make_vec2f(CranleyPattersonRotation(halton.x, rot.x), CranleyPattersonRotation(halton.y, rot.y));
      const vec3f ao_dir = localToWorld * cosineSampleHemisphere(r);
      if (dot(ao_dir, dg.Ns) < 0.01f) { s.shHit[e] = make_uint2(0u, 0u); continue; } // counts as occluded, no ray
      s.shA[e] = make_float4(dg.P.x, dg.P.y, dg.P.z, ren.aoRadius);
      s.shB[e] = make_float4(ao_dir.x, ao_dir.y, ao_dir.z, dg.epsilon);
      ptAppend(k.shadowOut, k.cntShadowOut, PT_SHADOW_RAY | (uint32_t)e);
    }
    s.con[p] = make_float4(color.x, color.y, color.z, opacity);
    s.lastP[p] = make_float4(intensity, 0.f, 0.f, 0.f);
  } else
    outputColor = aoBlendOver(outputColor, make_vec4f(color * intensity, opacity));
  if (!alive && !needAO) {
    ptFinishPath<false>(k, p, make_vec3f(outputColor), outputColor.w);
    return 0u;
  }
  s.thr[p] = make_float4(outputColor.x, outputColor.y, outputColor.z, outputColor.w);
  if (alive) {
    s.rayO[p] = make_float4(ray.org.x, ray.org.y, ray.org.z, pd.w);
    s.rayD[p] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, ray.t + dg.epsilon);
  }
  s.flags[p] = ((depth + 1) & PT_DEPTH_MASK) | (needAO ? PT_HAS_SHADOW : 0u) | (alive ? 0u : PT_DONE);
  return alive ? 1u : 2u;
}

__global__ void __launch_bounds__(128, 5) aoShadeKernel(const __grid_constant__ PtArgs k)
{
  if (frameCancelled(k.a.fb)) return;
  if (k.a.fb.control && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(k.a.fb.control + 1, k.progressUnit);
  const uint32_t n = *k.cntPathIn;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = k.pathIn[i] & ~PT_Q_DONE;
    uint32_t ix, iy;
    ptSlotToPixel(k, p, ix, iy);
    const uint32_t r = aoShadeStep(k, p, ix, iy);
    if (r) ptAppend(k.pathOut, k.cntPathOut, r == 2u ? (p | PT_Q_DONE) : p);
  }
}

} // namespace oc
