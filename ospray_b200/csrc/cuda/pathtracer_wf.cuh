// Wavefront path tracer: the pathtracer renderer as TRACE / SHADE kernel pairs over compacted path queues.
//
// Same algorithm, same sampler call order and the same per-path arithmetic as the megakernel in render.cu
// (reference: render/pathtracer/PathTracer.ispc:75-249, PathSampler.ispc:109-415, NextEventEstimation.ispc:25-160,
// TransparentShadow.ispc:18-85, VirtualLight.ispc:38-107); only the schedule differs:
//
//   TRACE<first>(d=0)   one thread per pixel sample: sampler init, pixel filter, camera ray, closest hit
//   SHADE(d)            everything between two traces of samplePath(): background, virtual lights, postIntersect,
//                       BSDF, NEE light sampling (shadow rays are EMITTED, not traced), BSDF sampling, roulette
//   TRACE(d+1)          the shadow rays emitted at vertex d (transparent-shadow loop) and the continuation ray
//   ...                 until maxPathLength; finished paths accumulate into the framebuffer where they end
//
// Why: the megakernel ran with 11.6 of 32 lanes active per instruction and stalled 7.3 issue slots per
// instruction on instruction fetch (289 KB of SASS, profiles/r1_pt_megakernel_ncu.txt). Here every kernel is a
// persistent grid-stride loop over a queue of live paths, appended with warp-aggregated atomics, so warps stay
// full as paths die, and each kernel's code fits the instruction cache. Path state lives in HBM as float4
// SoA arrays (128-bit coalesced accesses, ~180 B read + written per path and bounce).
#pragma once

#include <cooperative_groups.h>

namespace oc {
namespace cg = cooperative_groups;

enum : uint32_t {
  PT_DEPTH_MASK = 0xffu, PT_SCAT_SHIFT = 8, PT_SCAT_MASK = 0xff00u,
  PT_STRAIGHT = 1u << 16, PT_SPECTRANS = 1u << 17, PT_FIRSTBOUNCE = 1u << 18, PT_DONE = 1u << 19, PT_HAS_SHADOW = 1u << 20,
  PT_VOLUME = 1u << 21 // the vertex TRACE found is a scattering event inside a volume (hit1.xyz = its albedo)
};

struct PtScratch
{
  float4 *rayO;   // org.xyz, t (far)
  float4 *rayD;   // dir.xyz, t0
  float4 *hit0;   // t, u, v, primID
  float4 *hit1;   // Ng.xyz, geomID
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
  float4 *accC, *accZ, *accPos, *accFN, *accNrm, *accAlb; // ScreenSample sums across the spp passes (spp > 1)
  uint4 *rnd;     // worlds with volumes: pathState.randomSampler (PCG32 state lo/hi, stream)
  float *shT0;    // worlds with volumes: t0 of shadow entry [slot * nls + i] after the transparent-shadow loop
  uint32_t *queue[2];
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
  const uint32_t *qIn;
  uint32_t *qOut;
  const uint32_t *cntIn;
  uint32_t *cntOut;
  uint32_t *fetchCursor;  // ptFusedKernel's work-fetch cursor into qIn (zeroed with the counters every pass)
  uint32_t vol;           // the world holds volumes (host copy of OcWorld.hasVolumes): free-path sampling + transmittance
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

// a path ended: clamp, fold into the pixel's ScreenSample (spp > 1) and accumulate into the framebuffer
__device__ __forceinline__ void ptFinishPath(const PtArgs &k, uint32_t p, vec3f contribution, float alpha)
{
  const PtScratch &s = k.s;
  const OcFrameBuffer &fb = k.a.fb;
  uint32_t x, y;
  ptSlotToPixel(k, p, x, y);
  ScreenSample sm;
  sm.rgb = contribution;
  sm.alpha = alpha;
  if (isnan(contribution.x) || isnan(contribution.y) || isnan(contribution.z)) { sm.rgb = make_vec3f(0.f); sm.alpha = 1.0f; }
  sm.rgb = vmin(sm.rgb, make_vec3f(k.a.ren.maxRadiance));
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

// Does the surface hit by the any-hit query `ray` certainly take the shadow-ray contribution to exactly 0, as the
// reference's closest-hit loop would (TransparentShadow.ispc:18-85: lightContrib *= transparency per hit, result 0 as
// soon as reduce_max(lightContrib) <= minContribution)?  OBJ_getTransparency (OBJ.ispc:93-111) is T = Tf*d + (1-d) with
// d = material.d * dg.color.w.
//  * no colour alpha, d == 1, Tf == 0: T == 0 exactly.
//  * vec4 vertex colours whose alphas are ALL 1.0f (cornell_box): the interpolated alpha is 1 up to a few ulp
//    (barycentric / bilinear weights sum to 1 +- 4 ulp), so |T| <= 1e-6 and the product is <= maxContrib * 1e-6, which the
//    reference zeroes whenever that is <= minContribution. Checked here with the actual contribution.
// Anything else (per-primitive model colours, real alpha, thin glass, ...) answers "no" and takes the reference loop.
__device__ __forceinline__ bool ptOccluderIsOpaque(const OcWorld *world, const OcRenderer &ren, const Ray &ray, float maxContrib)
{
  const OcInstance *inst = world->instances + ray.instID;
  const OcModel *m = inst->group->models + ray.geomID;
  if (m->color.addr) return false;
  const uint32_t gf = m->geom->flags;
  if (gf & OC_GEOMFLAG_HAS_ALPHA) {
    if (!(gf & OC_GEOMFLAG_ALPHA_ALL_ONE) || !(maxContrib * 1e-6f <= ren.minContribution)) return false;
  }
  const OcMaterial *mat = getMaterial(world, ren, m, ray.primID);
  if (mat && mat->type == OC_MATERIAL_THINGLASS) return false;
  if (!mat || mat->type != OC_MATERIAL_OBJ) return true; // no / foreign material: transparency 0 (shading.cuh)
  return mat->d == 1.f && mat->Tf.x == 0.f && mat->Tf.y == 0.f && mat->Tf.z == 0.f;
}

struct PtCounts { uint32_t primary, secondary, shadow; };

// The reference's closest-hit transparent-shadow loop (TransparentShadow.ispc:18-85), taken only when the any-hit
// query of ptTraceStep found a non-opaque occluder (or the contribution is not finite). Kept out of line: inlined,
// postIntersect + the material transparencies (OBJ, thin glass) doubled the spill traffic of the TRACE kernels for a
// path that opaque scenes never take.
// Returns the attenuated contribution in xyz and the number of additional shadow segments traced in w (by value: a
// by-reference counter would pin the caller's ray counters into local memory).
static __device__ __noinline__ float4 ptTransparentShadowSlow(const PtArgs &k, size_t e, float cx, float cy, float cz, uint32_t *stackBase)
{
  uint32_t shadowCount = 0;
  Stack stack;
  stack.base = stackBase; stack.stride = PT_TRACE_BLOCK;
  vec3f lightContrib = make_vec3f(cx, cy, cz);
  const float4 a = k.s.shA[e], b = k.s.shB[e];
  Ray sr;
  setRay(sr, make_vec3f(a.x, a.y, a.z), make_vec3f(b.x, b.y, b.z), 0.f, a.w);
  const OcRenderer &ren = k.a.ren;
  uint32_t maxDepth = ren.maxDepth;
  const float tOriginal = sr.t;
  bool counted = isfinite(lightContrib.x) && isfinite(lightContrib.y) && isfinite(lightContrib.z); // the any-hit query was this segment
  while (1) {
    if (!counted) shadowCount++;
    counted = false;
    traceWorld<false>(k.a.world, sr, stack);
    if (!hadHit(sr)) break;
    DifferentialGeometry dg;
    postIntersect(k.a.world, ren, dg, sr, DG_NS | DG_NG | DG_FACEFORWARD | DG_NORMALIZE | DG_TEXCOORD | DG_COLOR);
    lightContrib = lightContrib * materialTransparency(dg, sr.dir);
    if (reduce_max(lightContrib) <= ren.minContribution || --maxDepth <= 0) { lightContrib = make_vec3f(0.f); break; }
    setRay(sr, sr.org, sr.dir, sr.t + dg.epsilon, tOriginal);
  }
  if (k.vol) k.s.shT0[e] = sr.t0; // where the ray stands after the transparent surfaces it crossed (volumeTransmittance starts there)
  return make_float4(lightContrib.x, lightContrib.y, lightContrib.z, __uint_as_float(shadowCount));
}

// TRACE step of one path: resolve the shadow rays emitted at the previous vertex, then (unless the path is
// done) trace the continuation ray. Returns false when the path finished here.
// VOL: the world holds volumes. A template parameter, not a run-time flag: with the volume calls compiled into the one
// instantiation every world used, SHADE ran 12 % slower on cornell_box (register allocation around the call sites).
template <bool FIRST, bool VOL>
__device__ __forceinline__ bool ptTraceStep(const PtArgs &k, uint32_t p, Stack stack, PtCounts &cnt)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcFrameBuffer &fb = k.a.fb;
  Ray ray;
  if (FIRST) {
    uint32_t ix, iy;
    ptSlotToPixel(k, p, ix, iy);
    if (ix >= (uint32_t)fb.sizeX || iy >= (uint32_t)fb.sizeY) return false;
    // adaptive accumulation: pixels of a converged render task start no path (LocalFB.cpp:245-257)
    if (fb.taskError && !(fb.taskError[(iy >> 3) * (((uint32_t)fb.sizeX + 7u) >> 3) + (ix >> 3)] > fb.errorThreshold)) return false;
    // PathTracer_renderPixel, one sample (PathTracer.ispc:149-193)
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
    cameraInitRay(k.a.cam, ray, screen, lens);
    // samplePath() prologue (PathSampler.ispc:116-133)
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
    cnt.primary++;
  } else {
    const uint32_t flags = s.flags[p];
    if (flags & PT_HAS_SHADOW) {
      // shadow rays of the previous vertex: transparentShadow (TransparentShadow.ispc:18-85) per light sample,
      // summed in sample order like nextEventEstimation's L (NextEventEstimation.ispc:140-156)
      vec3f L = make_vec3f(0.f);
      for (uint32_t j = 0; j < s.nls; j++) {
        const size_t e = (size_t)p * s.nls + j;
        const float4 c = s.shC[e];
        if (c.w == 0.f) continue;
        const float4 a = s.shA[e], b = s.shB[e];
        Ray sr;
        setRay(sr, make_vec3f(a.x, a.y, a.z), make_vec3f(b.x, b.y, b.z), 0.f, a.w);
        vec3f lightContrib = make_vec3f(c.x, c.y, c.z);
        // fast path: ANY opaque occluder makes the product of transparencies exactly 0, whatever lies in front
        // of it -> one any-hit query; only a non-opaque first find falls back to the reference's closest-hit loop
        bool resolved = false;
        float shadowT0 = 0.f;
        if (isfinite(lightContrib.x) && isfinite(lightContrib.y) && isfinite(lightContrib.z)) {
          Ray occ = sr;
          cnt.shadow++;
          if (!traceWorld<true>(k.a.world, occ, stack)) resolved = true; // unoccluded: lightContrib unchanged
          else if (ptOccluderIsOpaque(k.a.world, ren, occ, reduce_max(lightContrib))) { lightContrib = make_vec3f(0.f); resolved = true; }
        }
        if (!resolved) {
          const float4 r = ptTransparentShadowSlow(k, e, lightContrib.x, lightContrib.y, lightContrib.z, stack.base);
          lightContrib = make_vec3f(r.x, r.y, r.z);
          cnt.shadow += __float_as_uint(r.w);
          if (VOL) shadowT0 = s.shT0[e];
        }
        if (reduce_max(lightContrib) > 0) {
          // worlds with volumes: b.w = ls.pdf, c.w = fe.pdf (SHADE); the weight carries the transmittance estimates
          const float w = VOL ? ptVolumeShadowWeight(k, p, make_vec3f(a.x, a.y, a.z), make_vec3f(b.x, b.y, b.z), shadowT0, a.w, b.w, c.w == FLT_MIN ? 0.f : c.w)
                                : b.w;
          if (VOL) L = L + w * lightContrib; else L = L + lightContrib * w;
        }
      }
      float4 con = s.con[p];
      con.x += L.x; con.y += L.y; con.z += L.z;
      if (flags & PT_DONE) {
        ptFinishPath(k, p, make_vec3f(con.x, con.y, con.z), s.thr[p].w);
        return false;
      }
      s.con[p] = con;
    }
    const float4 o = s.rayO[p], d = s.rayD[p];
    setRay(ray, make_vec3f(o.x, o.y, o.z), make_vec3f(d.x, d.y, d.z), d.w, o.w);
    cnt.secondary++;
  }
  ray.u = ray.v = 0.f;
  ray.Ng = make_vec3f(0.f);
  const float rayLower = ray.t0, rayUpper = ray.t; // traceClippingRay without clipping geometry (World.ih:150-163)
  traceWorld<false>(k.a.world, ray, stack);
  if (VOL) { // volumeSampleFreePath (PathSampler.ispc:178-194): a scattering event before the surface replaces the hit
    const PtFreePath fp = ptVolumeFreePath(k, p, ray.org, ray.dir, ray.t0, ray.t, ray.geomID, rayLower, rayUpper);
    if (fp.albedoDist.w < INFINITY) {
      s.hit0[p] = make_float4(fp.albedoDist.w, 0.f, 0.f, __uint_as_float(0u));
      s.hit1[p] = make_float4(fp.albedoDist.x, fp.albedoDist.y, fp.albedoDist.z, __uint_as_float(fp.geomID));
      s.hitInst[p] = fp.instID;
      s.flags[p] |= PT_VOLUME;
      return true;
    }
  }
  s.hit0[p] = make_float4(ray.t, ray.u, ray.v, __uint_as_float(ray.primID));
  s.hit1[p] = make_float4(ray.Ng.x, ray.Ng.y, ray.Ng.z, __uint_as_float(ray.geomID));
  s.hitInst[p] = ray.instID;
  return true;
}

template <bool VOL>
__device__ bool ptShadeStep(const PtArgs &k, uint32_t p);

// evaluateVirtualLights in a world with volumes (VirtualLight.ispc:76-101): the MIS weight of a light the path runs into
// carries a transmittance estimate along the segment since the last vertex
static __device__ __noinline__ vec3f ptVirtualLightsVol(const PtArgs &k, uint32_t p, bool straightPath, vec3f throughput, vec3f lastP,
    vec3f lastNs, float lastPdf, uint32_t lastNumLightSamples, bool environment, vec3f rorg, vec3f rdir, float rt)
{
  const OcWorld *world = k.a.world;
  vec3f L = make_vec3f(0.f);
  const uint32_t numLights = world->numLights;
  if (numLights == 0) return L;
  const float lower = length(lastP - rorg);
  const float upper = environment ? INFINITY : length(lastP - (rorg + rt * rdir));
  const float selectionPDF = lastNumLightSamples / (float)numLights;
  RandomSampler rs = ptLoadRs(k.s, p);
  for (uint32_t i = 0; i < numLights; i++) {
    const OcLight &l = world->lights[i];
    if (!straightPath || l.isVisible) {
      const LightEval le = Light_eval(l, lastP, lastNs, rdir, lower, upper);
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

// ---- TRACE --------------------------------------------------------------------------------------------------
// Persistent grid-stride loop over the live-path queue. When few paths are left (RR has culled the frame down to
// a long thin tail), every further TRACE/SHADE launch would cost a launch + drain latency for almost no work
// (profiles/r1_pt_wavefront_launches.txt: 30 launches x ~35 us): instead each thread then carries its path to
// the end right here (TAIL mode); the queue it leaves behind is empty and the remaining launches exit at once.
template <bool FIRST, bool VOL>
__global__ void __launch_bounds__(PT_TRACE_BLOCK, 5) ptTraceKernel(const __grid_constant__ PtArgs k)
{
  __shared__ uint32_t s_stack[OC_STACK_SIZE * PT_TRACE_BLOCK];
  Stack stack;
  stack.base = s_stack + threadIdx.x;
  stack.stride = PT_TRACE_BLOCK;
  PtCounts cnt = {0, 0, 0};
  const uint32_t n = FIRST ? k.numPaths : *k.cntIn;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = FIRST ? i : k.qIn[i];
    if (ptTraceStep<FIRST, VOL>(k, p, stack, cnt)) ptAppend(k.qOut, k.cntOut, p);
  }
  ptCountRays(k.a.fb, cnt.primary, cnt.secondary, cnt.shadow);
}

// FUSED, launched ONCE at the depth where Russian roulette starts to cull (PathTracer.cpp:33 rouletteDepth): every
// remaining path is carried to its end here. Paths are independent, so instead of one launch + queue compaction per
// bounce (15 launches for ~20 % of the frame's rays, each paying a launch + drain latency for an ever thinner queue:
// profiles/r1_launches_bench_c1_wavefront.csv), each lane runs TRACE and SHADE back to back for its path and, when
// the path dies, immediately fetches the next live path from the input queue (warp-aggregated atomic on a fetch
// cursor). All lanes of a warp are at the same phase (trace, then shade) of *some* path until the queue runs dry.
template <bool VOL>
__global__ void __launch_bounds__(PT_TRACE_BLOCK) ptFusedKernel(const __grid_constant__ PtArgs k)
{
  __shared__ uint32_t s_stack[OC_STACK_SIZE * PT_TRACE_BLOCK];
  const uint32_t n = *k.cntIn;
  if (n == 0) return;
  Stack stack;
  stack.base = s_stack + threadIdx.x;
  stack.stride = PT_TRACE_BLOCK;
  PtCounts cnt = {0, 0, 0};
  const uint32_t lane = threadIdx.x & 31u;
  bool alive = false;
  uint32_t p = 0;
  while (true) {
    __syncwarp();
    const uint32_t need = __ballot_sync(0xffffffffu, !alive);
    if (need) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(k.fetchCursor, (uint32_t)__popc(need));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (!alive) {
        const uint32_t i = base + __popc(need & ((1u << lane) - 1u));
        if (i < n) { p = k.qIn[i]; alive = true; }
      }
      if (__ballot_sync(0xffffffffu, alive) == 0) break; // queue exhausted and every lane's path finished
    }
    if (alive) alive = ptTraceStep<false, VOL>(k, p, stack, cnt) && ptShadeStep<VOL>(k, p);
  }
  ptCountRays(k.a.fb, cnt.primary, cnt.secondary, cnt.shadow);
  if (k.a.fb.rayCounters) { // [3]: rays traced by this kernel
    uint32_t t = cnt.primary + cnt.secondary + cnt.shadow;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t) atomicAdd(k.a.fb.rayCounters + 3, (unsigned long long)t);
  }
}

// SHADE step of one path; returns true when the path goes on to the next TRACE step (alive, or done with
// shadow rays pending), false when it finished here.
template <bool VOL>
__device__ bool ptShadeStep(const PtArgs &k, uint32_t p)
{
  const PtScratch &s = k.s;
  const OcRenderer &ren = k.a.ren;
  const OcFrameBuffer &fb = k.a.fb;
  Stack dummy;
  dummy.base = nullptr; dummy.stride = 0;
  Ctx c(k.a.world, k.a.ren, k.a.cam, k.a.fb, dummy);
  {
    const bool volume = VOL && (s.flags[p] & PT_VOLUME) != 0; // scattering event inside a volume (ptVolumeFreePath)
    uint32_t flags = s.flags[p] & ~(PT_HAS_SHADOW | PT_DONE | PT_VOLUME);
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
    }
    float4 thr4 = s.thr[p], con4 = s.con[p];
    vec3f throughput = make_vec3f(thr4.x, thr4.y, thr4.z), contribution = make_vec3f(con4.x, con4.y, con4.z);
    float alpha = thr4.w;
    PathVertexLite last;
    {
      const float4 lp = s.lastP[p], ln = s.lastN[p];
      last.P = make_vec3f(lp.x, lp.y, lp.z); last.Ns = make_vec3f(ln.x, ln.y, ln.z);
      last.pdf_w = con4.w; last.numLightSamples = __float_as_uint(lp.w);
    }
    LDSampler ld;
    {
      uint32_t ix, iy, seed, so, sol, bits, bitsLight;
      ptSlotToPixel(k, p, ix, iy);
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
      const vec3f volumeAlbedo = ray.Ng; // hit1.xyz of a volume vertex
      const OcVolume *vmodel = nullptr;
      if (volume) {
        throughput = throughput * volumeAlbedo; // PathSampler.ispc:189-192
        vmodel = k.a.world->instances[ray.instID].group->volumes + ray.geomID;
      }
      const vec3f wo = neg(ray.dir);
      if (depth == 0) {
        float4 z = s.sZ[p];
        z.x = fb.projectedDepth ? ray.t * dot(ray.dir, toV(k.a.cam.direction)) : ray.t;
        s.sZ[p] = z;
      }
      if (environment && (ren.backgroundRefraction ? specularTransmissionPath : straightPath)) {
        const vec4f bg = getBackground(ren);
        alpha -= luminance(throughput) * (1.0f - bg.w);
        contribution = contribution + throughput * make_vec3f(bg);
      }
      if (!volume) { // PathSampler.ispc:220-228
        if (VOL)
          contribution = contribution + ptVirtualLightsVol(k, p, straightPath, throughput, last.P, last.Ns, last.pdf_w, last.numLightSamples,
                                            environment, ray.org, ray.dir, ray.t);
        else
          contribution = contribution + evaluateVirtualLights(c, straightPath, throughput, last, environment, ray);
      }
      if (environment) break;
      if (depth >= ren.maxDepth) break; // numGeoLights == 0

      DifferentialGeometry dg;
      if (volume) { // postIntersect of a volume vertex (PathSampler.ispc:51-67) + VolumetricModel_postIntersect (VolumetricModel.ih:16-33)
        dg.P = ray.org + ray.t * ray.dir;
        dg.Ng = dg.Ns = dg.dPds = make_vec3f(0.f);
        dg.color = make_vec4f(0.f, 0.f, 0.f, 0.f);
        dg.epsilon = 0.f; dg.material = nullptr; dg.primID = 0;
        dg.objID = vmodel->userID == OC_INVALID_ID ? ray.geomID : vmodel->userID;
        dg.instID = ray.instID;
        if (ray.instID != OC_INVALID_ID && k.a.world->instances[ray.instID].userID != OC_INVALID_ID)
          dg.instID = k.a.world->instances[ray.instID].userID;
      } else
        postIntersect(k.a.world, ren, dg, ray,
            DG_NS | DG_NG | DG_FACEFORWARD | DG_NORMALIZE | DG_TEXCOORD | DG_COLOR | DG_TANGENTS | DG_MOTIONBLUR);
      if (depth == 0) {
        float4 z = s.sZ[p];
        z.y = __uint_as_float(dg.primID); z.z = __uint_as_float(dg.objID); z.w = __uint_as_float(dg.instID);
        s.sZ[p] = z;
        if (fb.position) s.sPos[p] = make_float4(dg.P.x, dg.P.y, dg.P.z, 0.f);
        if (fb.firstNormal) s.sFN[p] = make_float4(dg.Ns.x, dg.Ns.y, dg.Ns.z, 0.f);
      }
      if (depth >= ren.maxDepth || scatteringEvents >= (uint32_t)ren.maxScatteringEvents) break;

      ObjBSDF bsdf;
      if (!volume) {
        OBJ_getBSDF(bsdf, dg);
        if (!bsdf.valid) break;
      }

      LDSampler_nextGroup(ld);
      uint32_t vertexLightSamples = 0;
      if (volume ? hgIsSmooth(vmodel->anisotropy) : (bsdf.scatteringType & SC_SMOOTH) != 0) { // isSmooth (PathTracerUtil.ih:15-28)
        if (scatteringEvents == 0) { // updateAuxilliaryData (PathTracerUtil.ih:48-60)
          if (fb.albedo) { const vec3f v = throughput * (volume ? volumeAlbedo : bsdf.albedo); s.sAlb[p] = make_float4(v.x, v.y, v.z, 0.f); }
          if (fb.normal) { const vec3f v = throughput * (volume ? dg.Ns : bsdf.frame.vz); s.sNrm[p] = make_float4(v.x, v.y, v.z, 0.f); }
        }
        // nextEventEstimation up to the shadow ray (NextEventEstimation.ispc:36-118): emit, TRACE(d+1) resolves
        const uint32_t numLightSamples = firstBounceLight ? k.nFirst : k.nIndirect;
        vertexLightSamples = numLightSamples;
        const uint32_t numLights = k.a.world->numLights;
        for (uint32_t j = 0; j < s.nls; j++) {
          const size_t e = (size_t)p * s.nls + j;
          bool valid = false;
          if (j < numLightSamples) {
            float ss;
            const vec2f sl = LDSampler_get3LightSamples(ld, j, numLightSamples > 1, ss);
            const Sample1D sel = Distribution1D_sample((int)numLights, k.a.world->lightsCDF, 0, ss);
            const float lightSelectionProb = sel.prob * numLightSamples;
            LightSample ls = Light_sample(k.a.world->lights[sel.idx], dg, sl);
            ls.weight = ls.weight / lightSelectionProb;
            ls.pdf *= lightSelectionProb;
            if (!(reduce_max(ls.weight) <= 0.0f || ls.pdf <= 0.0f)) {
              const EvalRes fe = volume ? HenyeyGreenstein_eval(vmodel->anisotropy, wo, ls.dir) : OBJ_eval(bsdf, wo, ls.dir);
              if (!(reduce_max(fe.value) <= 0.0f)) {
                vec3f org = dg.P;
                if (!volume && dot(dg.Ng, ls.dir) < 0.f) org = org - (2.0f * dg.epsilon) * dg.Ng;
                const vec3f unshaded = throughput * fe.value * ls.weight;
                s.shA[e] = make_float4(org.x, org.y, org.z, ls.dist);
                // worlds with volumes: TRACE needs the two pdfs apart (the MIS weight carries a transmittance estimate)
                s.shB[e] = make_float4(ls.dir.x, ls.dir.y, ls.dir.z, VOL ? ls.pdf : misHeuristic(ls.pdf, fe.pdf));
                s.shC[e] = make_float4(unshaded.x, unshaded.y, unshaded.z, VOL ? (fe.pdf > 0.f ? fe.pdf : FLT_MIN) : 1.f);
                valid = true;
                hasShadow = true;
              }
            }
          }
          if (!valid) s.shC[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        firstBounceLight = false;
      }

      float ss;
      uint32_t uu3;
      const vec2f sb = LDSampler_getNext4Samples(ld, ss, uu3);
      SampleRes fs = volume ? HenyeyGreenstein_sample(vmodel->anisotropy, wo, sb) : OBJ_sample(bsdf, wo, sb, ss);
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
      vec3f ray_org = dg.P;
      if (!volume && (fs.type & SC_TRANSMISSION)) ray_org = ray_org - (2.0f * dg.epsilon) * dg.Ng;
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
      return false;
    }
    s.thr[p] = make_float4(throughput.x, throughput.y, throughput.z, alpha);
    s.con[p] = make_float4(contribution.x, contribution.y, contribution.z, last.pdf_w);
    if (alive) {
      s.lastP[p] = make_float4(last.P.x, last.P.y, last.P.z, __uint_as_float(last.numLightSamples));
      s.lastN[p] = make_float4(last.Ns.x, last.Ns.y, last.Ns.z, 0.f);
      s.smp[p] = make_uint2(ld.revIndex, ld.scramble);
    }
    s.flags[p] = ((depth + 1) & PT_DEPTH_MASK) | (scatteringEvents << PT_SCAT_SHIFT) | (straightPath ? PT_STRAIGHT : 0u)
        | (specularTransmissionPath ? PT_SPECTRANS : 0u) | (firstBounceLight ? PT_FIRSTBOUNCE : 0u) | (hasShadow ? PT_HAS_SHADOW : 0u)
        | (alive ? 0u : PT_DONE);
    return true;
  }
}

template <bool VOL>
__global__ void __launch_bounds__(128, 5) ptShadeKernel(const __grid_constant__ PtArgs k)
{
  const uint32_t n = *k.cntIn;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = k.qIn[i];
    if (ptShadeStep<VOL>(k, p)) ptAppend(k.qOut, k.cntOut, p);
  }
}

} // namespace oc
