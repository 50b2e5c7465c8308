// Ray / primitive intersection and two-level wide-BVH traversal.
//
// Stands in for Embree 4.3.3 rtcIntersectV / rtcOccludedV as called from
// modules/cpu/common/World.ih:29-52, :311-335 (un-vendored dependency; algorithm restated
// from Embree's published kernels: Moeller-Trumbore triangle test with edges e1=v0-v1,
// e2=v2-v0 and Ng=cross(e2,e1); quads as (v0,v1,v3)+(v2,v3,v1) with (1-u,1-v) remap, the same
// split the reference mirrors in geometry/Mesh.ispc:315-345; sphere test with projected
// centre; OSPRay's own box slab test common/Intersect.ih:194-221 + FilterIntersect.ih:41-116).
//
// Arithmetic in the primitive tests is spelled with explicit fused ops so the CPU oracle
// (oracle/oracle.cpp, fmaf) reproduces hit/miss decisions bit for bit.
#pragma once
#include "../../../include/ospcuda.h"
#include "../common/osp_math.h"

namespace oc {
using namespace ospm;

#define OC_INVALID_ID 0xFFFFFFFFu
#define OC_TLAS_THRESHOLD 4u
#define OC_STACK_SIZE OSPCUDA_TRAVERSAL_STACK // include/ospcuda.h: the host checks every built hierarchy against it
#define OC_LINEAR_MAX 8u // groups with at most this many leaf records are scanned linearly (no hierarchy)
#define OC_MAX_BLOCK 128 // largest CTA of a kernel that traverses (shared stack = OC_STACK_SIZE * blockDim.x words)

struct Ray
{
  vec3f org; float t0;
  vec3f dir; float t;
  vec3f Ng;  float u, v;
  uint32_t primID, geomID, instID;
};

__device__ __forceinline__ void setRay(Ray &r, vec3f org, vec3f dir, float t0, float t1)
{
  r.org = org; r.dir = dir; r.t0 = t0; r.t = t1;
  r.primID = r.geomID = r.instID = OC_INVALID_ID;
}
__device__ __forceinline__ bool hadHit(const Ray &r) { return r.geomID != OC_INVALID_ID; }

__device__ __forceinline__ vec3f toV(const OcVec3f &v) { return make_vec3f(v.x, v.y, v.z); }
__device__ __forceinline__ affine3f toA(const OcAffine3f &a)
{
  affine3f r; r.l.vx = toV(a.vx); r.l.vy = toV(a.vy); r.l.vz = toV(a.vz); r.p = toV(a.p); return r;
}

// fused dot / cross (Embree AVX2 madd/msub pattern)
__device__ __forceinline__ float dotF(vec3f a, vec3f b)
{
  return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, __fmul_rn(a.z, b.z)));
}
__device__ __forceinline__ vec3f crossF(vec3f a, vec3f b)
{
  return make_vec3f(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
      __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// ray -> instance space with every product and sum rounded separately (no fma contraction), so that the
// local ray -- and with it every hit decision inside the instance -- is bit-identical to the oracle's
__device__ __forceinline__ vec3f xfmVectorExact(const linear3f &l, vec3f v)
{
  return make_vec3f(__fadd_rn(__fadd_rn(__fmul_rn(v.x, l.vx.x), __fmul_rn(v.y, l.vy.x)), __fmul_rn(v.z, l.vz.x)),
      __fadd_rn(__fadd_rn(__fmul_rn(v.x, l.vx.y), __fmul_rn(v.y, l.vy.y)), __fmul_rn(v.z, l.vz.y)),
      __fadd_rn(__fadd_rn(__fmul_rn(v.x, l.vx.z), __fmul_rn(v.y, l.vy.z)), __fmul_rn(v.z, l.vz.z)));
}
__device__ __forceinline__ vec3f xfmPointExact(const affine3f &a, vec3f v)
{
  const vec3f r = xfmVectorExact(a.l, v);
  return make_vec3f(__fadd_rn(r.x, a.p.x), __fadd_rn(r.y, a.p.y), __fadd_rn(r.z, a.p.z));
}

// ---- triangle (v0, e1 = v0 - v1, e2 = v2 - v0) ---------------------------------------
__device__ __forceinline__ bool intersectTriangle(const Ray &ray, vec3f v0, vec3f e1, vec3f e2,
    float &t, float &u, float &v, vec3f &Ng)
{
  const vec3f triNg = crossF(e2, e1);
  const vec3f C = make_vec3f(__fsub_rn(v0.x, ray.org.x), __fsub_rn(v0.y, ray.org.y), __fsub_rn(v0.z, ray.org.z));
  const vec3f R = crossF(C, ray.dir);
  const float den = dotF(triNg, ray.dir);
  const float absDen = fabsf(den);
  const uint32_t sgn = __float_as_uint(den) & 0x80000000u;
  // staged rejection (identical decisions: V >= 0 and fl(U+V) <= absDen imply U <= absDen); with coherent
  // rays most warps leave at the first test and never issue the V / T arithmetic
  const float U = __uint_as_float(__float_as_uint(dotF(R, e2)) ^ sgn);
  if (!(den != 0.f && U >= 0.f && U <= absDen))
    return false;
  const float V = __uint_as_float(__float_as_uint(dotF(R, e1)) ^ sgn);
  if (!(V >= 0.f && __fadd_rn(U, V) <= absDen))
    return false;
  const float T = __uint_as_float(__float_as_uint(dotF(triNg, C)) ^ sgn);
  if (!(__fmul_rn(absDen, ray.t0) < T && T <= __fmul_rn(absDen, ray.t)))
    return false;
  const float rcpAbsDen = __frcp_rn(absDen);
  u = __fmul_rn(U, rcpAbsDen);
  v = __fmul_rn(V, rcpAbsDen);
  t = __fmul_rn(T, rcpAbsDen);
  Ng = triNg;
  return true;
}

// ---- sphere (centre c, radius r) ------------------------------------------------------
__device__ __forceinline__ bool intersectSphere(const Ray &ray, vec3f c, float r, float &t, vec3f &Ng)
{
  const float rd2 = __frcp_rn(dotF(ray.dir, ray.dir));
  const vec3f c0 = make_vec3f(__fsub_rn(c.x, ray.org.x), __fsub_rn(c.y, ray.org.y), __fsub_rn(c.z, ray.org.z));
  const float projC0 = __fmul_rn(dotF(c0, ray.dir), rd2);
  const vec3f perp = make_vec3f(__fmaf_rn(-projC0, ray.dir.x, c0.x), __fmaf_rn(-projC0, ray.dir.y, c0.y),
      __fmaf_rn(-projC0, ray.dir.z, c0.z));
  const float l2 = dotF(perp, perp);
  const float r2 = __fmul_rn(r, r);
  if (!(l2 <= r2))
    return false;
  const float td = __fsqrt_rn(__fmul_rn(__fsub_rn(r2, l2), rd2));
  const float tFront = __fsub_rn(projC0, td);
  const float tBack = __fadd_rn(projC0, td);
  float tdSel;
  if (ray.t0 <= tFront && tFront <= ray.t) { t = tFront; tdSel = -td; }
  else if (ray.t0 <= tBack && tBack <= ray.t) { t = tBack; tdSel = td; }
  else return false;
  Ng = make_vec3f(__fmaf_rn(tdSel, ray.dir.x, -perp.x), __fmaf_rn(tdSel, ray.dir.y, -perp.y),
      __fmaf_rn(tdSel, ray.dir.z, -perp.z));
  return true;
}

// ---- box (OSPRay user geometry; common/Intersect.ih:194-221) ---------------------------
__device__ __forceinline__ float rcp_safe(float f)
{
  // rkcommon rcp_safe: avoid division by ~0 (un-vendored; restated)
  return __frcp_rn(fabsf(f) < 1e-8f ? (f < 0.f ? -1e-8f : 1e-8f) : f);
}
__device__ __forceinline__ bool intersectBox(const Ray &ray, vec3f lo, vec3f hi, float &t, vec3f &Ng)
{
  const vec3f rd = make_vec3f(rcp_safe(ray.dir.x), rcp_safe(ray.dir.y), rcp_safe(ray.dir.z));
  const vec3f mins = make_vec3f(__fmul_rn(__fsub_rn(lo.x, ray.org.x), rd.x), __fmul_rn(__fsub_rn(lo.y, ray.org.y), rd.y),
      __fmul_rn(__fsub_rn(lo.z, ray.org.z), rd.z));
  const vec3f maxs = make_vec3f(__fmul_rn(__fsub_rn(hi.x, ray.org.x), rd.x), __fmul_rn(__fsub_rn(hi.y, ray.org.y), rd.y),
      __fmul_rn(__fsub_rn(hi.z, ray.org.z), rd.z));
  const vec3f nears = vmin(mins, maxs), fars = vmax(mins, maxs);
  const float tEntry = reduce_max(nears), tExit = reduce_min(fars);
  if (!(tEntry < tExit))
    return false;
  if (tEntry > ray.t0 && tEntry <= ray.t) {
    t = tEntry;
    if (tEntry == nears.x) Ng = make_vec3f(ray.dir.x > 0.f ? -1.f : 1.f, 0.f, 0.f);
    else if (tEntry == nears.y) Ng = make_vec3f(0.f, ray.dir.y > 0.f ? -1.f : 1.f, 0.f);
    else Ng = make_vec3f(0.f, 0.f, ray.dir.z > 0.f ? -1.f : 1.f);
    return true;
  }
  if (tExit > ray.t0 && tExit <= ray.t) {
    t = tExit;
    if (tExit == fars.x) Ng = make_vec3f(ray.dir.x > 0.f ? 1.f : -1.f, 0.f, 0.f);
    else if (tExit == fars.y) Ng = make_vec3f(0.f, ray.dir.y > 0.f ? 1.f : -1.f, 0.f);
    else Ng = make_vec3f(0.f, 0.f, ray.dir.z > 0.f ? 1.f : -1.f);
    return true;
  }
  return false;
}

// ---- capsule = one piece of a round curve -------------------------------------------------------------------------
// Stands in for Embree's round-curve sweep intersector (kernels/geometry/curve_intersector_sweep.h, un-vendored):
// every curve segment is tessellated at BVH build into OC_CURVE_SUBDIV capsules whose end points are the curve basis
// evaluated at k/8 and whose radius is the mean of the two end radii (bvh_build.cu). Front hit only; u = u0 + du * s,
// Ng = P - axis(s). Every operation is rounded separately so that the oracle (plain C, -ffp-contract=off) reproduces it.
__device__ __forceinline__ float dotRN(vec3f a, vec3f b)
{
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ vec3f subRN(vec3f a, vec3f b) { return make_vec3f(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ vec3f maddRN(vec3f a, float s, vec3f b) // a + s * b
{
  return make_vec3f(__fadd_rn(a.x, __fmul_rn(s, b.x)), __fadd_rn(a.y, __fmul_rn(s, b.y)), __fadd_rn(a.z, __fmul_rn(s, b.z)));
}
// The ray origin is first moved along the ray to the point closest to pa (parameter tp) so that the discriminant of a
// grazing ray seen from far away does not cancel catastrophically (see oracle/curve_capsule.h); t = tp + t'.
// Out of line, hit returned by value: inlined into intersectRecord, the capsule's early-out control flow made the compiler
// restructure the whole leaf loop and traversal ran with half the lanes per warp (C3: 95 -> 190 ms, tools/ab_commits.sh).
struct CapsuleHit { float t, s; vec3f Ng; bool hit; };
static __device__ __noinline__ CapsuleHit intersectCapsuleNI(vec3f org, vec3f d, float rayT0, float rayT, vec3f pa, vec3f pb, float r)
{
  CapsuleHit res;
  res.hit = false; res.t = 0.f; res.s = 0.f; res.Ng = make_vec3f(0.f);
  float &t = res.t, &sOut = res.s;
  vec3f &Ng = res.Ng;
  struct { float t0, t; } ray = {rayT0, rayT};
  const vec3f ba = subRN(pb, pa), oa0 = subRN(org, pa);
  const float dd = dotRN(d, d);
  const float tp = -__fdiv_rn(dotRN(d, oa0), dd);
  const vec3f oa = maddRN(oa0, tp, d);
  const float baba = dotRN(ba, ba), bard = dotRN(ba, d), baoa = dotRN(ba, oa), rdoa = dotRN(d, oa), oaoa = dotRN(oa, oa);
  const float rr = __fmul_rn(r, r);
  const float a = __fsub_rn(__fmul_rn(baba, dd), __fmul_rn(bard, bard));
  float y;
  if (a > 0.f) {
    const float b = __fsub_rn(__fmul_rn(baba, rdoa), __fmul_rn(baoa, bard));
    const float c = __fsub_rn(__fsub_rn(__fmul_rn(baba, oaoa), __fmul_rn(baoa, baoa)), __fmul_rn(rr, baba));
    const float h = __fsub_rn(__fmul_rn(b, b), __fmul_rn(a, c));
    if (!(h >= 0.f)) return res; // misses the infinite cylinder, hence the capsule
    const float tb = __fdiv_rn(__fsub_rn(-b, __fsqrt_rn(h)), a);
    y = __fadd_rn(baoa, __fmul_rn(tb, bard));
    if (y > 0.f && y < baba) {
      const float tt = __fadd_rn(tp, tb);
      if (!(tt > ray.t0 && tt <= ray.t)) return res;
      const float s = __fdiv_rn(y, baba);
      t = tt; sOut = s;
      Ng = maddRN(maddRN(oa, tb, d), -s, ba);
      res.hit = true;
      return res;
    }
  } else
    y = bard > 0.f ? -1.f : __fadd_rn(baba, 1.f); // ray parallel to the axis: it enters through the cap it faces
  const bool first = y <= 0.f;
  const vec3f oc = first ? oa : subRN(oa, ba);
  const float b2 = dotRN(d, oc);
  const float c2 = __fsub_rn(dotRN(oc, oc), rr);
  const float h2 = __fsub_rn(__fmul_rn(b2, b2), __fmul_rn(dd, c2));
  if (!(h2 > 0.f)) return res;
  const float tc = __fdiv_rn(__fsub_rn(-b2, __fsqrt_rn(h2)), dd);
  const float tt = __fadd_rn(tp, tc);
  if (!(tt > ray.t0 && tt <= ray.t)) return res;
  t = tt; sOut = first ? 0.f : 1.f;
  Ng = maddRN(oc, tc, d);
  res.hit = true;
  return res;
}

// ---- leaf records -----------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const OcVec4f *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Every non-triangle primitive test sits behind ONE call: the triangle test stays inline in the leaf loop, whose code then
// has the shape the compiler predicates well (C3 89.7 -> 87.1 ms, C4 4.67 -> 4.53 ms against inlined sphere / box tests).
struct OtherHit { float t, u; vec3f Ng; bool hit; };
static __device__ __noinline__ OtherHit intersectOtherNI(vec3f org, vec3f dir, float rayT0, float rayT, float4 r0, float4 r1, float4 r2)
{
  Ray ray;
  ray.org = org; ray.dir = dir; ray.t0 = rayT0; ray.t = rayT;
  OtherHit o;
  o.u = 0.f; o.t = 0.f; o.Ng = make_vec3f(0.f);
  const uint32_t kind = __float_as_uint(r2.w);
  if (kind == OC_PRIM_SPHERE) o.hit = intersectSphere(ray, make_vec3f(r0.x, r0.y, r0.z), r1.x, o.t, o.Ng);
  else if (kind == OC_PRIM_CAPSULE) {
    const CapsuleHit ch = intersectCapsuleNI(org, dir, rayT0, rayT, make_vec3f(r0.x, r0.y, r0.z), make_vec3f(r1.x, r1.y, r1.z), r2.x);
    o.hit = ch.hit; o.t = ch.t; o.Ng = ch.Ng;
    if (o.hit) o.u = __fadd_rn(r2.y, __fmul_rn(r2.z, ch.s));
  } else o.hit = intersectBox(ray, make_vec3f(r0.x, r0.y, r0.z), make_vec3f(r1.x, r1.y, r1.z), o.t, o.Ng);
  return o;
}

template <bool OCCLUSION>
__device__ __forceinline__ bool intersectRecord(Ray &ray, const OcVec4f *rec, uint32_t instID)
{
  const float4 r0 = ldg4(rec), r1 = ldg4(rec + 1), r2 = ldg4(rec + 2);
  const uint32_t kind = __float_as_uint(r2.w);
  float t, u = 0.f, v = 0.f;
  vec3f Ng;
  bool hit;
  if (kind <= OC_PRIM_TRI_QUAD2ND) {
    hit = intersectTriangle(ray, make_vec3f(r0.x, r0.y, r0.z), make_vec3f(r1.x, r1.y, r1.z),
        make_vec3f(r2.x, r2.y, r2.z), t, u, v, Ng);
    if (hit && kind == OC_PRIM_TRI_QUAD2ND) { u = __fsub_rn(1.f, u); v = __fsub_rn(1.f, v); }
  }
  else {
    const OtherHit o = intersectOtherNI(ray.org, ray.dir, ray.t0, ray.t, r0, r1, r2);
    hit = o.hit; t = o.t; u = o.u; Ng = o.Ng;
  }
  if (!hit)
    return false;
  const uint32_t primID = __float_as_uint(r0.w), geomID = __float_as_uint(r1.w);
  if (!OCCLUSION && t == ray.t && ray.geomID != OC_INVALID_ID) {
    // exact tie with the hit already held (coplanar faces, shared edges): the primitive tests accept t <= ray.t, so a
    // scan in record order ends with the LAST of the tied primitives (what the oracle's and the Embree stand-in's loops
    // do). Same rule here whatever order the BVH visits them in: the later (instance, geometry, primitive, quad half) wins.
    const bool later = instID != ray.instID ? instID > ray.instID
        : geomID != ray.geomID              ? geomID > ray.geomID
        : primID != ray.primID              ? primID > ray.primID
                                            : kind == OC_PRIM_TRI_QUAD2ND;
    if (!later) return false;
  }
  ray.t = t;
  ray.primID = primID;
  ray.geomID = geomID;
  ray.instID = instID;
  if (!OCCLUSION) { ray.u = u; ray.v = v; ray.Ng = Ng; }
  return true;
}

// ---- wide-BVH traversal -------------------------------------------------------------------
struct NodeHit { float t[4]; };

// slab test of one 128-byte node against the ray: 8 x LDG.128, conservative (far side padded
// by 2 ulp) so that the BVH never culls a primitive the exact test would accept.
__device__ __forceinline__ uint32_t nodeTest(const OcNode4 *node, vec3f org, vec3f rdir, float t0, float t1,
    float (&tnear)[4], uint4 &child)
{
  const float4 *n = reinterpret_cast<const float4 *>(node);
  const float4 lox = __ldg(n + 0), hix = __ldg(n + 1), loy = __ldg(n + 2), hiy = __ldg(n + 3),
               loz = __ldg(n + 4), hiz = __ldg(n + 5);
  child = __ldg(reinterpret_cast<const uint4 *>(n + 6));
  uint32_t mask = 0;
#define OC_SLAB(i, LX, HX, LY, HY, LZ, HZ, CH)                                                   \
  {                                                                                              \
    const float ax = (LX - org.x) * rdir.x, bx = (HX - org.x) * rdir.x;                          \
    const float ay = (LY - org.y) * rdir.y, by = (HY - org.y) * rdir.y;                          \
    const float az = (LZ - org.z) * rdir.z, bz = (HZ - org.z) * rdir.z;                          \
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t0));       \
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t1)) * 1.0000005f; \
    tnear[i] = tn;                                                                               \
    if (CH != OC_INVALID_ID && tn <= tf) mask |= 1u << i;                                        \
  }
  OC_SLAB(0, lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, child.x)
  OC_SLAB(1, lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, child.y)
  OC_SLAB(2, lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, child.z)
  OC_SLAB(3, lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, child.w)
#undef OC_SLAB
  return mask;
}

__device__ __forceinline__ vec3f safeRcpDir(vec3f d)
{
  // __frcp_rn is the correctly rounded reciprocal == 1.f / x, without the full division sequence
  const float eps = 1e-30f;
  return make_vec3f(__frcp_rn(fabsf(d.x) < eps ? copysignf(eps, d.x) : d.x), __frcp_rn(fabsf(d.y) < eps ? copysignf(eps, d.y) : d.y),
      __frcp_rn(fabsf(d.z) < eps ? copysignf(eps, d.z) : d.z));
}

// Per-thread traversal stack in SHARED memory, laid out [depth][thread] so that the 32 lanes of a warp hit 32
// different banks whatever their individual depths are (a local-memory stack puts lanes at different depths
// into different 32-byte sectors). Kernels declare `__shared__ uint32_t s[OC_STACK_SIZE * blockDim.x]`.
struct Stack
{
  uint32_t *base;   // s + threadIdx.x
  uint32_t stride;  // blockDim.x
  __device__ __forceinline__ uint32_t &at(int d) const { return base[(uint32_t)d * stride]; }
};

// Traverse one BVH (group BLAS), "while-while": descend inner nodes until a leaf is reached, then test its
// records. Inlined: a by-reference Ray handed to a non-inlined function lives in local memory (309 LDL / 198 STL
// in the first wavefront TRACE kernel). Kernels with many trace call sites (the scivis / ao megakernels) go
// through the non-inlined wrapper traceWorldNI below instead, which bounds their code size.
template <bool OCCLUSION>
__device__ __forceinline__ bool traverseBLAS(const OcNode4 *__restrict__ nodes, const OcVec4f *__restrict__ prims, uint32_t numRecords,
    Ray &ray, uint32_t instID, Stack st, int sp0)
{
  bool any = false;
  if (nodes == nullptr) { // linear group (<= OC_LINEAR_MAX records): uniform loop, warp-broadcast loads
    const uint32_t n = numRecords;
    for (uint32_t i = 0; i < n; i++)
      if (intersectRecord<OCCLUSION>(ray, prims + 3 * (size_t)i, instID)) {
        any = true;
        if (OCCLUSION) return true;
      }
    return any;
  }
  const vec3f rdir = safeRcpDir(ray.dir);
  int sp = sp0;
  uint32_t cur = 0; // root
  while (true) {
    // ---- inner nodes ----
    while (!(cur & 0x80000000u)) {
      float tn[4];
      uint4 ch;
      const uint32_t mask = nodeTest(nodes + cur, ray.org, rdir, ray.t0, ray.t, tn, ch);
      if (mask == 0) {
        if (sp == sp0) return any;
        cur = st.at(--sp);
        continue;
      }
#ifndef OC_UNSORTED_PUSH // sorted: C3 106 -> 94 ms, C5s 51 -> 45 ms (tools/ab_all.sh)
      if (!OCCLUSION) {
        // closest hit: order the hit children by entry distance (5 compare-exchanges); the nearest is visited next, the
        // others are pushed far-to-near so that pops come back nearest-first and ray.t culls the rest early
        float k0 = (mask & 1u) ? tn[0] : INFINITY, k1 = (mask & 2u) ? tn[1] : INFINITY, k2 = (mask & 4u) ? tn[2] : INFINITY,
              k3 = (mask & 8u) ? tn[3] : INFINITY;
        uint32_t c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
#define OC_CSWAP(ka, ca, kb, cb) { const bool sw = kb < ka; const float tk = sw ? kb : ka; kb = sw ? ka : kb; ka = tk; const uint32_t tc = sw ? cb : ca; cb = sw ? ca : cb; ca = tc; }
        OC_CSWAP(k0, c0, k1, c1) OC_CSWAP(k2, c2, k3, c3) OC_CSWAP(k0, c0, k2, c2) OC_CSWAP(k1, c1, k3, c3) OC_CSWAP(k1, c1, k2, c2)
#undef OC_CSWAP
        if (k3 < INFINITY && sp < OC_STACK_SIZE) st.at(sp++) = c3;
        if (k2 < INFINITY && sp < OC_STACK_SIZE) st.at(sp++) = c2;
        if (k1 < INFINITY && sp < OC_STACK_SIZE) st.at(sp++) = c1;
#ifdef OC_PREFETCH_NEXT
        // the entry just pushed last is what the next pop returns: start fetching its node / first record now
        if (k1 < INFINITY) {
          const void *nx = (c1 & 0x80000000u) ? (const void *)(prims + 3 * (size_t)(c1 & 0x07FFFFFFu)) : (const void *)(nodes + c1);
          #ifdef OC_PREFETCH_L1
          asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
#else
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
#endif
        }
#endif
        cur = c0;
        continue;
      }
#endif
      const uint32_t chs[4] = {ch.x, ch.y, ch.z, ch.w};
      int best = 0;
      float bestT = INFINITY;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (((mask >> i) & 1u) && tn[i] < bestT) { bestT = tn[i]; best = i; }
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (((mask >> i) & 1u) && i != best && sp < OC_STACK_SIZE) st.at(sp++) = chs[i];
      cur = chs[best];
    }
    // ---- leaf ----
    const uint32_t first = cur & 0x07FFFFFFu;
    const uint32_t count = (cur >> 27) & 0xFu;
    for (uint32_t i = 0; i < count; i++) {
      if (intersectRecord<OCCLUSION>(ray, prims + 3 * (size_t)(first + i), instID)) {
        any = true;
        if (OCCLUSION) return true;
      }
    }
    if (sp == sp0) return any;
    cur = st.at(--sp);
  }
}

__device__ __forceinline__ bool rayBoxOverlap(const float *b, vec3f org, vec3f rdir, float t0, float t1)
{
  const float ax = (b[0] - org.x) * rdir.x, bx = (b[3] - org.x) * rdir.x;
  const float ay = (b[1] - org.y) * rdir.y, by = (b[4] - org.y) * rdir.y;
  const float az = (b[2] - org.z) * rdir.z, bz = (b[5] - org.z) * rdir.z;
  const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t0));
  const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t1)) * 1.0000005f;
  return tn <= tf;
}

// one instance: ray -> instance space (Embree keeps t, no renormalisation), one BLAS traversal, hit back
template <bool OCCLUSION>
__device__ __forceinline__ bool traceInstance(const OcInstance *__restrict__ inst, uint32_t instID, Ray &ray, Stack st, int sp0)
{
  const OcGroup *g = inst->group;
  const uint32_t numRecords = g->bvh.numRecords;
  if (numRecords == 0)
    return false;
  Ray local = ray;
  if (!inst->isIdentity) {
    const affine3f rx = toA(inst->rcp_xfm);
    local.org = xfmPointExact(rx, ray.org);
    local.dir = xfmVectorExact(rx.l, ray.dir);
  }
  const bool h = traverseBLAS<OCCLUSION>(g->bvh.nodes, g->bvh.prims, numRecords, local, instID, st, sp0);
  if (h) {
    ray.t = local.t;
    ray.primID = local.primID; ray.geomID = local.geomID; ray.instID = local.instID;
    if (!OCCLUSION) { ray.u = local.u; ray.v = local.v; ray.Ng = local.Ng; } // Ng stays in instance space (World.ih:411-420)
  }
  return h;
}

// closest hit / occlusion over the whole world (two-level). Equivalent of traceGeometryRay /
// isOccludedNoClipping (World.ih:29-52, 311-335).
template <bool OCCLUSION>
__device__ __forceinline__ bool traceWorld(const OcWorld *__restrict__ world, Ray &ray, Stack st)
{
  bool any = false;
  const uint32_t n = world->numInstances;
  const vec3f rdir = safeRcpDir(ray.dir);
  if (n <= OC_TLAS_THRESHOLD) {
    for (uint32_t i = 0; i < n; i++) {
      const OcInstance *inst = world->instances + i;
      if (!rayBoxOverlap(inst->bounds, ray.org, rdir, ray.t0, ray.t))
        continue;
      if (traceInstance<OCCLUSION>(inst, i, ray, st, 0)) {
        any = true;
        if (OCCLUSION) return true;
      }
    }
    return any;
  }
  // TLAS: leaves hold instance indices in record[0].w; the BLAS traversal continues on the same stack above sp
  const OcBVH &tlas = world->tlas;
  int sp = 0;
  uint32_t cur = 0;
  while (true) {
    while (!(cur & 0x80000000u)) {
      float tn[4];
      uint4 ch;
      const uint32_t mask = nodeTest(tlas.nodes + cur, ray.org, rdir, ray.t0, ray.t, tn, ch);
      if (mask == 0) {
        if (sp == 0) return any;
        cur = st.at(--sp);
        continue;
      }
      const uint32_t chs[4] = {ch.x, ch.y, ch.z, ch.w};
      int best = 0;
      float bestT = INFINITY;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (((mask >> i) & 1u) && tn[i] < bestT) { bestT = tn[i]; best = i; }
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (((mask >> i) & 1u) && i != best && sp < OC_STACK_SIZE / 2) st.at(sp++) = chs[i];
      cur = chs[best];
    }
    const uint32_t first = cur & 0x07FFFFFFu;
    const uint32_t count = (cur >> 27) & 0xFu;
    for (uint32_t i = 0; i < count; i++) {
      const uint32_t instID = __float_as_uint(ldg4(tlas.prims + 3 * (size_t)(first + i)).w);
      if (traceInstance<OCCLUSION>(world->instances + instID, instID, ray, st, sp)) {
        any = true;
        if (OCCLUSION) return true;
      }
    }
    if (sp == 0) return any;
    cur = st.at(--sp);
  }
}

// one shared, non-inlined copy per kernel for code with many call sites
template <bool OCCLUSION>
__device__ __noinline__ bool traceWorldNI(const OcWorld *__restrict__ world, Ray &ray, Stack st)
{
  return traceWorld<OCCLUSION>(world, ray, st);
}

} // namespace oc
