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
#define OC_STACK_TOTAL OSPCUDA_TRAVERSAL_STACK // include/ospcuda.h: the host checks every built hierarchy against it
#ifndef OC_STACK_SHARED
#define OC_STACK_SHARED 32 // entries per thread kept in shared memory; deeper ones go to the local overflow array (Stack)
#endif
#define OC_STACK_OVF (OC_STACK_TOTAL - OC_STACK_SHARED)
#define OC_MAX_BLOCK 128 // largest CTA of a kernel that traverses (shared stack = OC_STACK_SHARED * blockDim.x words)

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

// What traversal carries per ray: 5 registers. The geometric normal and the ids are NOT live during the walk; they are
// recomputed from the leaf record the hit points at once the ray is done (resolveHit below). `rec` indexes the records of
// the hit instance's group BVH; inst == OC_INVALID_ID: no hit.
struct Hit { float t, u, v; uint32_t rec, inst; };

// the triangle test without the normal output (traversal form of intersectTriangle: same arithmetic, same decisions)
__device__ __forceinline__ bool intersectTriangleT(vec3f org, vec3f dir, float t0, float t1, vec3f v0, vec3f e1, vec3f e2,
    float &t, float &u, float &v)
{
  const vec3f triNg = crossF(e2, e1);
  const vec3f C = make_vec3f(__fsub_rn(v0.x, org.x), __fsub_rn(v0.y, org.y), __fsub_rn(v0.z, org.z));
  const vec3f R = crossF(C, dir);
  const float den = dotF(triNg, dir);
  const float absDen = fabsf(den);
  const uint32_t sgn = __float_as_uint(den) & 0x80000000u;
  const float U = __uint_as_float(__float_as_uint(dotF(R, e2)) ^ sgn);
  if (!(den != 0.f && U >= 0.f && U <= absDen))
    return false;
  const float V = __uint_as_float(__float_as_uint(dotF(R, e1)) ^ sgn);
  if (!(V >= 0.f && __fadd_rn(U, V) <= absDen))
    return false;
  const float T = __uint_as_float(__float_as_uint(dotF(triNg, C)) ^ sgn);
  if (!(__fmul_rn(absDen, t0) < T && T <= __fmul_rn(absDen, t1)))
    return false;
  const float rcpAbsDen = __frcp_rn(absDen);
  u = __fmul_rn(U, rcpAbsDen);
  v = __fmul_rn(V, rcpAbsDen);
  t = __fmul_rn(T, rcpAbsDen);
  return true;
}

// one leaf record against the ray in the current (instance) space. Exact-t ties with the hit already held (coplanar
// faces, shared edges) are resolved by record identity -- the later (instance, geometry, primitive, quad half) wins, what
// a scan in record order (the oracle's and the Embree stand-in's loops) ends with -- so the result does not depend on the
// BVH topology or the visiting order. The ids of the held hit are read back from its record only in that rare case.
__device__ __forceinline__ bool intersectRecordH(vec3f org, vec3f dir, float t0, Hit &h, const OcVec4f *__restrict__ prims,
    uint32_t recIdx, uint32_t instID, bool occlusion)
{
  const OcVec4f *rec = prims + 3 * (size_t)recIdx;
  const float4 r0 = ldg4(rec), r1 = ldg4(rec + 1), r2 = ldg4(rec + 2);
  const uint32_t kind = __float_as_uint(r2.w);
  float t, u = 0.f, v = 0.f;
  bool hit;
  if (kind <= OC_PRIM_TRI_QUAD2ND) {
    hit = intersectTriangleT(org, dir, t0, h.t, make_vec3f(r0.x, r0.y, r0.z), make_vec3f(r1.x, r1.y, r1.z),
        make_vec3f(r2.x, r2.y, r2.z), t, u, v);
    if (hit && kind == OC_PRIM_TRI_QUAD2ND) { u = __fsub_rn(1.f, u); v = __fsub_rn(1.f, v); }
  } else {
    const OtherHit o = intersectOtherNI(org, dir, t0, h.t, r0, r1, r2);
    hit = o.hit; t = o.t; u = o.u;
  }
  if (!hit)
    return false;
  if (!occlusion && t == h.t && h.inst != OC_INVALID_ID) {
    bool later;
    if (instID != h.inst) later = instID > h.inst;
    else {
      const OcVec4f *held = prims + 3 * (size_t)h.rec;
      const uint32_t hPrim = __float_as_uint(ldg4(held).w), hGeom = __float_as_uint(ldg4(held + 1).w);
      const uint32_t primID = __float_as_uint(r0.w), geomID = __float_as_uint(r1.w);
      later = geomID != hGeom ? geomID > hGeom : primID != hPrim ? primID > hPrim : kind == OC_PRIM_TRI_QUAD2ND;
    }
    if (!later) return false;
  }
  h.t = t; h.u = u; h.v = v; h.rec = recIdx; h.inst = instID;
  return true;
}

// ---- wide-BVH traversal -------------------------------------------------------------------
// slab test of one 128-byte node against the ray: 7 x LDG.128, conservative (far side padded by 2 ulp) so that the BVH
// never culls a primitive the exact test would accept. The planes the ray enters / leaves through are picked by the sign of
// the direction when the node is LOADED (lo and hi rows sit 16 bytes apart), not by min / max afterwards: (lo - o) * rd <=
// (hi - o) * rd for rd >= 0 by monotonic rounding, so the values are the ones min / max would have kept -- 6 instructions
// less per child.
__device__ __forceinline__ uint32_t nodeTest(const OcNode4 *node, vec3f org, vec3f rdir, float t0, float t1,
    float (&tnear)[4], uint4 &child)
{
  const float4 *n = reinterpret_cast<const float4 *>(node);
  const int sx = rdir.x < 0.f, sy = rdir.y < 0.f, sz = rdir.z < 0.f;
  const float4 nx = __ldg(n + sx), fx = __ldg(n + (sx ^ 1)), ny = __ldg(n + 2 + sy), fy = __ldg(n + 2 + (sy ^ 1)),
               nz = __ldg(n + 4 + sz), fz = __ldg(n + 4 + (sz ^ 1));
  child = __ldg(reinterpret_cast<const uint4 *>(n + 6));
  uint32_t mask = 0;
#define OC_SLAB(i, NX, FX, NY, FY, NZ, FZ, CH)                                                   \
  {                                                                                              \
    const float ax = (NX - org.x) * rdir.x, bx = (FX - org.x) * rdir.x;                          \
    const float ay = (NY - org.y) * rdir.y, by = (FY - org.y) * rdir.y;                          \
    const float az = (NZ - org.z) * rdir.z, bz = (FZ - org.z) * rdir.z;                          \
    const float tn = fmaxf(fmaxf(ax, ay), fmaxf(az, t0));                                        \
    const float tf = fminf(fminf(bx, by), fminf(bz, t1)) * 1.0000005f;                           \
    tnear[i] = tn;                                                                               \
    if (CH != OC_INVALID_ID && tn <= tf) mask |= 1u << i;                                        \
  }
  OC_SLAB(0, nx.x, fx.x, ny.x, fy.x, nz.x, fz.x, child.x)
  OC_SLAB(1, nx.y, fx.y, ny.y, fy.y, nz.y, fz.y, child.y)
  OC_SLAB(2, nx.z, fx.z, ny.z, fy.z, nz.z, fz.z, child.z)
  OC_SLAB(3, nx.w, fx.w, ny.w, fy.w, nz.w, fz.w, child.w)
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

// Per-thread traversal stack. The first OC_STACK_SHARED entries live in SHARED memory, laid out [depth][thread] so that
// the 32 lanes of a warp hit 32 different banks whatever their individual depths are; deeper entries (up to
// OSPCUDA_TRAVERSAL_STACK in total: the host checks every hierarchy it builds against that) spill into a per-thread
// local-memory array that costs nothing until a walk actually gets that deep. Kernels declare
// `__shared__ uint32_t s[OC_STACK_SHARED * blockDim.x]` and `uint32_t ovf[OC_STACK_OVF]`.
struct Stack
{
  uint32_t *base;   // s + threadIdx.x
  uint32_t stride;  // blockDim.x
  uint32_t *ovf;    // local overflow
  __device__ __forceinline__ void push(int &sp, uint32_t v) const
  {
    if (sp < OC_STACK_SHARED) { base[(uint32_t)sp * stride] = v; sp++; }
    else if (sp < OC_STACK_TOTAL) { ovf[sp - OC_STACK_SHARED] = v; sp++; }
  }
  __device__ __forceinline__ uint32_t pop(int &sp) const
  {
    --sp;
    return sp < OC_STACK_SHARED ? base[(uint32_t)sp * stride] : ovf[sp - OC_STACK_SHARED];
  }
};

// where a walk re-reads its world-space ray from when it leaves an instance: values held in registers (megakernels,
// single queries) or the ray arrays of a queue kernel (global memory: 6 registers less across the whole walk)
struct RayRegs
{
  vec3f o, d;
  __device__ __forceinline__ vec3f org() const { return o; }
  __device__ __forceinline__ vec3f dir() const { return d; }
};
struct RayMem
{
  const float4 *po, *pd;
  __device__ __forceinline__ vec3f org() const { const float4 v = *po; return make_vec3f(v.x, v.y, v.z); }
  __device__ __forceinline__ vec3f dir() const { const float4 v = *pd; return make_vec3f(v.x, v.y, v.z); }
};

// Two-level traversal as ONE resumable state machine. The top level (World::commit builds it over the instance bounds,
// one instance per leaf) and the group BVHs share the loop and the stack: a top-level leaf switches the walk into the
// instance (ray -> instance space, Embree keeps t, no renormalisation; spBase marks where the group's entries start), an
// empty group part of the stack switches it back. Equivalent of traceGeometryRay / isOccludedNoClipping (World.ih:29-52,
// 311-335). The walk is cut into the two steps a warp wants to take together -- inner(): one node, leaf(): one leaf --
// so that a kernel can line its lanes up between them (ptTraceKernel); traceHit() below is the plain loop.
// The hit children of a node are ordered by entry distance (5 compare-exchanges), the nearest is visited next and the rest
// pushed far-to-near (C3 106 -> 94 ms against an unordered push); any-hit rays take the same order.
// MODE 0: closest hit, 1: occlusion, 2: decided per ray at run time (`occ`; queue kernels that mix both kinds).
template <int MODE>
struct Traversal
{
  vec3f org, dir, rdir;
  float t0;
  Hit h;
  const OcNode4 *nodes;
  const OcVec4f *prims;
  uint32_t cur, curInst;
  int sp, spBase;
  bool occ, xformed; // xformed: org / dir are in the space of a non-identity instance

  __device__ __forceinline__ bool occlusion() const { return MODE == 2 ? occ : MODE == 1; }
  __device__ __forceinline__ bool atLeaf() const { return (cur & 0x80000000u) != 0u; }

  __device__ __forceinline__ void enterInstance(const OcWorld *__restrict__ world, uint32_t instID, const OcGroup *__restrict__ g)
  {
    const OcInstance *__restrict__ inst = world->instances + instID;
    xformed = !inst->isIdentity;
    if (xformed) {
      const affine3f rx = toA(inst->rcp_xfm);
      const vec3f wo = org, wd = dir;
      org = xfmPointExact(rx, wo);
      dir = xfmVectorExact(rx.l, wd);
      rdir = safeRcpDir(dir);
    }
    nodes = g->bvh.nodes; prims = g->bvh.prims;
    curInst = instID; spBase = sp; cur = 0;
  }

  // false: nothing to traverse (empty world)
  __device__ __forceinline__ bool init(const OcWorld *__restrict__ world, vec3f rorg, vec3f rdirection, float rt0, float rt1, bool occlusionRay)
  {
    occ = occlusionRay; xformed = false;
    h.t = rt1; h.u = h.v = 0.f; h.rec = 0; h.inst = OC_INVALID_ID;
    t0 = rt0;
    nodes = world->tlas.nodes;
    prims = world->tlas.prims;
    if (nodes == nullptr) return false;
    org = rorg; dir = rdirection;
    rdir = safeRcpDir(dir);
    cur = 0; curInst = OC_INVALID_ID; sp = 0; spBase = 0;
    if (world->numInstances == 1) { // the top level is one leaf: start inside the instance (its root node does the bounds test)
      const OcGroup *__restrict__ g = world->instances[0].group;
      if (g->bvh.numRecords == 0) return false;
      enterInstance(world, 0u, g);
    }
    return true;
  }

  // the group part of the stack is empty: back to the top level; false when that is empty too (walk finished)
  template <typename Src>
  __device__ __forceinline__ bool popNext(const OcWorld *__restrict__ world, const Src &src, const Stack &st)
  {
    if (curInst != OC_INVALID_ID && sp == spBase) {
      if (sp == 0) return false;
      if (xformed) { org = src.org(); dir = src.dir(); rdir = safeRcpDir(dir); xformed = false; }
      nodes = world->tlas.nodes; prims = world->tlas.prims;
      curInst = OC_INVALID_ID;
    }
    if (sp == 0) return false;
    cur = st.pop(sp);
    return true;
  }

  // cur is an inner node: test it, descend into the nearest hit child (or pop). false: the walk is finished
  template <typename Src>
  __device__ __forceinline__ bool inner(const OcWorld *__restrict__ world, const Src &src, const Stack &st)
  {
    float tn[4];
    uint4 ch;
    const uint32_t mask = nodeTest(nodes + cur, org, rdir, t0, h.t, tn, ch);
    if (mask == 0) return popNext(world, src, st);
    if (occlusion()) { // any-hit walks: nearest child next, the others in node order (no sort; queue kernels fetch shadow rays
                       // before continuation rays, so warps are of one kind except at the boundary)
      const uint32_t chs[4] = {ch.x, ch.y, ch.z, ch.w};
      int best = 0;
      float bestT = INFINITY;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (((mask >> i) & 1u) && tn[i] < bestT) { bestT = tn[i]; best = i; }
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (((mask >> i) & 1u) && i != best) st.push(sp, chs[i]);
      cur = chs[best];
      return true;
    }
    float k0 = (mask & 1u) ? tn[0] : INFINITY, k1 = (mask & 2u) ? tn[1] : INFINITY, k2 = (mask & 4u) ? tn[2] : INFINITY,
          k3 = (mask & 8u) ? tn[3] : INFINITY;
    uint32_t c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
#define OC_CSWAP(ka, ca, kb, cb) { const bool sw = kb < ka; const float tk = sw ? kb : ka; kb = sw ? ka : kb; ka = tk; const uint32_t tc = sw ? cb : ca; cb = sw ? ca : cb; ca = tc; }
    OC_CSWAP(k0, c0, k1, c1) OC_CSWAP(k2, c2, k3, c3) OC_CSWAP(k0, c0, k2, c2) OC_CSWAP(k1, c1, k3, c3) OC_CSWAP(k1, c1, k2, c2)
#undef OC_CSWAP
    if (k3 < INFINITY) st.push(sp, c3);
    if (k2 < INFINITY) st.push(sp, c2);
    if (k1 < INFINITY) st.push(sp, c1);
    cur = c0;
    return true;
  }

  // cur is a leaf. Top level: enter the instance of its first record (the rest of the leaf -- the builder makes
  // one-instance leaves, this keeps any other count correct -- goes back on the stack). Group: test the records, pop.
  // false: the walk is finished (stack empty, or an occlusion ray found its hit)
  template <typename Src>
  __device__ __forceinline__ bool leaf(const OcWorld *__restrict__ world, const Src &src, const Stack &st)
  {
    const uint32_t first = cur & 0x07FFFFFFu;
    const uint32_t count = (cur >> 27) & 0xFu;
    if (curInst == OC_INVALID_ID) {
      if (count > 1) st.push(sp, 0x80000000u | ((count - 1u) << 27) | (first + 1u));
      const uint32_t instID = __float_as_uint(ldg4(prims + 3 * (size_t)first).w);
      const OcGroup *__restrict__ g = world->instances[instID].group;
      if (g->bvh.numRecords != 0) {
        enterInstance(world, instID, g);
        return true;
      }
    } else {
      for (uint32_t i = 0; i < count; i++)
        if (intersectRecordH(org, dir, t0, h, prims, first + i, curInst, occlusion()) && occlusion()) return false;
    }
    return popNext(world, src, st);
  }
};

// a whole walk, for code that traces one ray at a time (megakernels, batch queries, the tail). The lanes that arrive here
// together step together: every round the warp counts how many of them stand at an inner node and how many at a leaf and
// takes the step the majority wants (one node test, or one leaf), the others wait. Left to itself the compiler lets every
// lane run ahead on its own and the node / primitive tests execute a few lanes at a time.
template <int MODE>
__device__ __forceinline__ Hit traceHit(const OcWorld *__restrict__ world, vec3f org, vec3f dir, float t0, float t1, const Stack &st)
{
  RayRegs src; src.o = org; src.d = dir;
  Traversal<MODE> tr;
  const unsigned together = __activemask();
  bool go = tr.init(world, org, dir, t0, t1, MODE == 1);
  while (true) {
    const bool wantInner = go && !tr.atLeaf();
    const uint32_t nGo = __popc(__ballot_sync(together, go)), nInner = __popc(__ballot_sync(together, wantInner));
    if (nGo == 0) break;
    if (2 * nInner >= nGo) {
      if (wantInner) go = tr.inner(world, src, st);
    } else {
      if (go && !wantInner) go = tr.leaf(world, src, st);
    }
  }
  return tr.h;
}

// the hit's geometric normal (instance space, unnormalised: World.ih:411-420) and ids, recomputed from its leaf record.
// Triangles: Ng = cross(e2, e1) of the record. Everything else re-runs the primitive test on the ray clipped to the hit
// distance: the tests are deterministic in (org, dir, t0, record) and take ray.t only as an upper bound, so the same t and
// normal come back.
__device__ __forceinline__ void resolveHit(const OcWorld *__restrict__ world, const Hit &h, Ray &ray)
{
  ray.t = h.t; ray.u = h.u; ray.v = h.v;
  ray.instID = h.inst;
  if (h.inst == OC_INVALID_ID) { ray.primID = ray.geomID = OC_INVALID_ID; ray.Ng = make_vec3f(0.f); return; }
  const OcInstance *__restrict__ inst = world->instances + h.inst;
  const OcVec4f *rec = inst->group->bvh.prims + 3 * (size_t)h.rec;
  const float4 r0 = ldg4(rec), r1 = ldg4(rec + 1), r2 = ldg4(rec + 2);
  ray.primID = __float_as_uint(r0.w);
  ray.geomID = __float_as_uint(r1.w);
  const uint32_t kind = __float_as_uint(r2.w);
  if (kind <= OC_PRIM_TRI_QUAD2ND)
    ray.Ng = crossF(make_vec3f(r2.x, r2.y, r2.z), make_vec3f(r1.x, r1.y, r1.z));
  else {
    vec3f lo = ray.org, ld = ray.dir;
    if (!inst->isIdentity) {
      const affine3f rx = toA(inst->rcp_xfm);
      lo = xfmPointExact(rx, ray.org);
      ld = xfmVectorExact(rx.l, ray.dir);
    }
    ray.Ng = intersectOtherNI(lo, ld, ray.t0, h.t, r0, r1, r2).Ng;
  }
}
// ids only (occluder checks)
__device__ __forceinline__ void resolveHitIds(const OcWorld *__restrict__ world, uint32_t rec, uint32_t instID, uint32_t &primID, uint32_t &geomID)
{
  const OcVec4f *r = world->instances[instID].group->bvh.prims + 3 * (size_t)rec;
  primID = __float_as_uint(ldg4(r).w);
  geomID = __float_as_uint(ldg4(r + 1).w);
}

// one shared, non-inlined copy per kernel for code with many call sites (the scivis / ao / debug megakernels). The hit
// comes back by value and the Ray of the caller stays in registers.
template <int MODE>
static __device__ __noinline__ Hit traceHitNI(const OcWorld *__restrict__ world, vec3f org, vec3f dir, float t0, float t1, uint32_t *stackBase,
    uint32_t stackStride)
{
  uint32_t ovf[OC_STACK_OVF];
  Stack st;
  st.base = stackBase; st.stride = stackStride; st.ovf = ovf;
  return traceHit<MODE>(world, org, dir, t0, t1, st);
}
// Ray-in / Ray-out form of the above: closest hit fills t, u, v, Ng and the ids; occlusion returns whether anything was hit
template <bool OCCLUSION>
__device__ __forceinline__ bool traceWorldNI(const OcWorld *__restrict__ world, Ray &ray, const Stack &st)
{
  const Hit h = traceHitNI<OCCLUSION ? 1 : 0>(world, ray.org, ray.dir, ray.t0, ray.t, st.base, st.stride);
  if (OCCLUSION) return h.inst != OC_INVALID_ID;
  resolveHit(world, h, ray);
  return h.inst != OC_INVALID_ID;
}

} // namespace oc
