// curve_capsule.h -- the round-curve stand-in shared by oracle.cpp and ref_embree_stub.cpp.  TEST INFRASTRUCTURE ONLY.
//
// Embree's round-curve sweep intersector (kernels/geometry/curve_intersector_sweep.h) is not vendored. The restatement
// used on every side (CUDA path, oracle, Embree stand-in of oracle/_ref) is a tessellation: each curve segment becomes
// CURVE_SUBDIV capsules whose end points are the curve basis (geometry/Curves.ispc:40-75 polynomials) at k/8 and whose
// radius is the mean of the two end radii; front hits only, u = (k + s)/8, Ng = P - axis(s). Plain C float operations,
// compiled with -ffp-contract=off: the CUDA side spells the same operations with __f*_rn intrinsics.
#pragma once
#include <cmath>
#include <cstdint>

namespace curvecap {
const uint32_t CURVE_SUBDIV = 8;
struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };
inline float dot3(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline F3 sub3(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 madd3(F3 a, float s, F3 b) { return {a.x + s * b.x, a.y + s * b.y, a.z + s * b.z}; }

// the cubic basis polynomials of geometry/Curves.ispc:40-75 at t (basis = OSPCurveBasis: 1 Bezier, 2 B-spline, 4 Catmull-Rom):
// weights b[0..3] of the four control points and the common factor c
inline void curveBasisWeights(int basis, float t, float b[4], float &c)
{
  const float s = 1.0f - t;
  const float s2 = s * s, t2 = t * t, s2t = s2 * t, t2s = t2 * s, s3 = s2 * s, t3 = t2 * t;
  if (basis == 1) { b[0] = s3; b[1] = 3.0f * s2t; b[2] = 3.0f * t2s; b[3] = t3; c = 1.0f; }
  else if (basis == 2) {
    b[0] = s3; b[1] = (4.0f * s3 + t3) + (12.0f * s2t + 6.0f * t2s); b[2] = (4.0f * t3 + s3) + (12.0f * t2s + 6.0f * s2t); b[3] = t3;
    c = 1.0f / 6.0f;
  } else { b[0] = -s2t; b[1] = 2.0f + t2 * (3.0f * t - 5.0f); b[2] = 2.0f + s2 * (3.0f * s - 5.0f); b[3] = -t2s; c = 0.5f; }
}
// basis == 0: linear over (v0, v1); else the cubic basis over (v0..v3)
inline F4 curveEval(int basis, F4 v0, F4 v1, F4 v2, F4 v3, float t)
{
  const float s = 1.0f - t;
  if (basis == 0) return {s * v0.x + t * v1.x, s * v0.y + t * v1.y, s * v0.z + t * v1.z, s * v0.w + t * v1.w};
  float bw[4], cf;
  curveBasisWeights(basis, t, bw, cf);
  const float b0 = bw[0], b1 = bw[1], b2 = bw[2], b3 = bw[3];
  auto comb = [&](float a, float b, float c, float d) { return cf * (((b0 * a + b1 * b) + b2 * c) + b3 * d); };
  return {comb(v0.x, v1.x, v2.x, v3.x), comb(v0.y, v1.y, v2.y, v3.y), comb(v0.z, v1.z, v2.z, v3.z), comb(v0.w, v1.w, v2.w, v3.w)};
}

// The ray origin is first moved along the ray to the point closest to pa (parameter tp): the quadratic's coefficients
// then carry |org' - pa| ~ the capsule's own size instead of the camera distance, so the discriminant of a grazing ray
// does not cancel catastrophically and the accepted rays stay within a few ulp of the exact capsule (the BVH culling
// box of bvh_build.cu relies on that). Embree's sweep intersector re-centres the ray the same way.
inline bool intersectCapsule(F3 org, F3 d, float t0, float t1, F3 pa, F3 pb, float r, float &t, float &sOut, F3 &Ng)
{
  const F3 ba = sub3(pb, pa), oa0 = sub3(org, pa);
  const float dd = dot3(d, d);
  const float tp = -(dot3(d, oa0) / dd);
  const F3 oa = madd3(oa0, tp, d);
  const float baba = dot3(ba, ba), bard = dot3(ba, d), baoa = dot3(ba, oa), rdoa = dot3(d, oa), oaoa = dot3(oa, oa);
  const float rr = r * r;
  const float a = baba * dd - bard * bard;
  float y;
  if (a > 0.f) {
    const float b = baba * rdoa - baoa * bard;
    const float c = (baba * oaoa - baoa * baoa) - rr * baba;
    const float h = b * b - a * c;
    if (!(h >= 0.f)) return false;
    const float tb = (-b - sqrtf(h)) / a;
    y = baoa + tb * bard;
    if (y > 0.f && y < baba) {
      const float tt = tp + tb;
      if (!(tt > t0 && tt <= t1)) return false;
      const float s = y / baba;
      t = tt; sOut = s;
      Ng = madd3(madd3(oa, tb, d), -s, ba);
      return true;
    }
  } else
    y = bard > 0.f ? -1.f : baba + 1.f;
  const bool first = y <= 0.f;
  const F3 oc = first ? oa : sub3(oa, ba);
  const float b2 = dot3(d, oc);
  const float c2 = dot3(oc, oc) - rr;
  const float h2 = b2 * b2 - dd * c2;
  if (!(h2 > 0.f)) return false;
  const float tc = (-b2 - sqrtf(h2)) / dd;
  const float tt = tp + tc;
  if (!(tt > t0 && tt <= t1)) return false;
  t = tt; sOut = first ? 0.f : 1.f;
  Ng = madd3(oc, tc, d);
  return true;
}
} // namespace curvecap
