// oracle.cpp -- scalar CPU restatement of the OSPRay v3.3.0 render path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
// load liboracle.so, and only as the checker or the timed CPU baseline; nothing under
// ospray_b200/ links or imports it. The product path is libospcuda.so (sm_100a kernels).
//
// PARITY STATUS: pinned against the reference's own kernels for everything that lives in /root/reference. The
// reference's modules/cpu/**/*.ispc compile as scalar C++ (the mode its SYCL target uses); oracle/build_ref.py builds
// them where they lie into oracle/_ref/libospref.so, and tests/test_oracle_vs_reference.py checks this file against
// that build bit for bit (sampler streams, pixel-filter LUT, frame ops, accumulation) and pixel for pixel (PT / SciVis /
// AO renders of every tests/scenes.py case), live and through the committed vectors tests/golden/ref_*.npz.
// UNPINNED remains only the arithmetic of dependencies that are not vendored (needs ISPC, Embree 4.3.3, Open VKL 2.0.1,
// rkcommon 1.14.2 to build the real CPU device): both this file and the _ref build use the same restatement of those
// pieces, listed below. This file restates, function by function, the in-tree sources listed beside each routine
// (paths relative to /root/reference/modules/cpu); the arithmetic that lives in the absent
// dependencies is restated from their published algorithms and named where used:
//   Embree 4.3.3  triangle (Moeller-Trumbore, e1=v0-v1, e2=v2-v0, Ng=cross(e2,e1)), quad split
//                 (v0,v1,v3)+(v2,v3,v1) with (1-u,1-v) (mirrored in geometry/Mesh.ispc:315-345),
//                 sphere with projected centre
//   Open VKL 2.0.1 structuredRegular vertex-centred trilinear sample (NaN outside), forward
//                 difference gradient, nominalDeltaT = min_i spacing_i/|dir_i|
//                 round curves: Embree's sweep intersector is replaced by a stated tessellation into 8 capsules per segment
//                 (curve_capsule.h); under the pathtracer the interval iterator hands delta tracking ONE interval with the
//                 value range (-inf, inf) -> global majorant (real VKL: macrocell intervals with tighter ranges)
//   rkcommon 1.14.2 to_float_unorm, frame(), lerp(f,a,b), luminance, rcp_safe, cvt_uint32,
//                 linear_to_srgba8, makeRandomColor (debug renderer)
// Intersection is brute force (no BVH): the oracle is independent of the GPU builder/traversal.
// The volume interval iterator is restated as ONE interval over the clipped box (mathematically
// equal to any macrocell subdivision: skipped cells have zero opacity and `emptySpace`
// compensates dt), so the GPU's macrocell DDA is checked against plain marching.
//
// Scene description: the POD structs of include/ospcuda.h with HOST pointers (OcBVH members are
// ignored). Build: make -C oracle  (g++ -O2 -ffp-contract=off; fmaf is spelled where Embree's
// kernels fuse so that hit/miss decisions match the GPU bit for bit).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "../include/ospcuda.h"
#include "tables.inc"
#include "curve_capsule.h"

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace {

const float INF = INFINITY;
const uint32_t INVALID_ID = 0xFFFFFFFFu;
const float PI = 3.14159265358979323846f, TWO_PI = 6.28318530717958647692f, ONE_OVER_PI = 0.31830988618379067154f,
            ONE_OVER_TWO_PI = 0.15915494309189533577f, ONE_OVER_FOUR_PI = 0.079577471545947667884f,
            FOUR_PI = 12.566370614359172954f, LNTWO_OVER_TWO = 0.34657359027997265471f;
const float ULP_EPS = 0x1.0fp-21f; // common/DifferentialGeometry.ih:46

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct L3 { V3 vx, vy, vz; };
struct A3 { L3 l; V3 p; };

inline V3 v3(float x, float y, float z) { return {x, y, z}; }
inline V3 v3(float s) { return {s, s, s}; }
inline V3 v3(const OcVec3f &o) { return {o.x, o.y, o.z}; }
inline V3 v3(const V4 &o) { return {o.x, o.y, o.z}; }
inline V4 v4(V3 a, float w) { return {a.x, a.y, a.z, w}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator*(float s, V4 a) { return {s * a.x, s * a.y, s * a.z, s * a.w}; }
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, V2 b) { return {a.x * b.x, a.y * b.y}; }
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { return a * (1.f / sqrtf(dot(a, a))); }
inline float rmax(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
inline float rmin(V3 a) { return fminf(a.x, fminf(a.y, a.z)); }
inline V3 vabs(V3 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }
inline V3 vmin(V3 a, V3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline float sqr(float x) { return x * x; }
inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
inline float lerp(float f, float a, float b) { return (1.f - f) * a + f * b; } // rkcommon: factor first
inline V3 lerp(float f, V3 a, V3 b) { return (1.f - f) * a + f * b; }
inline V4 lerp(float f, V4 a, V4 b) { return (1.f - f) * a + f * b; }
inline float luminance(V3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }
inline bool eq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline V3 xfmVector(const L3 &l, V3 v) { return v.x * l.vx + v.y * l.vy + v.z * l.vz; }
inline V3 xfmPoint(const A3 &a, V3 v) { return xfmVector(a.l, v) + a.p; }
inline L3 transposed(const L3 &l) { return {{l.vx.x, l.vy.x, l.vz.x}, {l.vx.y, l.vy.y, l.vz.y}, {l.vx.z, l.vy.z, l.vz.z}}; }
inline A3 a3(const OcAffine3f &a) { return {{v3(a.vx), v3(a.vy), v3(a.vz)}, v3(a.p)}; }
inline L3 frame(V3 N) // rkcommon math/LinearSpace.h frame()
{
  const V3 dx0 = v3(0.f, N.z, -N.y), dx1 = v3(-N.z, 0.f, N.x);
  const V3 dx = normalize(dot(dx0, dx0) > dot(dx1, dx1) ? dx0 : dx1);
  const V3 dy = normalize(cross(N, dx));
  return {dx, dy, N};
}

// ================================ samplers (math/random.ih, sobol.ih, halton.ih) ===============
inline float to_float_unorm(uint32_t a) { return (float)a * 0x1.0p-32f; } // rkcommon
inline uint32_t reverseBits(uint32_t i) // halton.ih:38-47
{
  i = (i << 16) | (i >> 16);
  i = ((i & 0x00ff00ffu) << 8) | ((i & 0xff00ff00u) >> 8);
  i = ((i & 0x0f0f0f0fu) << 4) | ((i & 0xf0f0f0f0u) >> 4);
  i = ((i & 0x33333333u) << 2) | ((i & 0xccccccccu) >> 2);
  i = ((i & 0x55555555u) << 1) | ((i & 0xaaaaaaaau) >> 1);
  return i;
}
inline uint32_t MurmurHash3_mix(uint32_t hash, uint32_t k) // random.ih:15-32
{
  k *= 0xcc9e2d51u; k = (k << 15) | (k >> 17); k *= 0x1b873593u;
  hash ^= k;
  return ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
}
inline uint32_t MurmurHash3_finalize(uint32_t h) // random.ih:34-43
{
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
inline uint32_t tripleHash(uint32_t x) // random.ih:46-56
{
  x ^= x >> 17; x *= 0xed5ad4bbu; x ^= x >> 11; x *= 0xac4c1b51u; x ^= x >> 15; x *= 0x31848babu; x ^= x >> 14;
  return x;
}
inline float CranleyPattersonRotation(float x, float dx) { x += dx; if (x >= 1.f) x -= 1.f; return x; } // :84-91

struct RandomSampler { uint64_t state; uint32_t stream; }; // random.ih:116-149
inline uint32_t pcg32(RandomSampler &s)
{
  const uint64_t old = s.state;
  s.state = old * 6364136223846793005ULL + (s.stream | 1u);
  const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  const uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}
inline float RandomSampler_getFloat(RandomSampler &s) { return to_float_unorm(pcg32(s)); }
inline void RandomSampler_init(RandomSampler &s, uint32_t seed, uint32_t stream)
{
  s.state = 0;
  s.stream = (stream << 1u) | 1u;
  s.state = MurmurHash3_mix((uint32_t)s.state, seed);
  s.state = MurmurHash3_finalize((uint32_t)s.state);
  pcg32(s);
  s.state += seed;
  pcg32(s);
}
inline void tea8(uint32_t &v0, uint32_t &v1) // random.ih:159-172
{
  uint32_t sum = 0;
  for (int i = 0; i < 8; i++) {
    sum += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
  }
}
inline uint32_t OwenHash2(uint32_t x, uint32_t seed) // random.ih:234-243
{
  x ^= x * 0x3d20adeau; x += seed; x *= (seed >> 16) | 1u; x ^= x * 0x05526c56u; x ^= x * 0x53a22864u;
  return x;
}
inline uint32_t OwenScramble4(uint32_t x, uint32_t seed) // random.ih:252-266
{
  x = reverseBits(x);
  x ^= x * 0x3d20adeau;
  x ^= (x >> 1) & (x << 1) & 0x55555555u;
  x += seed;
  x *= (seed >> 16) | 1u;
  x ^= (x >> 1) & (x << 1) & 0x55555555u;
  x ^= x * 0x05526c56u;
  x ^= x * 0x53a22864u;
  return reverseBits(x);
}
inline uint32_t partitionZOrder8(uint32_t n) // render/util.ih:33-40
{
  n &= 0x000000FFu; n = (n | (n << 4)) & 0x0F0F0F0Fu; n = (n | (n << 2)) & 0x33333333u; n = (n | (n << 1)) & 0x55555555u;
  return n;
}
inline uint32_t interleaveZOrder16(uint32_t x, uint32_t y) { return partitionZOrder8(x) | (partitionZOrder8(y) << 1); }

inline void Sobol_revSample(uint32_t revIndex, int dims, uint32_t r[4]) // sobol.ih:38-80
{
  r[0] = r[1] = r[2] = r[3] = 0;
  for (uint32_t i = 0; revIndex; revIndex <<= 1, i += 4)
    if (revIndex & 0x80000000u)
      for (int d = 0; d < dims; d++) r[d] ^= kSobolRev[i + d];
}
const uint32_t SCRAMBLE0 = 0x9374b0d3u, SCRAMBLE1 = 0x1ec38426u, SCRAMBLE2 = 0xdb81c29cu, SCRAMBLE3 = 0x5628372bu,
               SCRAMBLE4 = 0xbcb745bdu;
struct LDSampler { uint32_t revIndex, scramble, baseIndexLight, idxMask, idxMaskLight; }; // random.ih:283-394
inline void LDSampler_init(LDSampler &s, uint32_t seed, uint32_t sampleIndex, uint32_t sampleOffsetLight, uint32_t bits,
    uint32_t bitsLight)
{
  s.revIndex = reverseBits(sampleIndex);
  s.scramble = seed;
  s.baseIndexLight = sampleOffsetLight;
  s.idxMask = bits > 31 ? ~0u : ~(~0u >> bits);
  s.idxMaskLight = bitsLight > 31 ? ~0u : ~(~0u >> bitsLight);
}
inline void LDSampler_nextGroup(LDSampler &s) { s.scramble = tripleHash(s.scramble); }
inline float LDSamplerScramble(uint32_t sample, uint32_t scramble) { return to_float_unorm(reverseBits(OwenHash2(sample, scramble))); }
inline V2 LDSampler_get3LightSamples(const LDSampler &s, uint32_t idx, bool split, float &ss)
{
  uint32_t revIndex = split ? reverseBits(s.baseIndexLight + idx) : s.revIndex;
  revIndex = OwenHash2(revIndex, s.scramble);
  revIndex &= split ? s.idxMaskLight : s.idxMask;
  const uint32_t r0 = reverseBits(revIndex);
  uint32_t r[4];
  Sobol_revSample(revIndex, 2, r);
  V2 res;
  res.x = LDSamplerScramble(r0, s.scramble ^ SCRAMBLE0);
  res.y = LDSamplerScramble(r[0], s.scramble ^ SCRAMBLE1);
  ss = LDSamplerScramble(r[1], s.scramble ^ SCRAMBLE2);
  return res;
}
inline V2 LDSampler_getNext4Samples(LDSampler &s, float &ss, uint32_t &u)
{
  LDSampler_nextGroup(s);
  const uint32_t revIndex = OwenHash2(s.revIndex, s.scramble) & s.idxMask;
  const uint32_t r0 = reverseBits(revIndex);
  uint32_t r[4];
  Sobol_revSample(revIndex, 3, r);
  V2 res;
  res.x = LDSamplerScramble(r0, s.scramble ^ SCRAMBLE0);
  res.y = LDSamplerScramble(r[0], s.scramble ^ SCRAMBLE1);
  ss = LDSamplerScramble(r[1], s.scramble ^ SCRAMBLE2);
  u = r[2];
  return res;
}
inline float LDSampler_finalizeDim3(const LDSampler &s, uint32_t revRes) { return LDSamplerScramble(revRes, s.scramble ^ SCRAMBLE3); }
inline V2 LDSampler_getNext5Samples(LDSampler &s, uint32_t ss[3])
{
  LDSampler_nextGroup(s);
  const uint32_t revIndex = OwenHash2(s.revIndex, s.scramble) & s.idxMask;
  const uint32_t r0 = reverseBits(revIndex);
  uint32_t r[4];
  Sobol_revSample(revIndex, 4, r);
  V2 res;
  res.x = LDSamplerScramble(r0, s.scramble ^ SCRAMBLE0);
  res.y = LDSamplerScramble(r[0], s.scramble ^ SCRAMBLE1);
  ss[0] = r[1]; ss[1] = r[2]; ss[2] = r[3];
  return res;
}
inline V2 LDSampler_finalizeDim23(const LDSampler &s, const uint32_t rr[3])
{
  return {LDSamplerScramble(rr[0], s.scramble ^ SCRAMBLE2), LDSamplerScramble(rr[1], s.scramble ^ SCRAMBLE3)};
}
inline float Halton_sample2(uint32_t index) { return to_float_unorm(reverseBits(index)); } // halton.ih:49-52
inline float Halton_sample3(uint32_t index) // halton.ih:54-67
{
  return (float)(kHaltonPerm3[index % 243u] * 14348907u + kHaltonPerm3[(index / 243u) % 243u] * 59049u
             + kHaltonPerm3[(index / 59049u) % 243u] * 243u + kHaltonPerm3[(index / 14348907u) % 243u])
      * (float)(1.0 / 3486784401u);
}
inline float PermHalton_sample5(uint32_t index) // halton.ih:70-77
{
  return (float)(kHaltonPerm5[index % 125u] * 1953125u + kHaltonPerm5[(index / 125u) % 125u] * 15625u
             + kHaltonPerm5[(index / 15625u) % 125u] * 125u + kHaltonPerm5[(index / 1953125u) % 125u])
      * (float)(1.0 / 244140625u);
}
inline float PermHalton_sample7(uint32_t index) // halton.ih:80-86
{
  return (float)(kHaltonPerm7[index % 343u] * 117649u + kHaltonPerm7[(index / 343u) % 343u] * 343u
             + kHaltonPerm7[(index / 117649u) % 343u])
      * (float)(1.0 / 40353607u);
}

// ================================ rays and primitive tests =====================================
struct Ray // common/Ray.ih:15-42
{
  V3 org; float t0; V3 dir; float t; V3 Ng; float u, v; uint32_t primID, geomID, instID;
};
inline void setRay(Ray &r, V3 org, V3 dir, float t0, float t1)
{
  r.org = org; r.dir = dir; r.t0 = t0; r.t = t1;
  r.Ng = v3(0.f); r.u = r.v = 0.f;
  r.primID = r.geomID = r.instID = INVALID_ID;
}
inline void setRayInterval(Ray &r, float t0, float t1) // setRay(ray, t0, t1): keeps org/dir, resets ids
{
  r.t0 = t0; r.t = t1; r.primID = r.geomID = r.instID = INVALID_ID;
}
inline bool hadHit(const Ray &r) { return r.geomID != INVALID_ID; }

inline float dotF(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline V3 crossF(V3 a, V3 b) { return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))}; }
inline float xorSign(float f, uint32_t sgn)
{
  uint32_t b; memcpy(&b, &f, 4); b ^= sgn; memcpy(&f, &b, 4); return f;
}
// Embree 4 TriangleM / Moeller-Trumbore (kernels/geometry/triangle_intersector_moeller.h, un-vendored)
inline bool intersectTriangle(const Ray &ray, V3 v0, V3 e1, V3 e2, float &t, float &u, float &v, V3 &Ng)
{
  const V3 triNg = crossF(e2, e1);
  const V3 C = v0 - ray.org;
  const V3 R = crossF(C, ray.dir);
  const float den = dotF(triNg, ray.dir);
  const float absDen = fabsf(den);
  uint32_t sgn; memcpy(&sgn, &den, 4); sgn &= 0x80000000u;
  const float U = xorSign(dotF(R, e2), sgn), V = xorSign(dotF(R, e1), sgn);
  if (!(den != 0.f && U >= 0.f && V >= 0.f && U + V <= absDen)) return false;
  const float T = xorSign(dotF(triNg, C), sgn);
  if (!(absDen * ray.t0 < T && T <= absDen * ray.t)) return false;
  const float rcpAbsDen = 1.f / absDen;
  u = U * rcpAbsDen; v = V * rcpAbsDen; t = T * rcpAbsDen; Ng = triNg;
  return true;
}
// Embree 4 sphere (kernels/geometry/sphere_intersector.h, un-vendored)
inline bool intersectSphere(const Ray &ray, V3 c, float r, float &t, V3 &Ng)
{
  const float rd2 = 1.f / dotF(ray.dir, ray.dir);
  const V3 c0 = c - ray.org;
  const float projC0 = dotF(c0, ray.dir) * rd2;
  const V3 perp = {fmaf(-projC0, ray.dir.x, c0.x), fmaf(-projC0, ray.dir.y, c0.y), fmaf(-projC0, ray.dir.z, c0.z)};
  const float l2 = dotF(perp, perp), r2 = r * r;
  if (!(l2 <= r2)) return false;
  const float td = sqrtf((r2 - l2) * rd2);
  const float tFront = projC0 - td, tBack = projC0 + td;
  float tdSel;
  if (ray.t0 <= tFront && tFront <= ray.t) { t = tFront; tdSel = -td; }
  else if (ray.t0 <= tBack && tBack <= ray.t) { t = tBack; tdSel = td; }
  else return false;
  Ng = {fmaf(tdSel, ray.dir.x, -perp.x), fmaf(tdSel, ray.dir.y, -perp.y), fmaf(tdSel, ray.dir.z, -perp.z)};
  return true;
}
inline float rcp_safe(float f) { return 1.f / (fabsf(f) < 1e-8f ? (f < 0.f ? -1e-8f : 1e-8f) : f); } // rkcommon
// common/Intersect.ih:194-221 + FilterIntersect.ih:41-116 (entry hit first, then exit hit)
inline bool intersectBox(const Ray &ray, V3 lo, V3 hi, float &t, V3 &Ng)
{
  const V3 rd = {rcp_safe(ray.dir.x), rcp_safe(ray.dir.y), rcp_safe(ray.dir.z)};
  const V3 mins = (lo - ray.org) * rd, maxs = (hi - ray.org) * rd;
  const V3 nears = vmin(mins, maxs), fars = vmax(mins, maxs);
  const float tEntry = rmax(nears), tExit = rmin(fars);
  if (!(tEntry < tExit)) return false;
  if (tEntry > ray.t0 && tEntry <= ray.t) {
    t = tEntry;
    if (tEntry == nears.x) Ng = v3(ray.dir.x > 0.f ? -1.f : 1.f, 0.f, 0.f);
    else if (tEntry == nears.y) Ng = v3(0.f, ray.dir.y > 0.f ? -1.f : 1.f, 0.f);
    else Ng = v3(0.f, 0.f, ray.dir.z > 0.f ? -1.f : 1.f);
    return true;
  }
  if (tExit > ray.t0 && tExit <= ray.t) {
    t = tExit;
    if (tExit == fars.x) Ng = v3(ray.dir.x > 0.f ? 1.f : -1.f, 0.f, 0.f);
    else if (tExit == fars.y) Ng = v3(0.f, ray.dir.y > 0.f ? 1.f : -1.f, 0.f);
    else Ng = v3(0.f, 0.f, ray.dir.z > 0.f ? 1.f : -1.f);
    return true;
  }
  return false;
}

// strided accessors (common/Data.ih)
inline V3 get3f(const OcData1D &d, uint32_t i) { const float *p = (const float *)(d.addr + (int64_t)i * d.byteStride); return {p[0], p[1], p[2]}; }
inline V4 get4f(const OcData1D &d, uint32_t i) { const float *p = (const float *)(d.addr + (int64_t)i * d.byteStride); return {p[0], p[1], p[2], p[3]}; }
inline uint32_t get1ui(const OcData1D &d, uint32_t i, int c) { return ((const uint32_t *)(d.addr + (int64_t)i * d.byteStride))[c]; }
inline V4 getColor(const OcGeometry *g, uint32_t i)
{
  if (g->flags & OC_GEOMFLAG_COLOR_IS_VEC4) return get4f(g->color, i);
  return v4(get3f(g->color, i), 1.f);
}

inline bool hitTri(Ray &ray, V3 a, V3 b, V3 c, bool second, uint32_t primID, uint32_t geomID, uint32_t instID, bool occl)
{
  float t, u, v; V3 Ng;
  if (!intersectTriangle(ray, a, a - b, c - a, t, u, v, Ng)) return false;
  ray.t = t;
  if (!occl) {
    ray.u = second ? 1.f - u : u; ray.v = second ? 1.f - v : v; ray.Ng = Ng;
    ray.primID = primID; ray.geomID = geomID; ray.instID = instID;
  }
  return true;
}

// all primitives of one group, in (model, prim) order; ray in instance space
bool traceGroup(const OcGroup *g, Ray &ray, uint32_t instID, bool occl)
{
  bool any = false;
  for (uint32_t m = 0; m < g->numModels; m++) {
    const OcGeometry *geo = g->models[m].geom;
    for (uint32_t p = 0; p < geo->numPrims; p++) {
      bool h = false;
      if (geo->type == OC_GEOM_TRIANGLES) {
        const V3 a = get3f(geo->vertex, get1ui(geo->index, p, 0)), b = get3f(geo->vertex, get1ui(geo->index, p, 1)),
                 c = get3f(geo->vertex, get1ui(geo->index, p, 2));
        h = hitTri(ray, a, b, c, false, p, m, instID, occl);
      } else if (geo->type == OC_GEOM_QUADS) {
        const V3 q0 = get3f(geo->vertex, get1ui(geo->index, p, 0)), q1 = get3f(geo->vertex, get1ui(geo->index, p, 1)),
                 q2 = get3f(geo->vertex, get1ui(geo->index, p, 2)), q3 = get3f(geo->vertex, get1ui(geo->index, p, 3));
        h = hitTri(ray, q0, q1, q3, false, p, m, instID, occl);
        if (h && occl) return true;
        h |= hitTri(ray, q2, q3, q1, true, p, m, instID, occl);
      } else if (geo->type == OC_GEOM_SPHERES) {
        const V4 s = get4f(geo->sphere, p);
        float t; V3 Ng;
        if (intersectSphere(ray, v3(s), s.w, t, Ng)) {
          h = true; ray.t = t;
          if (!occl) { ray.u = ray.v = 0.f; ray.Ng = Ng; ray.primID = p; ray.geomID = m; ray.instID = instID; }
        }
      } else if (geo->type == OC_GEOM_CURVES) { // round curve = 8 capsules per segment (curve_capsule.h)
        const uint32_t first = get1ui(geo->index, p, 0);
        curvecap::F4 cv[4];
        for (int k = 0; k < 4; k++) {
          const bool need = k < 2 || geo->curveBasis != OC_CURVE_BASIS_LINEAR;
          const V4 c = need ? get4f(geo->sphere, first + k) : V4{0, 0, 0, 0};
          cv[k] = {c.x, c.y, c.z, c.w};
        }
        const int basis = (int)geo->curveBasis;
        for (uint32_t k = 0; k < curvecap::CURVE_SUBDIV; k++) {
          const float u0 = (float)k * (1.f / curvecap::CURVE_SUBDIV), u1 = (float)(k + 1) * (1.f / curvecap::CURVE_SUBDIV);
          const curvecap::F4 a4 = curvecap::curveEval(basis, cv[0], cv[1], cv[2], cv[3], u0), b4 = curvecap::curveEval(basis, cv[0], cv[1], cv[2], cv[3], u1);
          const float rad = fmaxf(0.f, 0.5f * (a4.w + b4.w));
          float t, sPar; curvecap::F3 N;
          if (curvecap::intersectCapsule({ray.org.x, ray.org.y, ray.org.z}, {ray.dir.x, ray.dir.y, ray.dir.z}, ray.t0, ray.t, {a4.x, a4.y, a4.z},
                  {b4.x, b4.y, b4.z}, rad, t, sPar, N)) {
            h = true; ray.t = t;
            if (occl) break;
            ray.u = u0 + (1.f / curvecap::CURVE_SUBDIV) * sPar; ray.v = 0.f; ray.Ng = {N.x, N.y, N.z};
            ray.primID = p; ray.geomID = m; ray.instID = instID;
          }
        }
      } else {
        const float *b = (const float *)(geo->box.addr + (int64_t)p * geo->box.byteStride);
        float t; V3 Ng;
        if (intersectBox(ray, v3(b[0], b[1], b[2]), v3(b[3], b[4], b[5]), t, Ng)) {
          h = true; ray.t = t;
          if (!occl) { ray.u = ray.v = 0.f; ray.Ng = Ng; ray.primID = p; ray.geomID = m; ray.instID = instID; }
        }
      }
      if (h) { any = true; if (occl) return true; }
    }
  }
  return any;
}

// traceGeometryRay / isOccluded (common/World.ih:29-52, 311-335): two-level scene, instID = index
bool traceWorld(const OcWorld *w, Ray &ray, bool occl)
{
  bool any = false;
  for (uint32_t i = 0; i < w->numInstances; i++) {
    const OcInstance *inst = w->instances + i;
    if (inst->group->numModels == 0) continue;
    if (inst->isIdentity) {
      if (traceGroup(inst->group, ray, i, occl)) { any = true; if (occl) return true; }
    } else {
      const A3 rx = a3(inst->rcp_xfm);
      Ray local = ray;
      local.org = xfmPoint(rx, ray.org);
      local.dir = xfmVector(rx.l, ray.dir);
      if (traceGroup(inst->group, local, i, occl)) {
        any = true;
        ray.t = local.t;
        if (occl) return true;
        ray.u = local.u; ray.v = local.v; ray.Ng = local.Ng;
        ray.primID = local.primID; ray.geomID = local.geomID; ray.instID = local.instID;
      }
    }
  }
  return any;
}

struct Counters { uint64_t primary = 0, secondary = 0, shadow = 0; };

// ================================ DifferentialGeometry ===========================================
enum { DG_FACEFORWARD = 1, DG_NORMALIZE = 2, DG_NG = 4, DG_NS = 8, DG_COLOR = 16, DG_TEXCOORD = 32, DG_TANGENTS = 64, DG_MOTIONBLUR = 128 };
struct DG { V3 P, Ng, Ns, dPds; V4 color; float epsilon; uint32_t primID, objID, instID; const OcMaterial *material; };

inline float calcEpsilon(V3 P, float t) { return fmaxf(fmaxf(fabsf(P.x), fabsf(P.y)), fmaxf(fabsf(P.z), t)) * ULP_EPS; }
inline float calcEpsilon(V3 P, V3 dir, float t) { return calcEpsilon(P, rmax(vabs(dir)) * t); }

// geometry/Mesh.ispc:65-268, Spheres.ispc:24-42, Boxes.ispc:65-71
void geometryPostIntersect(const OcGeometry *g, DG &dg, const Ray &ray, uint32_t flags)
{
  dg.Ng = dg.Ns = ray.Ng;
  if (g->type == OC_GEOM_QUADS) {
    const uint32_t i0 = get1ui(g->index, ray.primID, 0), i1 = get1ui(g->index, ray.primID, 1), i2 = get1ui(g->index, ray.primID, 2),
                   i3 = get1ui(g->index, ray.primID, 3);
    const uint32_t fo = ray.primID * 4;
    const float u = ray.u, v = ray.v;
    V4 uv;
    if (i2 == i3) uv = {1.f - u - v, u, v, 0.f};
    else uv = {(1 - v) * (1 - u), (1 - v) * u, v * u, v * (1 - u)};
    if ((flags & DG_NS) && g->normal.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_NORMAL_FACEVARYING;
      dg.Ns = uv.x * get3f(g->normal, fv ? fo : i0) + uv.y * get3f(g->normal, fv ? fo + 1 : i1)
          + uv.z * get3f(g->normal, fv ? fo + 2 : i2) + uv.w * get3f(g->normal, fv ? fo + 3 : i3);
    }
    if ((flags & DG_COLOR) && g->color.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_COLOR_FACEVARYING;
      dg.color = uv.x * getColor(g, fv ? fo : i0) + uv.y * getColor(g, fv ? fo + 1 : i1) + uv.z * getColor(g, fv ? fo + 2 : i2)
          + uv.w * getColor(g, fv ? fo + 3 : i3);
      if (!(g->flags & OC_GEOMFLAG_HAS_ALPHA)) dg.color.w = 1.f;
    }
    const V3 va = get3f(g->vertex, i0), vb = get3f(g->vertex, i1), vc = get3f(g->vertex, i2), vd = get3f(g->vertex, i3);
    dg.epsilon = rmax(vmax(vmax(vabs(va), vabs(vb)), vmax(vabs(vc), vabs(vd)))) * ULP_EPS;
    if (flags & DG_TANGENTS) dg.dPds = vb - va; // no texcoords in scope (Mesh.ispc:170-173)
  } else if (g->type == OC_GEOM_TRIANGLES) {
    const uint32_t i0 = get1ui(g->index, ray.primID, 0), i1 = get1ui(g->index, ray.primID, 1), i2 = get1ui(g->index, ray.primID, 2);
    const uint32_t fo = ray.primID * 3;
    const V3 uv = {1.0f - ray.u - ray.v, ray.u, ray.v};
    if ((flags & DG_NS) && g->normal.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_NORMAL_FACEVARYING;
      dg.Ns = uv.x * get3f(g->normal, fv ? fo : i0) + uv.y * get3f(g->normal, fv ? fo + 1 : i1) + uv.z * get3f(g->normal, fv ? fo + 2 : i2);
    }
    if ((flags & DG_COLOR) && g->color.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_COLOR_FACEVARYING;
      dg.color = uv.x * getColor(g, fv ? fo : i0) + uv.y * getColor(g, fv ? fo + 1 : i1) + uv.z * getColor(g, fv ? fo + 2 : i2);
      if (!(g->flags & OC_GEOMFLAG_HAS_ALPHA)) dg.color.w = 1.f;
    }
    const V3 va = get3f(g->vertex, i0), vb = get3f(g->vertex, i1), vc = get3f(g->vertex, i2);
    dg.epsilon = rmax(vmax(vmax(vabs(va), vabs(vb)), vabs(vc))) * ULP_EPS;
    if (flags & DG_TANGENTS) dg.dPds = vb - va;
  } else if (g->type == OC_GEOM_SPHERES) {
    dg.epsilon = get4f(g->sphere, ray.primID).w * ULP_EPS;
  } else if (g->type == OC_GEOM_CURVES) { // Curves_postIntersect + interpolate<vec4f> (geometry/Curves.ispc:17-94)
    if ((flags & DG_COLOR) && g->color.addr) {
      const uint32_t index = get1ui(g->index, ray.primID, 0);
      const float t = ray.u;
      const V4 v0 = get4f(g->color, index + 0), v1 = get4f(g->color, index + 1);
      if (g->curveBasis == OC_CURVE_BASIS_LINEAR) dg.color = lerp(t, v0, v1);
      else {
        const V4 v2 = get4f(g->color, index + 2), v3 = get4f(g->color, index + 3);
        float bw[4], cf;
        curvecap::curveBasisWeights((int)g->curveBasis, t, bw, cf);
        dg.color = cf * (((bw[0] * v0 + bw[1] * v1) + bw[2] * v2) + bw[3] * v3);
      }
    }
  }
}

// geometry/GeometricModel.ih:21-47
const OcMaterial *getMaterial(const OcWorld *w, const OcRenderer &ren, const OcModel *m, uint32_t primID)
{
  if (m->materialMode == OC_MAT_NONE || m->numMaterials == 0) return nullptr;
  uint32_t idx = primID;
  if (m->index.addr) idx = *(m->index.addr + (int64_t)std::min(idx, m->index.numItems - 1) * m->index.byteStride);
  uint32_t matIdx = m->material[std::min(idx, m->numMaterials - 1)];
  if (m->materialMode == OC_MAT_FROM_RENDERER) {
    if (!ren.materials || ren.numMaterials == 0) return nullptr;
    matIdx = ren.materials[std::min(matIdx, ren.numMaterials - 1)];
  }
  return w->materials + matIdx;
}

// common/World.ih:389-465 + Instance.ih:44-105 + GeometricModel.ih:49-76
void postIntersect(const OcWorld *w, const OcRenderer &ren, DG &dg, const Ray &ray, uint32_t flags)
{
  dg.primID = ray.primID;
  dg.material = nullptr;
  if (flags & DG_COLOR) dg.color = {1.f, 1.f, 1.f, 1.f};
  if (flags & DG_TANGENTS) dg.dPds = v3(1.f, 0.f, 0.f);
  dg.P = ray.org + ray.t * ray.dir;
  dg.epsilon = 0.f;
  if (ray.instID != INVALID_ID) {
    const OcInstance *inst = w->instances + ray.instID;
    const OcModel *m = inst->group->models + ray.geomID;
    geometryPostIntersect(m->geom, dg, ray, flags);
    dg.objID = (m->userID == INVALID_ID) ? ray.geomID : m->userID;
    if ((flags & DG_COLOR) && m->color.addr) {
      uint32_t idx = ray.primID;
      if (m->index.addr) idx = *(m->index.addr + (int64_t)std::min(idx, m->index.numItems - 1) * m->index.byteStride);
      dg.color = get4f(m->color, std::min(idx, m->color.numItems - 1));
    }
    dg.material = getMaterial(w, ren, m, ray.primID);
    if ((flags & DG_NS) && m->invertedNormals) dg.Ns = -dg.Ns;
    dg.instID = (inst->userID == INVALID_ID) ? ray.instID : inst->userID;
    // the reference always applies the instance transform; identity is skipped here only because
    // x*1+y*0+z*0 is exact
    if (!inst->isIdentity) {
      const A3 xfm = a3(inst->xfm), rxfm = a3(inst->rcp_xfm);
      const L3 rt = transposed(rxfm.l);
      dg.Ns = xfmVector(rt, dg.Ns);
      dg.Ng = xfmVector(rt, dg.Ng);
      dg.epsilon *= fmaxf(fabsf(xfm.l.vx.x), fmaxf(fabsf(xfm.l.vy.y), fabsf(xfm.l.vz.z)));
      if (flags & DG_TANGENTS) dg.dPds = xfmVector(xfm.l, dg.dPds);
    }
  } else {
    dg.Ns = dg.Ng = ray.Ng;
    dg.objID = dg.instID = INVALID_ID;
  }
  dg.epsilon = fmaxf(dg.epsilon, calcEpsilon(dg.P, ray.dir, ray.t));
  V3 ffnng = normalize(dg.Ng);
  if ((flags & (DG_NG | DG_NORMALIZE)) == (DG_NG | DG_NORMALIZE)) dg.Ng = ffnng;
  if ((flags & (DG_NS | DG_NORMALIZE)) == (DG_NS | DG_NORMALIZE)) dg.Ns = normalize(dg.Ns);
  const bool flip = dot(ray.dir, dg.Ng) >= 0.f;
  if (flip) ffnng = -ffnng;
  if ((flags & (DG_NG | DG_FACEFORWARD)) == (DG_NG | DG_FACEFORWARD) && flip) dg.Ng = -dg.Ng;
  if ((flags & (DG_NS | DG_FACEFORWARD)) == (DG_NS | DG_FACEFORWARD))
    if (dot(dg.Ng, dg.Ns) < 0.f) dg.Ns = -dg.Ns;
  dg.P = dg.P + dg.epsilon * ffnng;
}

// ================================ sampling routines (math/sampling.ih) ============================
inline V3 cartesian(float phi, float sinTheta, float cosTheta) { return {cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta}; }
inline float cos2sin(float c) { return sqrtf(fmaxf(0.f, 1.f - c * c)); }
inline V3 cosineSampleHemisphere(V2 s) { return cartesian(TWO_PI * s.x, sqrtf(1.0f - s.y), sqrtf(s.y)); }
inline V3 powerCosineSampleHemisphere(float n, V2 s)
{
  const float cosTheta = powf(s.y, 1.0f / (n + 1.0f));
  return cartesian(TWO_PI * s.x, cos2sin(cosTheta), cosTheta);
}
inline float powerCosineSampleHemispherePDF(float cosTheta, float n) { return (n + 1.0f) * ONE_OVER_TWO_PI * powf(cosTheta, n); }
inline V3 uniformSampleCone(float cosAngle, V2 s)
{
  const float cosTheta = 1.0f - s.y * (1.0f - cosAngle);
  return cartesian(TWO_PI * s.x, cos2sin(cosTheta), cosTheta);
}
inline V3 uniformSampleSphere(float radius, V2 s)
{
  return cartesian(TWO_PI * s.x, 2.f * radius * sqrtf(s.y * (1.f - s.y)), radius * (1.f - 2.f * s.y));
}

// ================================ OBJ material (materials/OBJ.ispc:29-144) =========================
enum {
  SC_SPECULAR_REFLECTION = 1, SC_GLOSSY_REFLECTION = 2, SC_DIFFUSE_REFLECTION = 4, SC_SPECULAR_TRANSMISSION = 8,
  SC_GLOSSY_TRANSMISSION = 16, SC_DIFFUSE_TRANSMISSION = 32,
  SC_DIFFUSE = 4 | 32, SC_GLOSSY = 2 | 16, SC_SPECULAR = 1 | 8, SC_REFLECTION = 1 | 2 | 4, SC_TRANSMISSION = 8 | 16 | 32,
  SC_SMOOTH = SC_DIFFUSE | SC_GLOSSY
};
struct EvalRes { V3 value; float pdf; };
struct SampleRes { V3 weight, wi; float pdf; uint32_t type; };
struct ObjBSDF
{
  L3 frame; V3 Kd, T, ks, specWeight, albedo; float ns; float imp[3], importanceSum; uint32_t scatteringType; bool valid;
  bool thinGlass; // RobustThinDielectric (materials/ThinGlass.ispc): eta in `ns`, attenuation in `Kd`
};
// ---- thin glass: bsdfs/Fresnel.ih:19-47, Optics.ih:13-28, RobustThinDielectric.ih, materials/ThinGlass.ispc ----
inline float sqrCosT(float cosI, float eta) { return 1.0f - sqr(eta) * (1.0f - sqr(cosI)); }
inline float fresnelDielectric3(float cosI, float cosT, float eta)
{
  const float Rper = (eta * cosI - cosT) * (1.f / (eta * cosI + cosT));
  const float Rpar = (cosI - eta * cosT) * (1.f / (cosI + eta * cosT));
  return 0.5f * (sqr(Rpar) + sqr(Rper));
}
inline float fresnelDielectric(float cosI, float eta)
{
  const float sqrcost = sqrCosT(cosI, eta);
  if (sqrcost < 0.0f) return 1.0f;
  return fresnelDielectric3(cosI, sqrtf(sqrcost), eta);
}
inline float fresnelDielectricEx(float cosI, float &cosT, float eta)
{
  const float sqrcost = sqrCosT(cosI, eta);
  if (sqrcost < 0.0f) { cosT = 0.0f; return 1.0f; }
  cosT = sqrtf(sqrcost);
  return fresnelDielectric3(cosI, cosT, eta);
}
inline V3 expf3(V3 v) { return {expf(v.x), expf(v.y), expf(v.z)}; }
inline V3 logf3(V3 v) { return {logf(v.x), logf(v.y), logf(v.z)}; }
inline V3 rcp3(V3 v) { return {1.f / v.x, 1.f / v.y, 1.f / v.z}; }
inline V3 operator+(float a, V3 b) { return {a + b.x, a + b.y, a + b.z}; }
inline V3 operator-(float a, V3 b) { return {a - b.x, a - b.y, a - b.z}; }
inline V3 ThinGlass_getAttenuation(const OcMaterial *m, const DG &dg) // ThinGlass.ispc:17-30
{
  return v3(m->attenuation) + logf3(v3(dg.color)) * m->attenuationScale;
}
inline V3 RobustThinDielectric_getTransparency(float cosThetaO, float eta, V3 attenuation) // RobustThinDielectric.ih:99-117
{
  float cosThetaT;
  const float Fo = fresnelDielectricEx(cosThetaO, cosThetaT, eta);
  const float Ft = fresnelDielectric(cosThetaT, 1.f / eta);
  const float length = 1.f / cosThetaT;
  const V3 A = expf3(attenuation * length);
  return (1.f - Fo) * A * (1.f - Ft) * (1.f + (Ft * A) * (Ft * A) * rcp3(1.f - (Ft * A) * (Ft * A)));
}
inline L3 makeShadingFrame(const DG &dg) // render/Material.ih:25-37
{
  const V3 N = dg.Ns;
  V3 U = normalize(dg.dPds);
  V3 V = cross(N, U);
  if (dot(V, V) > 0.f) { V = normalize(V); U = cross(V, N); return {U, V, N}; }
  return frame(N);
}
void OBJ_getBSDF(ObjBSDF &b, const DG &dg)
{
  const OcMaterial *m = dg.material;
  b.valid = false;
  b.thinGlass = false;
  if (m && m->type == OC_MATERIAL_THINGLASS) { // ThinGlass_getBSDF (ThinGlass.ispc:32-50)
    b.valid = b.thinGlass = true;
    b.frame = frame(dg.Ns);
    b.Kd = ThinGlass_getAttenuation(m, dg);
    b.ns = m->eta;
    b.albedo = expf3(b.Kd);
    b.scatteringType = SC_SPECULAR;
    b.importanceSum = 1.f; b.imp[0] = b.imp[1] = b.imp[2] = 0.f;
    return;
  }
  if (!m || m->type != OC_MATERIAL_OBJ) return;
  b.valid = true;
  b.frame = makeShadingFrame(dg);
  b.albedo = v3(0.f); b.scatteringType = 0; b.importanceSum = 0.f; b.imp[0] = b.imp[1] = b.imp[2] = 0.f;
  const float d = m->d * dg.color.w;
  const V3 Kd = v3(m->Kd) * d * v3(dg.color);
  b.Kd = Kd;
  if (rmax(Kd) > 0.0f) { b.albedo = b.albedo + Kd; b.scatteringType |= SC_DIFFUSE_REFLECTION; b.imp[0] = luminance(Kd); b.importanceSum += b.imp[0]; }
  const V3 T = v3(m->Tf) * d + v3(1.f - d);
  b.T = T;
  if (rmax(T) > 0.0f) { b.albedo = b.albedo + T; b.scatteringType |= SC_SPECULAR_TRANSMISSION; b.imp[1] = luminance(T); b.importanceSum += b.imp[1]; }
  const V3 Ks = d * v3(m->Ks);
  if (rmax(Ks) > 0.0f) {
    b.albedo = b.albedo + Ks; b.scatteringType |= SC_GLOSSY_REFLECTION; b.imp[2] = luminance(Ks); b.importanceSum += b.imp[2];
    b.ns = m->Ns * 4.f; // bsdfs/Specular.ih:71-79
    const float tmp = (b.ns + 2.f) * (1.f / (2.f - expf(b.ns * -LNTWO_OVER_TWO)));
    b.ks = Ks * (ONE_OVER_FOUR_PI * tmp);
    b.specWeight = Ks * (2.f * tmp * (1.f / (b.ns + 1.f)));
  }
}
inline EvalRes Lambert_eval(const ObjBSDF &b, V3 wi) // bsdfs/Lambert.ih:11-20
{
  const float c = fmaxf(dot(wi, b.frame.vz), 0.f);
  return {b.Kd * ONE_OVER_PI * c, c * ONE_OVER_PI};
}
inline EvalRes Specular_eval(const ObjBSDF &b, V3 wo, V3 wi) // bsdfs/Specular.ih:22-40
{
  const V3 N = b.frame.vz;
  const float cosThetaO = dot(wo, N), cosThetaI = dot(wi, N);
  if (cosThetaO <= 0.0f || cosThetaI <= 0.0f) return {v3(0.f), 0.f};
  const V3 wh = normalize(wo + wi);
  const float cosThetaH = dot(wh, N);
  EvalRes r;
  r.pdf = powerCosineSampleHemispherePDF(cosThetaH, b.ns) * (1.f / (4.0f * dot(wo, wh)));
  r.value = b.ks * powf(cosThetaH, b.ns);
  return r;
}
inline void evalChildren(const ObjBSDF &b, V3 wo, V3 wi, int choice, EvalRes &e) // MultiBSDF.ih:70-78
{
  if (b.imp[0] > 0.f && choice != 0) { const EvalRes c = Lambert_eval(b, wi); e.value = e.value + c.value; e.pdf += c.pdf * b.imp[0]; }
  if (b.imp[2] > 0.f && choice != 2) { const EvalRes c = Specular_eval(b, wo, wi); e.value = e.value + c.value; e.pdf += c.pdf * b.imp[2]; }
}
inline EvalRes OBJ_eval(const ObjBSDF &b, V3 wo, V3 wi)
{
  EvalRes e = {v3(0.f), 0.f};
  if (b.thinGlass) return e; // purely specular
  evalChildren(b, wo, wi, -1, e);
  e.pdf *= 1.f / b.importanceSum;
  return e;
}
SampleRes RobustThinDielectric_sample(const ObjBSDF &b, V3 wo, float ss) // RobustThinDielectric.ih:17-62
{
  SampleRes res;
  const V3 N = b.frame.vz;
  const float eta = b.ns;
  const float cosThetaO = fmaxf(dot(wo, N), 0.f);
  float cosThetaT;
  const float Fo = fresnelDielectricEx(cosThetaO, cosThetaT, eta);
  const float Ft = fresnelDielectric(cosThetaT, 1.f / eta);
  const float length = 1.f / cosThetaT;
  const V3 A = expf3(b.Kd * length);
  const V3 R = Fo + (1.f - Fo) * (1.f - Ft) * (Ft * (A * A)) * rcp3(1.f - (Ft * A) * (Ft * A));
  const V3 T = (1.f - Fo) * A * (1.f - Ft) * (1.f + (Ft * A) * (Ft * A) * rcp3(1.f - (Ft * A) * (Ft * A)));
  const float pR = luminance(R), pT = luminance(T);
  const float p = fmaxf(0.25f, fminf(pR * (1.f / (pR + pT)), 0.75f));
  if (ss <= p) { res.wi = (2.0f * cosThetaO) * N - wo; res.type = SC_SPECULAR_REFLECTION; res.weight = R * (1.f / p); }
  else { res.wi = -wo; res.type = SC_SPECULAR_TRANSMISSION; res.weight = T * (1.f / (1.f - p)); }
  res.pdf = INF;
  return res;
}
SampleRes OBJ_sample(const ObjBSDF &b, V3 wo, V2 s, float ss) // MultiBSDF.ih:83-128
{
  if (b.thinGlass) return RobustThinDielectric_sample(b, wo, ss);
  SampleRes res = {v3(0.f), v3(0.f), 0.f, 0};
  if (b.importanceSum == 0.f) return res;
  const float x = ss * b.importanceSum;
  int choice = 0;
  float prefix = 0.f;
  if (x < (prefix += b.imp[0])) { // Lambert_sample (Lambert.ih:22-31)
    choice = 0;
    const V3 localDir = cosineSampleHemisphere(s);
    res.wi = xfmVector(b.frame, localDir); res.pdf = localDir.z * ONE_OVER_PI; res.type = SC_DIFFUSE_REFLECTION; res.weight = b.Kd;
  } else if (x < (prefix += b.imp[1])) { // Transmission_sample (Transmission.ih:11-21)
    choice = 1;
    res.wi = -wo; res.pdf = INF; res.type = SC_SPECULAR_TRANSMISSION; res.weight = b.T;
  } else if (x < (prefix += b.imp[2])) { // Specular_sample (Specular.ih:42-66)
    choice = 2;
    const V3 whl = powerCosineSampleHemisphere(b.ns, s);
    const V3 wh = xfmVector(b.frame, whl);
    const float cosThetaOH = dot(wo, wh);
    res.wi = 2.f * cosThetaOH * wh - wo;
    if (dot(res.wi, b.frame.vz) <= 0.0f) res = {v3(0.f), v3(0.f), 0.f, 0};
    else {
      res.pdf = powerCosineSampleHemispherePDF(whl.z, b.ns) * (1.f / (4.0f * cosThetaOH));
      res.type = SC_GLOSSY_REFLECTION;
      res.weight = b.specWeight * cosThetaOH;
    }
  }
  if (eq(res.weight, v3(0.f)) || res.pdf == 0.0f) return {v3(0.f), v3(0.f), 0.f, 0};
  if (res.type & SC_SPECULAR) { res.weight = res.weight * (b.importanceSum * (1.f / b.imp[choice])); return res; }
  EvalRes e;
  e.value = res.weight * res.pdf;
  res.pdf *= b.imp[choice];
  e.pdf = res.pdf;
  evalChildren(b, wo, res.wi, choice, e);
  res.pdf = e.pdf;
  res.pdf *= 1.f / b.importanceSum;
  res.weight = e.value * (1.f / res.pdf);
  return res;
}
inline V3 OBJ_getTransparency(const DG &dg, V3 rayDir) // OBJ.ispc:93-111, ThinGlass.ispc:52-64; NULL material -> 0
{
  const OcMaterial *m = dg.material;
  if (m && m->type == OC_MATERIAL_THINGLASS)
    return RobustThinDielectric_getTransparency(fmaxf(-dot(rayDir, dg.Ns), 0.0f), m->eta, ThinGlass_getAttenuation(m, dg));
  if (!m || m->type != OC_MATERIAL_OBJ) return v3(0.f);
  const float d = m->d * dg.color.w;
  return v3(m->Tf) * d + v3(1.f - d);
}

// ================================ lights ==========================================================
struct LightSample { V3 weight, dir; float dist, pdf; };
struct LightEval { V3 radiance; float pdf; };
struct SphericalQuad { V3 e0, e1, z; float S, k, n0z, n2z, x0, x1, y0, y1, z0; };
SphericalQuad SphericalQuad_create(V3 q0, V3 e0, V3 e1, V3 n, V3 P) // lights/SphericalQuadSampling.ih:16-80
{
  SphericalQuad q;
  const float e0l = length(e0), e1l = length(e1);
  q.e0 = e0 / e0l; q.e1 = e1 / e1l; q.z = n;
  const V3 d = q0 - P;
  q.z0 = dot(d, q.z);
  q.z = q.z0 > 0.f ? -q.z : q.z;
  q.z0 = -fabsf(q.z0);
  q.x0 = dot(d, q.e0); q.y0 = dot(d, q.e1);
  q.x1 = q.x0 + e0l; q.y1 = q.y0 + e1l;
  const V3 v00 = {q.x0, q.y0, q.z0}, v01 = {q.x0, q.y1, q.z0}, v10 = {q.x1, q.y0, q.z0}, v11 = {q.x1, q.y1, q.z0};
  const V3 n0 = normalize(cross(v00, v10)), n1 = normalize(cross(v10, v11)), n2 = normalize(cross(v11, v01)),
           n3 = normalize(cross(v01, v00));
  q.n0z = n0.z; q.n2z = n2.z;
  const float g0 = acosf(clampf(-dot(n0, n1), -1.f, 1.f)), g1 = acosf(clampf(-dot(n1, n2), -1.f, 1.f)),
              g2 = acosf(clampf(-dot(n2, n3), -1.f, 1.f)), g3 = acosf(clampf(-dot(n3, n0), -1.f, 1.f));
  q.k = TWO_PI - g2 - g3;
  q.S = fmaxf(0.f, g0 + g1 - q.k);
  return q;
}
inline float signf(float x) { return (float)(0.f < x) - (float)(x < 0.f); }
inline float rsqrt_(float x) { return 1.f / sqrtf(x); }
V3 sampleSphericalQuad(const SphericalQuad &q, V2 s) // SphericalQuadSampling.ih:87-110
{
  const float au = s.x * q.S + q.k;
  const float sinau = sinf(au), cosau = cosf(au);
  const float fu = (cosau * q.n0z - q.n2z) * (1.f / sinau);
  const float cu = clampf(rsqrt_(fu * fu + sqr(q.n0z)) * signf(fu), -1.f, 1.f);
  const float x0 = -(cu * q.z0) * rsqrt_(1.f - cu * cu);
  const float xu = clampf(x0, q.x0, q.x1);
  const float ds = sqrtf(sqr(xu) + sqr(q.z0));
  const float ds2 = sqr(ds);
  const float h0 = q.y0 * rsqrt_(ds2 + sqr(q.y0));
  const float h1 = q.y1 * rsqrt_(ds2 + sqr(q.y1));
  const float hv = s.y * (h1 - h0) + h0;
  const float hv2 = sqr(hv);
  const float yv = (hv2 < 1.f - 1e-4f) ? (hv * ds) * rsqrt_(1.f - hv2) : q.y1;
  return xu * q.e0 + yv * q.e1 + q.z0 * q.z;
}
float solve_quadratic(float A, float B, float C) // :112-133
{
  if (A == 0.f) return -C / B;
  float rad = B * B - 4.f * A * C;
  if (rad < 0.f) return 0.f;
  rad = sqrtf(rad);
  float root;
  if (B >= 0.f) { root = (-B - rad) / (2.f * A); if (root < 0.f || root > 1.f) root = 2.f * C / (-B - rad); }
  else { root = 2.f * C / (-B + rad); if (root < 0.f || root > 1.f) root = (-B + rad) / (2.f * A); }
  return root;
}
V3 computeCosineWeightedRNG(V3 q0, V3 e0, V3 e1, V3 P, V3 Ng, V2 s) // :138-198
{
  V3 wp[4]; float prob[4];
  wp[0] = q0; wp[1] = q0 + e0; wp[2] = q0 + e1; wp[3] = wp[1] + e1;
  for (int i = 0; i < 4; i++) prob[i] = fmaxf(0.f, dot(normalize(wp[i] - P), Ng));
  if ((prob[0] + prob[2]) == 0.f) return {s.x, s.y, 1.f};
  const float totProb = prob[0] + prob[2] + prob[1] + prob[3];
  const float W00 = prob[0], W01 = prob[2], W10 = prob[1], W11 = prob[3];
  float u, v;
  { const float A = W10 + W11 - W01 - W00, B = 2.f * (W01 + W00);
    u = (A + B == 0.f) ? s.x : solve_quadratic(A, B, -(W00 + W01 + W10 + W11) * s.x); }
  { const float A = (1.f - u) * W01 + u * W11 - (1.f - u) * W00 - u * W10, B = 2.f * (1.f - u) * W00 + 2.f * u * W10;
    v = (A + B == 0.f) ? s.y : solve_quadratic(A, B, -((1.f - u) * W01 + u * W11 + (1.f - u) * W00 + u * W10) * s.y); }
  const float pdf = 4.f * ((1.f - u) * (1.f - v) * prob[0] + u * (1.f - v) * prob[1] + u * v * prob[3] + (1.f - u) * v * prob[2]) / totProb;
  return {u, v, pdf};
}
LightSample QuadLight_sampleArea(const OcLight &l, V3 P, V2 s) // lights/QuadLight.ispc:43-80
{
  LightSample r;
  const V3 p = v3(l.position) + v3(l.edge1) * s.x + v3(l.edge2) * s.y;
  const V3 dir = p - P;
  const float dist = length(dir);
  r.dir = dir / dist; r.dist = dist;
  const float cosd = dot(v3(l.nnormal), r.dir);
  r.pdf = l.ppdf * sqr(dist) / fabsf(cosd);
  r.weight = v3(l.radiance) * (cosd > 0.f ? 1.f / r.pdf : 0.f);
  return r;
}
// PointLight ("sphere" light): lights/PointLight.ispc:38-92 (Sample), :131-160 (Eval); no intensity distribution
LightSample PointLight_sample(const OcLight &l, const DG &dg, V2 s)
{
  LightSample res;
  const V3 dir = v3(l.position) - dg.P;
  const float dist2 = dot(dir, dir);
  const float invdist = 1.f / sqrtf(dist2);
  res.dir = dir * invdist;
  res.dist = dist2 * invdist;
  res.pdf = INF;
  const float sinTheta = l.radius * invdist;
  if ((l.radius > 0.f) & (sinTheta > 0.005f)) {
    if (sinTheta < 1.f) {
      const float cosTheta = sqrtf(1.f - sinTheta * sinTheta);
      const V3 localDir = uniformSampleCone(cosTheta, s);
      res.dir = xfmVector(frame(res.dir), localDir);
      res.pdf = 1.f / (TWO_PI * (1.0f - cosTheta));
      const float c = localDir.z;
      const float disc = sqr(l.radius) - (1.f - c * c) * dist2;
      res.dist = c * res.dist - sqrtf(disc > 0.f ? disc : 0.f);
      res.weight = v3(l.radiance) / res.pdf;
    } else
      res.weight = v3(0.f);
  } else
    res.weight = v3(l.intensity) * sqr(invdist);
  return res;
}
LightEval PointLight_eval(const OcLight &l, V3 P, V3 dir, float minDist, float maxDist)
{
  LightEval res = {v3(0.f), 0.f};
  const V3 A = v3(l.position) - P;
  const float centerDist2 = dot(A, A);
  const float sinTheta2 = sqr(l.radius) * (1.f / centerDist2);
  if (l.radius > 0.f && sqrtf(sinTheta2) > 0.005f) {
    // intersectSphere (common/Intersect.ih:29-78), entry hit only
    const float rd2 = 1.0f / dot(dir, dir);
    const float projCO = dot(A, dir) * rd2;
    const V3 perp = A - projCO * dir;
    const float l2 = dot(perp, perp), r2 = sqr(l.radius);
    if (!(l2 > r2)) {
      const float t = projCO - sqrtf((r2 - l2) * rd2);
      if (t > minDist && t < maxDist) {
        const float cosTheta = sqrtf(1.f - sinTheta2);
        res.pdf = 1.f / (TWO_PI * (1.0f - cosTheta));
        res.radiance = v3(l.radiance);
      }
    }
  }
  return res;
}
// SpotLight: lights/SpotLight.ispc:26-86 (AngularAttenuation, Sample), :124-164 (Eval); no intensity distribution
inline float SpotLight_attenuation(const OcLight &l, float cosAngle) { return fmaxf(0.f, fminf((cosAngle - l.cosAngle) * l.cosAngleScale, 1.f)); }
LightSample SpotLight_sample(const OcLight &l, const DG &dg, V2 s)
{
  LightSample res;
  const V3 direction = v3(l.fz);
  res.dir = v3(l.position) - dg.P;
  if (l.radius > 0.0f) {
    const float r = sqrtf(lerp(s.x, sqr(l.innerRadius), sqr(l.radius))); // uniformSampleRing (math/sampling.ih:114-122)
    const float phi = TWO_PI * s.y;
    res.dir = xfmVector(L3{v3(l.fx), v3(l.fy), direction}, v3(r * cosf(phi), r * sinf(phi), 0.f)) + res.dir;
  }
  const float dist2 = dot(res.dir, res.dir);
  const float invdist = 1.f / sqrtf(dist2);
  res.dir = res.dir * invdist;
  res.dist = dist2 * invdist;
  const float cosAngle = -dot(direction, res.dir);
  float weight = SpotLight_attenuation(l, cosAngle);
  if (l.radius > 0.0f) {
    res.pdf = l.ppdf * (dist2 / fabsf(cosAngle));
    weight /= res.pdf;
    res.weight = v3(l.radiance) * weight;
  } else {
    res.pdf = INF;
    weight *= fabsf(cosAngle);
    res.weight = v3(l.intensity) * (sqr(invdist) * weight);
  }
  return res;
}
LightEval SpotLight_eval(const OcLight &l, V3 P, V3 dir, float minDist, float maxDist)
{
  LightEval res = {v3(0.f), 0.f};
  if (l.radius > 0.f) {
    const V3 direction = v3(l.fz);
    const float cosAngle = -dot(direction, dir);
    if (cosAngle > l.cosAngle) {
      const V3 vp = P - v3(l.position);
      const float t = dot(vp, direction) * (1.f / cosAngle);
      if ((t > minDist) & (t <= maxDist)) {
        const V3 vd = vp + t * dir;
        const float d2 = dot(vd, vd);
        if (d2 < sqr(l.radius) && d2 > sqr(l.innerRadius)) {
          res.radiance = v3(l.radiance) * SpotLight_attenuation(l, cosAngle);
          res.pdf = l.ppdf * (sqr(t) / fabsf(cosAngle));
        }
      }
    }
  }
  return res;
}
LightSample Light_sample(const OcLight &l, const DG &dg, V2 s)
{
  LightSample r;
  if (l.type == OC_LIGHT_SPHERE) return PointLight_sample(l, dg, s);
  if (l.type == OC_LIGHT_SPOT) return SpotLight_sample(l, dg, s);
  if (l.type == OC_LIGHT_QUAD) { // QuadLight.ispc:82-135
    if (s.x == 0.0f && s.y == 0.0f) return QuadLight_sampleArea(l, dg.P, {0.5f, 0.5f});
    const V3 pos = v3(l.position), e1 = v3(l.edge1), e2 = v3(l.edge2), nn = v3(l.nnormal);
    const SphericalQuad quad = SphericalQuad_create(pos, e1, e2, -nn, dg.P);
    const V3 cosW = computeCosineWeightedRNG(pos, e1, e2, dg.P, dg.Ng, s);
    const V3 dir = sampleSphericalQuad(quad, {cosW.x, cosW.y});
    const float dist = length(dir);
    r.dir = dir / dist; r.dist = dist;
    const float cosd = dot(nn, r.dir);
    r.pdf = quad.S == 0.f ? 0.f : cosW.z / quad.S;
    r.weight = (cosd > 0.f) && (r.pdf != 0.f) ? v3(l.radiance) * (1.f / r.pdf) : v3(0.f);
    return r;
  }
  if (l.type == OC_LIGHT_AMBIENT) { // AmbientLight.ispc:17-37
    if (rmax(vabs(dg.Ns)) > 0) {
      const V3 localDir = cosineSampleHemisphere(s);
      r.dir = xfmVector(frame(dg.Ns), localDir);
      r.pdf = localDir.z * ONE_OVER_PI;
    } else { r.dir = uniformSampleSphere(1.f, s); r.pdf = 1.f / FOUR_PI; }
    r.dist = INF;
    r.weight = v3(l.radiance) * (1.f / r.pdf);
    return r;
  }
  // DirectionalLight.ispc:14-31
  r.dir = v3(l.fz); r.dist = INF; r.pdf = l.pdf;
  if (l.cosAngle < 0.99999988f) r.dir = xfmVector(L3{v3(l.fx), v3(l.fy), v3(l.fz)}, uniformSampleCone(l.cosAngle, s));
  r.weight = v3(l.radiance);
  return r;
}
LightEval Light_eval(const OcLight &l, V3 P, V3 Ns, V3 dir, float minDist, float maxDist)
{
  LightEval r = {v3(0.f), 0.f};
  if (l.type == OC_LIGHT_SPHERE) return PointLight_eval(l, P, dir, minDist, maxDist);
  if (l.type == OC_LIGHT_SPOT) return SpotLight_eval(l, P, dir, minDist, maxDist);
  if (l.type == OC_LIGHT_QUAD) { // QuadLight.ispc:169-215
    const V3 nn = v3(l.nnormal), pos = v3(l.position), e1 = v3(l.edge1), e2 = v3(l.edge2);
    const float cosd = dot(nn, dir);
    if (cosd <= 0.f) return r;
    const V3 c = pos - P;
    const V3 rr = cross(c, dir);
    const float rcosd = 1.f / cosd;
    const float u = dot(rr, e1) * rcosd, v = -dot(rr, e2) * rcosd;
    if (fminf(u, v) < 0.f || fmaxf(u, v) * l.ppdf > 1.0f) return r;
    const float dist = dot(nn, c) * rcosd;
    if (dist <= minDist || dist > maxDist) return r;
    const SphericalQuad quad = SphericalQuad_create(pos, e1, e2, -nn, P);
    if (quad.S > 0.f) { r.radiance = v3(l.radiance); r.pdf = 1.f / quad.S; }
    return r;
  }
  if (l.type == OC_LIGHT_AMBIENT) { // AmbientLight.ispc:39-57
    r.radiance = INF <= maxDist ? v3(l.radiance) : v3(0.f);
    r.pdf = rmax(vabs(Ns)) > 0 ? fmaxf(dot(Ns, dir), 0.f) * ONE_OVER_PI : 1.f / FOUR_PI;
    return r;
  }
  if (INF <= maxDist && l.cosAngle < 0.99999988f && dot(v3(l.fz), dir) > l.cosAngle) { // DirectionalLight.ispc:64-80
    r.radiance = v3(l.radiance) * l.pdf; r.pdf = l.pdf;
  }
  return r;
}

// math/Distribution1D.ih:26-57
struct Sample1D { int idx; float frac, pdf, prob; };
Sample1D Distribution1D_sample(int size, const float *cdf, int start, float s)
{
  int first = start, len = size;
  while (len > 0) {
    const int half = len >> 1, mid = first + half;
    if (s < cdf[mid]) len = half;
    else { first = mid + 1; len -= half + 1; }
  }
  Sample1D r;
  r.idx = first - start;
  const float bef = first == start ? 0.0f : cdf[first - 1];
  const float dpdf = cdf[first] - bef;
  r.prob = dpdf; r.pdf = dpdf * size; r.frac = (s - bef) * (1.f / dpdf);
  return r;
}
float Distribution1D_create(int size, float *cdf) // math/Distribution1D.ispc:14-41
{
  float sum = cdf[0];
  for (int i = 1; i < size; i++) cdf[i] = sum += cdf[i];
  const float rcpSum = 1.0f / sum;
  for (int i = 0; i < size; i++) { const float c = cdf[i]; cdf[i] = c >= sum ? 0x1.000002p+0f : c * rcpSum; }
  return sum;
}
// pf/PixelFilter.ih:20-28, pf/LUTPixelFilter.ispc:9-15, math/Distribution2D.ispc:9-25
V2 pixelFilterSample(const OcRenderer &r, V2 s)
{
  if (r.pfType == OC_PF_POINT) return {0.f, 0.f};
  if (r.pfType == OC_PF_BOX) return (s + V2{-0.5f, -0.5f}) * r.pfWidth;
  const Sample1D sy = Distribution1D_sample((int)r.pfLutH, r.pfCdfY, 0, s.y);
  const Sample1D sx = Distribution1D_sample((int)r.pfLutW, r.pfCdfX, sy.idx * (int)r.pfLutW, s.x);
  const V2 uv = {(sx.idx + sx.frac) * (1.f / (float)r.pfLutW), (sy.idx + sy.frac) * (1.f / (float)r.pfLutH)};
  return (uv + V2{-0.5f, -0.5f}) * r.pfWidth;
}
// camera/PerspectiveCamera.ispc:16-64 (no stereo / motion blur)
void cameraInitRay(const OcCamera &cam, Ray &ray, V2 screen, V2 lens, V2 pixelCenter)
{
  if (cam.type == OC_CAMERA_PANORAMIC) { // PanoramicCamera_initRay (camera/PanoramicCamera.ispc:14-92)
    float offset = 0.f;
    const float ipd = cam.ipdOffset.x;
    if (cam.stereoMode == 1u) offset = -ipd;
    else if (cam.stereoMode == 2u) offset = ipd;
    else if (cam.stereoMode >= 3u) {
      const bool sbs = cam.stereoMode == 3u;
      float &split = sbs ? screen.x : screen.y;
      const float pc = sbs ? pixelCenter.x : pixelCenter.y;
      split *= 2.f;
      if (pc < 0.5f) offset = -ipd;
      else { offset = ipd; split -= 1.f; }
    }
    screen = {lerp(screen.x, cam.subImage[0], cam.subImage[2]), lerp(screen.y, cam.subImage[1], cam.subImage[3])};
    const float phi = TWO_PI * screen.x, theta = (float)M_PI * screen.y;
    const float sinTheta = sinf(theta), cosTheta = cosf(theta);
    const V3 localDir = cartesian(phi, sinTheta, cosTheta);
    const V3 vx = v3(cam.du_size), vy = v3(cam.dv_up), vz = v3(cam.dir_00);
    const V3 dir = (-localDir.y) * vx + (-localDir.z) * vy + localDir.x * vz;
    setRay(ray, v3(cam.org) + offset * cross(dir, vy), dir, cam.nearClip, INF);
    return;
  }
  V3 eyeOffset = v3(0.f);
  if (cam.type != OC_CAMERA_ORTHOGRAPHIC && cam.stereoMode >= 3u) { // side by side / top-bottom stereo (:28-40)
    const bool sbs = cam.stereoMode == 3u;
    float &split = sbs ? screen.x : screen.y;
    const float pc = sbs ? pixelCenter.x : pixelCenter.y;
    split *= 2.f;
    if (pc < 0.5f) eyeOffset = -1.f * v3(cam.ipdOffset);
    else { eyeOffset = v3(cam.ipdOffset); split -= 1.f; }
  }
  screen = {lerp(screen.x, cam.subImage[0], cam.subImage[2]), lerp(screen.y, cam.subImage[1], cam.subImage[3])};
  if (cam.type == OC_CAMERA_ORTHOGRAPHIC) { // OrthographicCamera_initRay (camera/OrthographicCamera.ispc:12-50)
    setRay(ray, v3(cam.org) + screen.x * v3(cam.du_size) + screen.y * v3(cam.dv_up), v3(cam.dir_00), cam.nearClip, INF);
    return;
  }
  V3 org = v3(cam.org) + eyeOffset;
  V3 dir = v3(cam.dir_00) + screen.x * v3(cam.du_size) + screen.y * v3(cam.dv_up);
  if (cam.scaledAperture > 0.f) { // uniformSampleDisk(radius, s)
    const float r = sqrtf(lens.x) * cam.scaledAperture, phi = TWO_PI * lens.y;
    const V3 lp = ((r * cosf(phi)) * v3(cam.du_size)) + (((r * sinf(phi)) * cam.aspect) * v3(cam.dv_up));
    org = org + lp; dir = dir - lp;
  }
  setRay(ray, org, normalize(dir), cam.nearClip, INF);
}

// ================================ renderers =======================================================
struct ScreenSample { V3 rgb; float alpha, z; V3 position, firstNormal, normal, albedo; uint32_t primID, geomID, instID; };
inline ScreenSample make_ScreenSample_zero() // render/ScreenSample.ih:39-56
{
  ScreenSample s;
  s.rgb = v3(0.f); s.alpha = 0.f; s.z = INF; s.position = s.firstNormal = s.normal = s.albedo = v3(0.f);
  s.primID = s.geomID = s.instID = INVALID_ID;
  return s;
}
inline void ScreenSample_accumulate(ScreenSample &a, const ScreenSample &s) // :58-76
{
  a.rgb = a.rgb + s.rgb; a.alpha = a.alpha + s.alpha; a.position = a.position + s.position;
  a.firstNormal = a.firstNormal + s.firstNormal; a.normal = a.normal + s.normal; a.albedo = a.albedo + s.albedo;
  if (fabsf(s.z) < fabsf(a.z)) { a.z = s.z; a.primID = s.primID; a.geomID = s.geomID; a.instID = s.instID; }
}
inline void ScreenSample_normalize(ScreenSample &s, float rspp) // :78-86
{
  s.rgb = s.rgb * rspp; s.alpha *= rspp; s.position = s.position * rspp; s.firstNormal = s.firstNormal * rspp;
  s.normal = s.normal * rspp; s.albedo = s.albedo * rspp;
}

struct Ctx { const OcWorld *w; const OcRenderer &ren; const OcCamera &cam; const OcFrameBuffer &fb; Counters cnt; };
inline V4 getBackground(const OcRenderer &r) { return {r.bgColor[0], r.bgColor[1], r.bgColor[2], r.bgColor[3]}; }

// ---- path tracer -------------------------------------------------------------------------------
inline float misHeuristic(float pdf1, float pdf2) // pathtracer/PathTracerUtil.ih:30-46
{
  const float p = sqr(pdf1) * (1.f / (sqr(pdf1) + sqr(pdf2)));
  return pdf1 > 1e17f ? 1.0f : p;
}
// ---- path tracer volumes: pathtracer/volumes/VolumeSampler.ispc:13-282 in its single-volume-interval form (the form the
// reference compiles for OSPRAY_TARGET_SYCL), pathtracer/volumes/HenyeyGreenstein.ih:11-74 ----
struct VolumeHit;
float volumeSample(const OcVolume &v, V3 p);
V4 transferFunctionGet(const OcTransferFunction &tf, float value);
struct VolumeHitFwd { const OcVolume *vol; const OcInstance *inst; uint32_t instID, geomID; float t0, t1; };
bool traceVolumeRayFwd(const OcWorld *w, const Ray &ray, VolumeHitFwd &vh);
// LinearTransferFunction_getMaxOpacity (LinearTransferFunction.ispc:46-64) over the table precomputeMaxOpacityRanges
// fills (LinearTransferFunction.cpp:110-160): entry [i][j] = max of the opacities floor(D*i/31) .. ceil(D*j/31), D = n - 1
float tfMaxOpacity(const OcTransferFunction &tf, float lo, float hi)
{
  const int maxDim = 31, D = (int)tf.numOpacities - 1;
  const float denom = tf.valueUpper - tf.valueLower;
  const int i = std::min((int)floorf(clampf((lo - tf.valueLower) / denom, 0.f, 1.f) * maxDim), maxDim);
  const int j = std::min((int)ceilf(clampf((hi - tf.valueLower) / denom, 0.f, 1.f) * maxDim), maxDim);
  if (j < i) return 0.f; // below the diagonal the table keeps its initial zeros
  const int lowI = (int)floorf(D * (float)i / maxDim), highI = (int)ceilf(D * (float)i / maxDim), highJ = (int)ceilf(D * (float)j / maxDim);
  float m = tf.opacity[lowI];
  for (int k = lowI; k <= std::max(highI, highJ); k++) m = fmaxf(m, tf.opacity[k]);
  return m;
}
// delta_tracking (VolumeSampler.ispc:13-103) over the ONE interval the structured-volume stand-in of Open VKL reports
// (whole clipped range, value range (-inf, inf): oracle/ref_vkl_stub.cpp) -> global majorant
float deltaTracking(const OcVolume &v, float lower, float upper, RandomSampler &rs, V3 o, V3 w, float &mu_t, V3 &albedo)
{
  if (!(lower < upper)) return INF; // the iterator reports nothing for an empty range
  float t = lower;
  const float mu_max = v.densityScale * tfMaxOpacity(v.tf, -INF, INF);
  if (mu_max <= 0.f) return INF;
  while (true) {
    float xi = RandomSampler_getFloat(rs);
    const float dt = -logf(1.f - xi) / mu_max;
    t += dt;
    if (t > upper) break;
    xi = RandomSampler_getFloat(rs);
    const float sample = volumeSample(v, o + t * w);
    if (std::isnan(sample)) continue;
    const V4 color = transferFunctionGet(v.tf, sample);
    mu_t = v.densityScale * color.w;
    if (xi < mu_t / mu_max) { albedo = v3(color); return t; }
  }
  return INF;
}
// volumeSampleFreePath (VolumeSampler.ispc:105-196); rayInterval = the interval traceClippingRay hands out without
// clipping geometry (common/World.ih:150-163)
float volumeSampleFreePath(Ctx &c, Ray &ray, float rayLower, float rayUpper, RandomSampler &rs, const OcVolume *&sampled, V3 &sampledAlbedo)
{
  sampled = nullptr;
  VolumeHitFwd vh;
  if (!traceVolumeRayFwd(c.w, ray, vh)) return INF;
  float lower = fmaxf(0.f, vh.t0), upper = hadHit(ray) ? fminf(vh.t1, ray.t) : vh.t1;
  if (lower > upper) return INF;
  V3 org = ray.org, dir = ray.dir;
  if (!vh.inst->isIdentity) { const A3 rx = a3(vh.inst->rcp_xfm); org = xfmPoint(rx, ray.org); dir = xfmVector(rx.l, ray.dir); }
  if (upper < rayLower || rayUpper < lower) return INF;
  lower = fmaxf(rayLower, lower); upper = fminf(rayUpper, upper);
  float extinction; V3 albedo;
  const float dist = deltaTracking(*vh.vol, lower, upper, rs, org, dir, extinction, albedo);
  if (dist < INF) {
    sampled = vh.vol; sampledAlbedo = albedo;
    ray.t = dist; ray.primID = 0; ray.geomID = vh.geomID; ray.instID = vh.instID;
  }
  return dist;
}
// volumeTransmittance (VolumeSampler.ispc:198-279)
float volumeTransmittance(Ctx &c, const Ray &ray, float rayLower, float rayUpper, RandomSampler &rs)
{
  VolumeHitFwd vh;
  if (!traceVolumeRayFwd(c.w, ray, vh)) return 1.f;
  float lower = fmaxf(ray.t0, vh.t0), upper = hadHit(ray) ? fminf(vh.t1, ray.t) : vh.t1;
  if (lower > upper) return 1.f;
  V3 org = ray.org, dir = ray.dir;
  if (!vh.inst->isIdentity) { const A3 rx = a3(vh.inst->rcp_xfm); org = xfmPoint(rx, ray.org); dir = xfmVector(rx.l, ray.dir); }
  if (upper < rayLower || rayUpper < lower) return 1.f;
  lower = fmaxf(rayLower, lower); upper = fminf(rayUpper, upper);
  float extinction; V3 albedo;
  const float dist = deltaTracking(*vh.vol, lower, upper, rs, org, dir, extinction, albedo);
  return (dist < upper) ? 0.f : 1.f;
}
inline float hgPhase(float cosTheta, float g) // HenyeyGreenstein.ih:11-15
{
  const float denom = 1.f + g * g - 2.f * g * cosTheta;
  return (1.f / (4.f * (float)M_PI)) * (1.f - g * g) * (1.f / (denom * sqrtf(denom)));
}
inline bool hgIsSmooth(float g) { return !(fabsf(g) > 1.f - 1e-3f); }
inline EvalRes HenyeyGreenstein_eval(float g, V3 wo, V3 wi) // :24-35 (the non-smooth branch is overwritten there too)
{
  EvalRes r;
  r.pdf = hgPhase(dot(-wo, wi), g);
  r.value = v3(r.pdf);
  return r;
}
inline SampleRes HenyeyGreenstein_sample(float g, V3 wo, V2 s) // :37-72
{
  SampleRes res;
  res.weight = v3(1.f);
  float cosTheta;
  if (!hgIsSmooth(g)) {
    res.wi = g > 0 ? -wo : wo;
    res.pdf = INF;
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
  const float phi = (float)(2.f * M_PI * s.y); // M_PI is <cmath>'s double in the reference's scalar C++ form
  res.wi = xfmVector(frame(-wo), cartesian(phi, sinTheta, cosTheta));
  res.pdf = hgPhase(cosTheta, g);
  return res;
}

V3 transparentShadow(Ctx &c, V3 lightContrib, Ray &shadowRay) // pathtracer/TransparentShadow.ispc:18-85
{
  uint32_t maxDepth = c.ren.maxDepth;
  const float tOriginal = shadowRay.t;
  while (1) {
    c.cnt.shadow++;
    traceWorld(c.w, shadowRay, false);
    if (!hadHit(shadowRay)) return lightContrib;
    DG dg;
    postIntersect(c.w, c.ren, dg, shadowRay, DG_NS | DG_NG | DG_FACEFORWARD | DG_NORMALIZE | DG_TEXCOORD | DG_COLOR | DG_MOTIONBLUR);
    lightContrib = lightContrib * OBJ_getTransparency(dg, shadowRay.dir);
    if (rmax(lightContrib) <= c.ren.minContribution || --maxDepth <= 0) return v3(0.f);
    setRayInterval(shadowRay, shadowRay.t + dg.epsilon, tOriginal);
  }
}
struct LastVertex { V3 P, Ns; float pdf_w; uint32_t numLightSamples; };
V3 evaluateVirtualLights(Ctx &c, bool straightPath, V3 throughput, const LastVertex &last, bool environment, const Ray &ray, RandomSampler &rs)
{ // pathtracer/VirtualLight.ispc:17-107
  V3 L = v3(0.f);
  const uint32_t numLights = c.w->numLights;
  if (numLights == 0) return L;
  const float lower = length(last.P - ray.org);
  const float upper = environment ? INF : length(last.P - (ray.org + ray.t * ray.dir));
  const float selectionPDF = last.numLightSamples / (float)numLights;
  for (uint32_t i = 0; i < numLights; i++) {
    const OcLight &l = c.w->lights[i];
    if (!straightPath || l.isVisible) {
      const LightEval le = Light_eval(l, last.P, last.Ns, ray.dir, lower, upper);
      if (rmax(le.radiance) > 0.0f) {
        float misWeight = 1.0f;
        if (selectionPDF) {
          float T = 1.f;
          if (c.w->hasVolumes) { // :76-101
            Ray shadowRay;
            setRay(shadowRay, last.P, ray.dir, lower, upper);
            T = volumeTransmittance(c, shadowRay, lower, upper, rs);
          }
          misWeight = misHeuristic(last.pdf_w * T, selectionPDF * le.pdf);
        }
        L = L + throughput * le.radiance * misWeight;
      }
    }
  }
  return L;
}
V3 nextEventEstimation(Ctx &c, const LDSampler &ld, uint32_t numLightSamples, V3 throughput, const DG &dg, const ObjBSDF &bsdf, V3 wo,
    const OcVolume *volume, RandomSampler &rs)
{ // pathtracer/NextEventEstimation.ispc:25-160 (volume != nullptr: the vertex is a volume scattering event)
  V3 L = v3(0.f);
  for (uint32_t i = 0; i < numLightSamples; i++) {
    float ss;
    const V2 s = LDSampler_get3LightSamples(ld, i, numLightSamples > 1, ss);
    const Sample1D sel = Distribution1D_sample((int)c.w->numLights, c.w->lightsCDF, 0, ss);
    const float lightSelectionProb = sel.prob * numLightSamples;
    LightSample ls = Light_sample(c.w->lights[sel.idx], dg, s);
    ls.weight = ls.weight / lightSelectionProb;
    ls.pdf *= lightSelectionProb;
    if (rmax(ls.weight) <= 0.0f || ls.pdf <= 0.0f) continue;
    const EvalRes fe = volume ? HenyeyGreenstein_eval(volume->anisotropy, wo, ls.dir) : OBJ_eval(bsdf, wo, ls.dir);
    if (rmax(fe.value) <= 0.0f) continue;
    V3 org = dg.P;
    if (!volume && dot(dg.Ng, ls.dir) < 0.f) org = org - (2.0f * dg.epsilon) * dg.Ng;
    Ray shadowRay;
    setRay(shadowRay, org, ls.dir, 0.f, ls.dist);
    const V3 unshadedLightContrib = throughput * fe.value * ls.weight;
    const V3 lightContrib = transparentShadow(c, unshadedLightContrib, shadowRay);
    if (rmax(lightContrib) > 0) {
      if (c.w->hasVolumes) { // :126-151: two independent transmittance estimates, one of them for the MIS weight
        const float T = volumeTransmittance(c, shadowRay, 0.f, ls.dist, rs);
        if (T > 0) {
          const float T_mis = volumeTransmittance(c, shadowRay, 0.f, ls.dist, rs);
          L = L + T * lightContrib * misHeuristic(ls.pdf, fe.pdf * T_mis);
        }
      } else
        L = L + lightContrib * misHeuristic(ls.pdf, fe.pdf);
    }
  }
  return L;
}
void samplePath(Ctx &c, LDSampler &ld, RandomSampler &rs, uint32_t nFirst, uint32_t nIndirect, Ray &ray, ScreenSample &sample)
{ // pathtracer/PathSampler.ispc:109-415 (no shadow catcher, no geometry lights, no media)
  LastVertex last = {ray.org, ray.dir, INF, 0};
  bool straightPath = true, specularTransmissionPath = true, firstBounceLight = true;
  uint32_t depth = 0, scatteringEvents = 0;
  V3 throughput = v3(1.f), contribution = v3(0.f);
  sample.alpha = 1.0f;
  const OcRenderer &ren = c.ren;
  do {
    if (depth == 0) c.cnt.primary++; else c.cnt.secondary++;
    const float rayLower = ray.t0, rayUpper = ray.t; // traceClippingRay without clipping geometry (World.ih:150-163)
    traceWorld(c.w, ray, false);
    bool surface = hadHit(ray);
    const V3 wo = -ray.dir;
    const OcVolume *volume = nullptr; // != nullptr: this vertex is a scattering event inside that volume (:178-194)
    V3 volumeAlbedo = v3(0.f);
    if (c.w->hasVolumes) {
      const float freePath = volumeSampleFreePath(c, ray, rayLower, rayUpper, rs, volume, volumeAlbedo);
      if (freePath < INF) { surface = false; throughput = throughput * volumeAlbedo; } else volume = nullptr;
    }
    const bool environment = !surface && !volume;
    if (depth == 0) sample.z = c.fb.projectedDepth ? ray.t * dot(ray.dir, v3(c.cam.direction)) : ray.t;
    if (environment && (ren.backgroundRefraction ? specularTransmissionPath : straightPath)) {
      const V4 bg = getBackground(ren);
      sample.alpha -= luminance(throughput) * (1.0f - bg.w);
      contribution = contribution + throughput * v3(bg);
    }
    if (!volume) contribution = contribution + evaluateVirtualLights(c, straightPath, throughput, last, environment, ray, rs);
    if (environment) break;
    if (depth >= ren.maxDepth) break; // numGeoLights == 0
    DG dg;
    if (volume) { // postIntersect for a volume vertex (:51-67) + VolumetricModel_postIntersect (volume/VolumetricModel.ih:16-33)
      dg.P = ray.org + ray.t * ray.dir;
      dg.Ng = dg.Ns = dg.dPds = v3(0.f);
      dg.color = {0.f, 0.f, 0.f, 0.f};
      dg.epsilon = 0.f; dg.material = nullptr; dg.primID = 0;
      dg.objID = volume->userID == INVALID_ID ? ray.geomID : volume->userID;
      dg.instID = ray.instID;
      if (ray.instID != INVALID_ID && c.w->instances[ray.instID].userID != INVALID_ID) dg.instID = c.w->instances[ray.instID].userID;
    } else
      postIntersect(c.w, ren, dg, ray, DG_NS | DG_NG | DG_FACEFORWARD | DG_NORMALIZE | DG_TEXCOORD | DG_COLOR | DG_TANGENTS | DG_MOTIONBLUR);
    if (depth == 0) {
      sample.position = dg.P; sample.firstNormal = dg.Ns;
      sample.primID = dg.primID; sample.geomID = dg.objID; sample.instID = dg.instID;
    }
    if (depth >= ren.maxDepth || scatteringEvents >= (uint32_t)ren.maxScatteringEvents) break;
    ObjBSDF bsdf;
    if (!volume) {
      OBJ_getBSDF(bsdf, dg);
      if (!bsdf.valid) break;
    }
    LDSampler_nextGroup(ld);
    uint32_t vertexLightSamples = 0;
    if (volume ? hgIsSmooth(volume->anisotropy) : (bsdf.scatteringType & SC_SMOOTH) != 0) { // isSmooth (PathTracerUtil.ih:15-28)
      if (scatteringEvents == 0) { // updateAuxilliaryData (PathTracerUtil.ih:48-60)
        sample.albedo = throughput * (volume ? volumeAlbedo : bsdf.albedo);
        sample.normal = throughput * (volume ? dg.Ns : bsdf.frame.vz);
      }
      vertexLightSamples = firstBounceLight ? nFirst : nIndirect;
      contribution = contribution + nextEventEstimation(c, ld, vertexLightSamples, throughput, dg, bsdf, wo, volume, rs);
      firstBounceLight = false;
    }
    float ss; uint32_t uu3;
    const V2 s = LDSampler_getNext4Samples(ld, ss, uu3);
    SampleRes fs = volume ? HenyeyGreenstein_sample(volume->anisotropy, wo, s) : OBJ_sample(bsdf, wo, s, ss);
    if (!volume) {
      const bool sameSide = dot(fs.wi, dg.Ng) >= 0.f;
      const bool reflection = fs.type & SC_REFLECTION;
      if (sameSide != reflection) fs.weight = v3(0.0f);
    }
    if (rmax(fs.weight) <= 0.0f || fs.pdf <= 0.0f) break;
    throughput = throughput * fs.weight;
    if (depth >= (uint32_t)ren.rouletteDepth) {
      const float rr = LDSampler_finalizeDim3(ld, uu3);
      const float contProb = fminf(rmax(throughput), 0.95f);
      if (rr > contProb) break;
      throughput = throughput * (1.f / contProb);
    }
    V3 ray_org = dg.P;
    if (!volume && (fs.type & SC_TRANSMISSION)) ray_org = ray_org - (2.0f * dg.epsilon) * dg.Ng;
    if ((fs.type & ~SC_SPECULAR_TRANSMISSION) || volume) last = {dg.P, dg.Ns, fs.pdf, vertexLightSamples};
    if (!eq(ray.dir, fs.wi)) straightPath = false;
    if (!(fs.type & SC_SPECULAR_TRANSMISSION)) specularTransmissionPath = false;
    setRay(ray, ray_org, fs.wi, 0.f, INF);
    depth++;
    if (fs.type & SC_SMOOTH) scatteringEvents++;
  } while (rmax(throughput) > ren.minContribution);
  sample.rgb = contribution;
  if (std::isnan(contribution.x) || std::isnan(contribution.y) || std::isnan(contribution.z)) { sample.rgb = v3(0.f); sample.alpha = 1.0f; }
}
inline uint32_t clz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
ScreenSample PathTracer_renderPixel(Ctx &c, uint32_t ix, uint32_t iy) // pathtracer/PathTracer.ispc:75-216
{
  const OcRenderer &ren = c.ren;
  const OcFrameBuffer &fb = c.fb;
  ScreenSample screenSample = make_ScreenSample_zero();
  const uint32_t numLights = c.w->numLights;
  const uint32_t nFirst = (ren.numLightSamples >= 0 && numLights > 0) ? (uint32_t)ren.numLightSamples : std::min(1024u, numLights);
  const uint32_t nIndirect = (ren.limitIndirectLightSamples && nFirst >= 1) ? 1u : nFirst;
  const int spp = ren.spp;
  uint32_t sampleOffs, sampleOffsLight, sampleBits, sampleBitsLight, seed = 0x49c5f37eu;
  if (fb.targetFrames) {
    const uint32_t targetSpp = (uint32_t)fb.targetFrames * spp;
    sampleOffs = OwenScramble4(interleaveZOrder16(ix, iy), seed) * targetSpp;
    sampleOffsLight = sampleOffs * nFirst;
    sampleBits = 16 + 32 - clz32(targetSpp - 1);
    sampleBitsLight = 48 - clz32(targetSpp * nFirst - 1);
  } else {
    seed ^= (uint32_t)fb.sizeX * iy + ix;
    sampleOffs = 0; sampleOffsLight = 0; sampleBits = 16; sampleBitsLight = 16;
  }
  const V2 rcpSize = {1.f / (float)fb.sizeX, 1.f / (float)fb.sizeY};
  for (int s = 0; s < spp; s++) {
    const uint32_t sampleID = (uint32_t)fb.frameID * spp + s;
    LDSampler ld;
    LDSampler_init(ld, seed, sampleOffs + sampleID, sampleOffsLight + sampleID * nFirst, sampleBits, sampleBitsLight);
    const V2 pixel_center = {((float)ix + 0.5f) * rcpSize.x, ((float)iy + 0.5f) * rcpSize.y};
    uint32_t ss234[3];
    const V2 pixelSample = LDSampler_getNext5Samples(ld, ss234);
    const V2 pfSample = pixelFilterSample(ren, pixelSample);
    const V2 screen = pixel_center + pfSample * rcpSize;
    V2 lens = {0.5f, 0.5f};
    if (c.cam.needLensSample) lens = LDSampler_finalizeDim23(ld, ss234);
    Ray ray;
    cameraInitRay(c.cam, ray, screen, lens, pixel_center);
    ScreenSample sample = make_ScreenSample_zero(); // PathTraceIntegrator_Li (PathTracer.ispc:37-63)
    RandomSampler rs;
    RandomSampler_init(rs, seed, sampleOffs + sampleID); // pathState.randomSampler (PathTracer.ispc:157): volumes only
    samplePath(c, ld, rs, nFirst, nIndirect, ray, sample);
    sample.rgb = vmin(sample.rgb, v3(ren.maxRadiance));
    ScreenSample_accumulate(screenSample, sample);
  }
  ScreenSample_normalize(screenSample, 1.f / (float)spp);
  return screenSample;
}

// ---- volumes (volume/Volume.ih:26-56 -> Open VKL structuredRegular, un-vendored) --------------------
float volumeSample(const OcVolume &v, V3 p)
{
  const V3 lc = {(p.x - v.gridOrigin.x) / v.gridSpacing.x, (p.y - v.gridOrigin.y) / v.gridSpacing.y, (p.z - v.gridOrigin.z) / v.gridSpacing.z};
  const float ux = (float)(v.dims[0] - 1), uy = (float)(v.dims[1] - 1), uz = (float)(v.dims[2] - 1);
  if (!(lc.x >= 0.f && lc.y >= 0.f && lc.z >= 0.f && lc.x <= ux && lc.y <= uy && lc.z <= uz)) return NAN;
  const float cx = fminf(lc.x, nextafterf(ux, 0.f)), cy = fminf(lc.y, nextafterf(uy, 0.f)), cz = fminf(lc.z, nextafterf(uz, 0.f));
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  const size_t ix = (size_t)fx0, iy = (size_t)fy0, iz = (size_t)fz0;
  const float fx = cx - fx0, fy = cy - fy0, fz = cz - fz0;
  const size_t sy = v.dims[0], sz = (size_t)v.dims[0] * v.dims[1];
  const float *b = v.voxels + ix + iy * sy + iz * sz;
  const float v00 = b[0] + fx * (b[1] - b[0]), v01 = b[sy] + fx * (b[sy + 1] - b[sy]);
  const float v10 = b[sz] + fx * (b[sz + 1] - b[sz]), v11 = b[sz + sy] + fx * (b[sz + sy + 1] - b[sz + sy]);
  const float v0 = v00 + fy * (v01 - v00), v1 = v10 + fy * (v11 - v10);
  return v0 + fz * (v1 - v0);
}
V3 volumeGradient(const OcVolume &v, V3 p)
{
  const float s = volumeSample(v, p);
  V3 g = {volumeSample(v, {p.x + v.gridSpacing.x, p.y, p.z}) - s, volumeSample(v, {p.x, p.y + v.gridSpacing.y, p.z}) - s,
      volumeSample(v, {p.x, p.y, p.z + v.gridSpacing.z}) - s};
  g = {g.x / v.gridSpacing.x, g.y / v.gridSpacing.y, g.z / v.gridSpacing.z};
  if (std::isnan(g.x)) g.x = 1.f; // Volume.ih:49-54
  if (std::isnan(g.y)) g.y = 1.f;
  if (std::isnan(g.z)) g.z = 1.f;
  return g;
}
V4 transferFunctionGet(const OcTransferFunction &tf, float value) // LinearTransferFunction.ispc:11-43
{
  value = (value - tf.valueLower) * (1.f / (tf.valueUpper - tf.valueLower));
  value = clampf(value, 0.0f, 0x1.fffffep-1f);
  const int maxIdxC = (int)tf.numColors - 1;
  const float idxCf = value * maxIdxC;
  const float fracC = idxCf - floorf(idxCf);
  const int idxC = (int)idxCf;
  const V3 col = lerp(fracC, v3(tf.color[idxC]), v3(tf.color[std::min(maxIdxC, idxC + 1)]));
  const int maxIdxO = (int)tf.numOpacities - 1;
  const float idxOf = value * maxIdxO;
  const float fracO = idxOf - floorf(idxOf);
  const int idxO = (int)idxOf;
  return v4(col, lerp(fracO, tf.opacity[idxO], tf.opacity[std::min(maxIdxO, idxO + 1)]));
}
struct VolumeHit { const OcVolume *vol; const OcInstance *inst; uint32_t instID, geomID; float t0, t1; };
// Volume_intersect_kernel (volume/Volume.ispc:22-73): first volume interval wins (single interval)
bool traceVolumeRay(const OcWorld *w, const Ray &ray, VolumeHit &vh)
{
  vh.vol = nullptr;
  float best = INF;
  for (uint32_t i = 0; i < w->numInstances; i++) {
    const OcInstance *inst = w->instances + i;
    const OcGroup *g = inst->group;
    if (g->numVolumes == 0) continue;
    V3 org = ray.org, dir = ray.dir;
    if (!inst->isIdentity) { const A3 rx = a3(inst->rcp_xfm); org = xfmPoint(rx, ray.org); dir = xfmVector(rx.l, ray.dir); }
    for (uint32_t k = 0; k < g->numVolumes; k++) {
      const OcVolume *v = g->volumes + k;
      const V3 rd = {rcp_safe(dir.x), rcp_safe(dir.y), rcp_safe(dir.z)};
      const V3 mins = (v3(v->bounds[0], v->bounds[1], v->bounds[2]) - org) * rd, maxs = (v3(v->bounds[3], v->bounds[4], v->bounds[5]) - org) * rd;
      const float tEntry = rmax(vmin(mins, maxs)), tExit = rmin(vmax(mins, maxs));
      const float tmin = fmaxf(ray.t0, tEntry), tmax = fminf(ray.t, tExit);
      if (tmin < tmax && tmin < best) { best = tmin; vh = {v, inst, i, k, tmin, tmax}; }
    }
  }
  return vh.vol != nullptr;
}
bool traceVolumeRayFwd(const OcWorld *w, const Ray &ray, VolumeHitFwd &vh)
{
  VolumeHit h;
  const bool r = traceVolumeRay(w, ray, h);
  vh = {h.vol, h.inst, h.instID, h.geomID, h.t0, h.t1};
  return r;
}
inline float nominalDeltaT(const OcVolume &v, V3 dir)
{
  float n = INF;
  if (fabsf(dir.x) > 0.f) n = fminf(n, v.gridSpacing.x / fabsf(dir.x));
  if (fabsf(dir.y) > 0.f) n = fminf(n, v.gridSpacing.y / fabsf(dir.y));
  if (fabsf(dir.z) > 0.f) n = fminf(n, v.gridSpacing.z / fabsf(dir.z));
  return n;
}
// render/ao/volumes.ispc (single-interval form :251-353): returns (rgb, alpha)
V4 integrateVolumeAO(Ctx &c, const VolumeHit &vh, const Ray &ray, RandomSampler &rs, float samplingRate)
{
  const OcVolume &v = *vh.vol;
  V3 org = ray.org, dir = ray.dir;
  if (!vh.inst->isIdentity) { const A3 rx = a3(vh.inst->rcp_xfm); org = xfmPoint(rx, ray.org); dir = xfmVector(rx.l, ray.dir); }
  V3 color = v3(0.f);
  float alpha = 0.f;
  const float jitter = RandomSampler_getFloat(rs);
  const float samplingStep = nominalDeltaT(v, dir) / samplingRate;
  const float iuLength = (vh.t1 - vh.t0) / samplingStep;
  float distance = vh.t0, iuDistance = jitter;
  while (alpha < .99f) {
    if (iuDistance > iuLength) break;
    const float newDistance = vh.t0 + iuDistance * samplingStep;
    const float sampleVal = volumeSample(v, org + newDistance * dir);
    if (!std::isnan(sampleVal)) {
      V4 sc = transferFunctionGet(v.tf, sampleVal);
      const float dt = newDistance - distance;
      sc.w = 1.f - expf(-sc.w * dt * v.densityScale);
      color = color + ((1.f - alpha) * sc.w * v3(sc));
      alpha = alpha + ((1.f - alpha) * sc.w);
    }
    iuDistance += 1.f;
    distance = newDistance;
  }
  return v4(color, alpha);
}

// ---- scivis -------------------------------------------------------------------------------------
struct SciVisBSDF { V3 diffuse, specular, transmission, albedo; float shininess, opacity; };
SciVisBSDF evalMaterial(const DG &dg) // render/scivis/surfaces.ih:49-92
{
  SciVisBSDF r;
  r.diffuse = v3(dg.color); r.specular = v3(0.f); r.shininess = 10.f; r.transmission = v3(1.f - dg.color.w); r.opacity = dg.color.w;
  const OcMaterial *m = dg.material;
  if (m && m->type == OC_MATERIAL_OBJ) {
    const float d = m->d * dg.color.w;
    r.diffuse = r.diffuse * v3(m->Kd);
    r.diffuse = r.diffuse * d;
    r.specular = d * v3(m->Ks);
    r.shininess = m->Ns;
    r.transmission = v3(m->Tf) * d + v3(1.f - d);
    r.opacity = d;
  }
  r.albedo = r.diffuse + r.specular;
  r.shininess *= 4.0f;
  r.diffuse = r.diffuse * ONE_OVER_PI;
  r.specular = r.specular * ((r.shininess + 2.f) * ONE_OVER_FOUR_PI * (1.f / (2.f - expf(r.shininess * -LNTWO_OVER_TWO))));
  return r;
}
V3 lightAlpha(Ctx &c, Ray &ray, RandomSampler &rs, V3 weight, float quality) // render/scivis/lightAlpha.ispc:19-122
{
  V3 alpha = v3(1.f);
  const float org_t_max = ray.t;
  while (true) {
    c.cnt.shadow++;
    traceWorld(c.w, ray, false);
    if (c.w->hasVolumes) {
      VolumeHit vh;
      if (traceVolumeRay(c.w, ray, vh)) {
        const V4 vc = integrateVolumeAO(c, vh, ray, rs, c.ren.volumeSamplingRate * quality);
        alpha = alpha * v3(1.f - vc.w);
      }
    }
    if (hadHit(ray)) {
      DG dg;
      postIntersect(c.w, c.ren, dg, ray, DG_COLOR | DG_TEXCOORD);
      alpha = alpha * evalMaterial(dg).transmission;
      setRayInterval(ray, ray.t + dg.epsilon, org_t_max);
    } else break;
    if (luminance(alpha * weight) < c.ren.minContribution) { alpha = v3(0.f); break; }
  }
  return alpha;
}
struct SSI { V3 shadedColor, transmission, albedo; float opacity; };
V3 SciVis_computeAO(Ctx &c, RandomSampler &rs, const DG &dg, int sampleCnt, float aoRadius, uint32_t px, uint32_t py, int sampleIDz)
{ // render/scivis/SciVis.ispc:224-281
  const int accumID = sampleIDz * sampleCnt;
  uint32_t v0 = px, v1 = py;
  tea8(v0, v1);
  const V2 rot = {to_float_unorm(v0), to_float_unorm(v1)};
  V3 hits = v3(0.f);
  const L3 localToWorld = frame(dg.Ns);
  for (int i = 0; i < sampleCnt; i++) {
    const V2 r = {CranleyPattersonRotation(PermHalton_sample5(accumID + i), rot.x), CranleyPattersonRotation(PermHalton_sample7(accumID + i), rot.y)};
    const V3 ao_dir = xfmVector(localToWorld, cosineSampleHemisphere(r));
    if (dot(ao_dir, dg.Ns) < 0.01f) { hits = hits + v3(1.f); continue; }
    Ray ao_ray;
    setRay(ao_ray, dg.P, ao_dir, dg.epsilon, aoRadius);
    hits = hits + (v3(1.f) - lightAlpha(c, ao_ray, rs, v3(1.f), 0.1f));
  }
  return v3(1.0f) - (hits / (float)sampleCnt);
}
SSI SciVis_computeShading(Ctx &c, const DG &dg, RandomSampler &rs, V3 inDir, uint32_t px, uint32_t py, int sampleIDz)
{ // render/scivis/surfaces.ispc:17-128
  SSI ret;
  const SciVisBSDF bsdf = evalMaterial(dg);
  ret.albedo = bsdf.albedo;
  V3 color = v3(0.f);
  const OcWorld *w = c.w;
  for (uint32_t i = w->numScivisLightsVisibleOnly; i < w->numScivisLights; i++) {
    const LightSample light = Light_sample(w->scivisLights[i], dg, {0.f, 0.f});
    if (rmax(light.weight) > 0.f) {
      const float cosNL = dot(light.dir, dg.Ns);
      if (cosNL < 0.0f) continue;
      const V3 H = normalize(light.dir - inDir);
      const float cosNH = dot(H, dg.Ns);
      const V3 specular = bsdf.specular * powf(cosNH, bsdf.shininess);
      const V3 diffuse = bsdf.diffuse * cosNL;
      const V3 light_contrib = (diffuse + specular) * light.weight;
      if (c.ren.shadowsEnabled) {
        if (luminance(light_contrib) > c.ren.minContribution) {
          V3 P = dg.P;
          if (dot(dg.Ng, light.dir) < 0.0f) P = P - (2.f * dg.epsilon) * dg.Ng;
          Ray shadowRay;
          setRay(shadowRay, P, light.dir, 0.0f, light.dist);
          color = color + lightAlpha(c, shadowRay, rs, light_contrib, 0.25f) * light_contrib;
        }
      } else color = color + light_contrib;
    }
  }
  V3 ao = v3(1.f);
  const V3 aoColorPi = v3(w->aoColorPi);
  if (c.ren.aoSamples > 0 && luminance(aoColorPi) > c.ren.minContribution)
    ao = SciVis_computeAO(c, rs, dg, c.ren.aoSamples, c.ren.aoRadius, px, py, sampleIDz);
  color = color + bsdf.diffuse * ao * aoColorPi;
  ret.shadedColor = color; ret.transmission = bsdf.transmission; ret.opacity = bsdf.opacity;
  return ret;
}
// render/scivis/volumes.ispc: sampling loop of the single-interval form (:302-439); depth / ID
// bookkeeping of the CPU form (sampleAllVolumes :160-169): the first sample in front of the
// surface hit with alpha above minContribution defines the volume hit.
V4 integrateVolumeSciVis(Ctx &c, const VolumeHit &vh, const Ray &ray, RandomSampler &rs, float samplingRate, uint32_t px, uint32_t py,
    int sampleIDz, float &volT, bool &volHit)
{
  const OcVolume &v = *vh.vol;
  V3 org = ray.org, dir = ray.dir;
  if (!vh.inst->isIdentity) { const A3 rx = a3(vh.inst->rcp_xfm); org = xfmPoint(rx, ray.org); dir = xfmVector(rx.l, ray.dir); }
  V3 color = v3(0.f);
  float transmission = 1.f;
  const float jitter = RandomSampler_getFloat(rs);
  const float samplingStep = nominalDeltaT(v, dir) / samplingRate;
  const float iuLength = (vh.t1 - vh.t0) / samplingStep;
  float distance = vh.t0, iuDistance = jitter;
  while (transmission > c.ren.minContribution) {
    if (iuDistance > iuLength) break;
    const float newDistance = vh.t0 + iuDistance * samplingStep;
    const V3 p = org + newDistance * dir;
    const float sampleVal = volumeSample(v, p);
    if (!std::isnan(sampleVal)) {
      V4 sc = transferFunctionGet(v.tf, sampleVal);
      const float dt = newDistance - distance;
      if (v.gradientShadingScale > 0.f && sc.w > 0.0f) {
        V3 ns = volumeGradient(v, p);
        if (dot(ns, ns) > 1e-6f) {
          ns = -ns;
          DG dg;
          dg.color = sc; dg.material = nullptr; dg.epsilon = dt / 2.f;
          dg.Ns = dg.Ng = normalize(xfmVector(transposed(a3(vh.inst->rcp_xfm).l), ns));
          dg.P = ray.org + newDistance * ray.dir;
          dg.dPds = v3(1.f, 0.f, 0.f); dg.primID = 0; dg.objID = dg.instID = INVALID_ID;
          const SSI shading = SciVis_computeShading(c, dg, rs, ray.dir, px, py, sampleIDz);
          sc = lerp(v.gradientShadingScale, sc, v4(shading.shadedColor, 1.f - luminance(shading.transmission)));
        }
      }
      sc.w = expf(-sc.w * dt * v.densityScale);
      if (newDistance < volT && sc.w < 1.0f - c.ren.minContribution) { volT = newDistance; volHit = true; }
      color = color + transmission * (1.f - sc.w) * v3(sc);
      transmission *= sc.w;
    }
    iuDistance += 1.f;
    distance = newDistance;
  }
  return v4(color, transmission);
}
V3 evaluateLightsSciVis(Ctx &c, const Ray &ray) // surfaces.ispc:130-150
{
  V3 color = v3(0.f);
  for (uint32_t i = 0; i < c.w->numScivisLights; i++) {
    const OcLight &l = c.w->scivisLights[i];
    if (!l.isVisible) continue;
    color = color + Light_eval(l, ray.org, v3(0.f), ray.dir, ray.t0, ray.t).radiance;
  }
  return color;
}
void SciVis_renderSample(Ctx &c, Ray &ray, ScreenSample &sample, uint32_t px, uint32_t py, int sampleIDz) // SciVis.ispc:24-222
{
  const OcRenderer &ren = c.ren;
  RandomSampler rs;
  RandomSampler_init(rs, px + (uint32_t)c.fb.sizeX * py, (uint32_t)sampleIDz);
  bool firstHit = true, allHitsFullyTransmissive = true, primary = true;
  const float originalRayTFar = ray.t;
  V3 outColor = v3(0.f), outTransmission = v3(1.f);
  while (true) {
    if (primary) c.cnt.primary++; else c.cnt.secondary++;
    primary = false;
    traceWorld(c.w, ray, false);
    float volumeT = ray.t;
    if (c.w->hasVolumes) {
      VolumeHit vh;
      if (traceVolumeRay(c.w, ray, vh)) {
        bool volHit = false;
        const V4 vc = integrateVolumeSciVis(c, vh, ray, rs, ren.volumeSamplingRate, px, py, sampleIDz, volumeT, volHit);
        outColor = outColor + outTransmission * v3(vc);
        outTransmission = outTransmission * vc.w;
        if (volHit && sample.geomID == INVALID_ID) {
          sample.primID = 0;
          sample.instID = (vh.inst->userID == INVALID_ID) ? vh.instID : vh.inst->userID;
          sample.geomID = (vh.vol->userID == INVALID_ID) ? vh.geomID : vh.vol->userID;
        }
      }
    }
    if (ren.visibleLights) outColor = outColor + outTransmission * evaluateLightsSciVis(c, ray);
    if (firstHit || allHitsFullyTransmissive) {
      sample.z = fminf(ray.t, volumeT);
      if (c.fb.projectedDepth) sample.z *= dot(ray.dir, v3(c.cam.direction));
    }
    if (hadHit(ray)) {
      DG dg;
      postIntersect(c.w, ren, dg, ray, DG_NG | DG_NS | DG_NORMALIZE | DG_FACEFORWARD | DG_COLOR | DG_TANGENTS | DG_TEXCOORD);
      const SSI sh = SciVis_computeShading(c, dg, rs, ray.dir, px, py, sampleIDz);
      if (firstHit || allHitsFullyTransmissive) {
        sample.position = dg.P; sample.firstNormal = dg.Ns; sample.normal = dg.Ns; sample.albedo = sh.albedo;
        if (sample.geomID == INVALID_ID) { sample.instID = dg.instID; sample.geomID = dg.objID; sample.primID = dg.primID; }
        firstHit = false;
      }
      allHitsFullyTransmissive = allHitsFullyTransmissive && sh.opacity < ren.minContribution;
      outColor = outColor + outTransmission * sh.shadedColor;
      outTransmission = outTransmission * sh.transmission;
      if (luminance(outTransmission) < ren.minContribution) { outTransmission = v3(0.f); break; }
      setRayInterval(ray, ray.t + dg.epsilon, originalRayTFar);
    } else {
      const V4 bg = getBackground(ren);
      outColor = outColor + outTransmission * v3(bg);
      outTransmission = outTransmission * (1.f - bg.w);
      break;
    }
  }
  sample.rgb = outColor;
  sample.alpha = 1.f - luminance(outTransmission);
}

// ---- ao ------------------------------------------------------------------------------------------
float computeAO(Ctx &c, const DG &dg, int sampleCnt, float aoRadius, uint32_t px, uint32_t py, int sampleIDz) // render/util.ispc:12-54
{
  const int accumID = sampleIDz * sampleCnt;
  uint32_t v0 = px, v1 = py;
  tea8(v0, v1);
  const V2 rot = {to_float_unorm(v0), to_float_unorm(v1)};
  int hits = 0;
  const L3 localToWorld = frame(dg.Ns);
  for (int i = 0; i < sampleCnt; i++) {
    const V2 r = {CranleyPattersonRotation(PermHalton_sample5(accumID + i), rot.x), CranleyPattersonRotation(PermHalton_sample7(accumID + i), rot.y)};
    const V3 ao_dir = xfmVector(localToWorld, cosineSampleHemisphere(r));
    if (dot(ao_dir, dg.Ns) < 0.01f) { hits++; continue; }
    Ray ao_ray;
    setRay(ao_ray, dg.P, ao_dir, dg.epsilon, aoRadius);
    c.cnt.shadow++;
    if (traceWorld(c.w, ao_ray, true)) hits++;
  }
  return 1.0f - (hits / (float)sampleCnt);
}
void AO_renderSample(Ctx &c, Ray &ray, ScreenSample &sample, uint32_t px, uint32_t py, int sampleIDz) // ao/AORenderer.ispc:23-149
{
  const OcRenderer &ren = c.ren;
  RandomSampler rs;
  RandomSampler_init(rs, px + (uint32_t)c.fb.sizeX * py, (uint32_t)sampleIDz);
  bool firstHit = true;
  const float originalRayTFar = ray.t;
  const V3 primaryDir = ray.dir;
  V4 outputColor = {0.f, 0.f, 0.f, 0.f};
  while (outputColor.w < 0.99f) {
    if (firstHit) c.cnt.primary++; else c.cnt.secondary++;
    traceWorld(c.w, ray, false);
    if (c.w->hasVolumes) {
      VolumeHit vh;
      if (traceVolumeRay(c.w, ray, vh)) outputColor = outputColor + (1.f - outputColor.w) * integrateVolumeAO(c, vh, ray, rs, ren.volumeSamplingRate);
    }
    V4 blendedColor;
    const bool rayHadHit = hadHit(ray);
    if (rayHadHit) {
      DG dg;
      postIntersect(c.w, ren, dg, ray, DG_NG | DG_NS | DG_NORMALIZE | DG_FACEFORWARD | DG_COLOR | DG_TEXCOORD | DG_TANGENTS);
      V3 surfaceColor = v3(1.f); // getSurfaceColor (ao/surfaces.ispc:18-42)
      float opacity = 1.f;
      if (dg.material && dg.material->type == OC_MATERIAL_OBJ) { surfaceColor = v3(dg.material->Kd); opacity *= dg.material->d; }
      const V3 color = surfaceColor * v3(dg.color);
      opacity = opacity * dg.color.w;
      float intensity = fabsf(dot(dg.Ns, primaryDir)) * opacity; // AORenderer_computeShading (:44-74)
      if (ren.aoSamples > 0 && ren.aoIntensity > 0.01f) {
        const float ao = computeAO(c, dg, ren.aoSamples, ren.aoRadius, px, py, sampleIDz);
        intensity *= (1.0f - ren.aoIntensity) + (ao * ren.aoIntensity);
      }
      blendedColor = v4(color * intensity, opacity);
      if (firstHit) {
        sample.z = ray.t;
        if (c.fb.projectedDepth) sample.z *= dot(ray.dir, v3(c.cam.direction));
        sample.albedo = color; sample.position = dg.P; sample.firstNormal = dg.Ns; sample.normal = dg.Ns;
        sample.instID = dg.instID; sample.geomID = dg.objID; sample.primID = dg.primID;
      }
      setRayInterval(ray, ray.t + dg.epsilon, originalRayTFar);
    } else {
      blendedColor = getBackground(ren);
      if (firstHit) { sample.z = ray.t; sample.albedo = v3(blendedColor); sample.normal = primaryDir; }
    }
    outputColor = outputColor + (1.f - outputColor.w) * blendedColor;
    firstHit = false;
    if (!rayHadHit) break;
  }
  sample.rgb = v3(outputColor);
  sample.alpha = outputColor.w;
}
// ---- debug renderer: render/debug/DebugRenderer.ispc:27-420 (texCoord, dPdt and volume methods are not restated) ----
inline V3 makeRandomColor(uint32_t i) // rkcommon/utility/random.ih (un-vendored; restated)
{
  const uint32_t mx = 13 * 17 * 43, my = 11 * 29, mz = 7 * 23 * 63;
  const uint32_t g = i * (3 * 5 * 127) + 12312314;
  if (i == 0xFFFFFFFFu) return {0.f, 0.f, 0.f};
  return {(g % mx) * (1.f / (mx - 1)), (g % my) * (1.f / (my - 1)), (g % mz) * (1.f / (mz - 1))};
}
inline float debugEyeLight(const Ray &ray) { return 0.2f + 0.8f * fabsf(dot(normalize(ray.Ng), ray.dir)); } // :39-42
inline V3 debugTangentColor(V3 t) // :133-141
{
  V3 c = normalize(t);
  if (c.x < 0.f) c.x = c.x * -0.3f;
  if (c.y < 0.f) c.y = c.y * -0.3f;
  if (c.z < 0.f) c.z = c.z * -0.3f;
  return c;
}
void Debug_renderSample(Ctx &c, Ray &ray, ScreenSample &sample, uint32_t px, uint32_t py, int sampleIDz)
{
  const OcRenderer &ren = c.ren;
  const uint32_t method = ren.debugMethod;
  if (method == OC_DEBUG_TEST_FRAME) { // :47-62
    sample.rgb = {(px % 256) / 255.f, (py % 256) / 255.f, ((px + py + (uint32_t)sampleIDz) % 256) / 255.f};
    sample.alpha = 1.f; sample.z = 1.f;
  } else if (method == OC_DEBUG_RAY_DIR) { // :66-72
    sample.rgb = vabs(ray.dir);
    sample.alpha = 1.f; sample.z = 1.f;
  } else {
    c.cnt.primary++;
    traceWorld(c.w, ray, false); // hitBackground :27-38
    sample.z = ray.t;
    sample.alpha = 1.f;
    sample.rgb = v3(getBackground(ren));
    if (hadHit(ray)) {
      DG dg;
      switch (method) {
      case OC_DEBUG_EYE_LIGHT: sample.rgb = v3(debugEyeLight(ray)); break;
      case OC_DEBUG_NG: postIntersect(c.w, ren, dg, ray, DG_NORMALIZE | DG_NG); sample.rgb = vabs(dg.Ng); break;
      case OC_DEBUG_NS: postIntersect(c.w, ren, dg, ray, DG_NORMALIZE | DG_NS); sample.rgb = vabs(dg.Ns); break;
      case OC_DEBUG_COLOR:
        postIntersect(c.w, ren, dg, ray, DG_COLOR | DG_NS);
        sample.rgb = v3(dg.color) * fabsf(dot(normalize(ray.dir), normalize(dg.Ns)));
        break;
      case OC_DEBUG_DPDS: postIntersect(c.w, ren, dg, ray, DG_TANGENTS); sample.rgb = debugTangentColor(dg.dPds); break;
      case OC_DEBUG_PRIM_ID: sample.rgb = debugEyeLight(ray) * makeRandomColor(ray.primID); break;
      case OC_DEBUG_GEOM_ID: sample.rgb = debugEyeLight(ray) * makeRandomColor(ray.geomID); break;
      case OC_DEBUG_INST_ID: sample.rgb = debugEyeLight(ray) * makeRandomColor(ray.instID); break;
      case OC_DEBUG_BACKFACING_NG:
        sample.rgb = v3(debugEyeLight(ray));
        if (dot(ray.Ng, ray.dir) > 0.f) sample.rgb.y = 0.f;
        break;
      case OC_DEBUG_BACKFACING_NS: {
        postIntersect(c.w, ren, dg, ray, DG_NORMALIZE | DG_NS);
        const float cs = dot(dg.Ns, ray.dir);
        sample.rgb = v3(.2f + .8f * fabsf(cs));
        if (cs > 0.f) sample.rgb.y = 0.f;
        break;
      }
      default: break;
      }
    }
  }
  if (c.fb.projectedDepth) sample.z *= dot(ray.dir, v3(c.cam.direction)); // :397-398
}
// render/RendererRenderTaskFn.inl:30-86 -- NOTE `sample` lives outside the spp loop (not re-zeroed)
ScreenSample defaultRenderPixel(Ctx &c, uint32_t x, uint32_t y)
{
  const OcRenderer &ren = c.ren;
  const OcFrameBuffer &fb = c.fb;
  const int spp = ren.spp;
  const int startSampleID = fb.frameID * spp + 1;
  ScreenSample screenSample = make_ScreenSample_zero();
  ScreenSample sample = make_ScreenSample_zero();
  const V2 rcpSize = {1.f / (float)fb.sizeX, 1.f / (float)fb.sizeY};
  for (int s = 0; s < spp; s++) {
    const float pixel_du = Halton_sample2(startSampleID + s);
    const float pixel_dv = CranleyPattersonRotation(Halton_sample3(startSampleID + s), 1.f / 6.f);
    const V2 pfSample = pixelFilterSample(ren, {pixel_du, pixel_dv});
    const V2 pixel_center = {((float)x + 0.5f) * rcpSize.x, ((float)y + 0.5f) * rcpSize.y};
    const V2 screen = pixel_center + pfSample * rcpSize;
    Ray ray;
    cameraInitRay(c.cam, ray, screen, {0.f, 0.f}, pixel_center);
    if (ren.type == OC_RENDERER_SCIVIS) SciVis_renderSample(c, ray, sample, x, y, startSampleID + s);
    else if (ren.type == OC_RENDERER_DEBUG) Debug_renderSample(c, ray, sample, x, y, startSampleID + s);
    else AO_renderSample(c, ray, sample, x, y, startSampleID + s);
    ScreenSample_accumulate(screenSample, sample);
  }
  ScreenSample_normalize(screenSample, 1.f / (float)spp);
  return screenSample;
}

// ================================ framebuffer (fb/LocalFB.ispc:14-78, FrameBuffer.ih:17-22) =========
inline float accum1(float value, float nv, float f) { return (f > 0.8f) ? nv : lerp(f, value, nv); }
inline void accum3(OcVec3f *b, size_t i, V3 nv, float f) { b[i].x = accum1(b[i].x, nv.x, f); b[i].y = accum1(b[i].y, nv.y, f); b[i].z = accum1(b[i].z, nv.z, f); }
inline void accum4(OcVec4f *b, size_t i, V4 nv, float f)
{
  b[i].x = accum1(b[i].x, nv.x, f); b[i].y = accum1(b[i].y, nv.y, f); b[i].z = accum1(b[i].z, nv.z, f); b[i].w = accum1(b[i].w, nv.w, f);
}
void accumulateSample(const OcFrameBuffer &fb, uint32_t x, uint32_t y, const ScreenSample &s)
{
  const size_t pixelID = x + (size_t)fb.sizeX * y;
  const int accumID = fb.frameID;
  const float accScale = 1.f / (float)(accumID + 1);
  const V4 colorSample = v4(s.rgb, s.alpha);
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

void renderRows(const OcWorld *w, const OcCamera *cam, const OcRenderer *ren, const OcFrameBuffer *fb, std::atomic<int> *nextRow,
    Counters *out)
{
  Ctx c{w, *ren, *cam, *fb, Counters()};
  const uint32_t numRanks = fb->numRanks ? fb->numRanks : 1, rank = fb->numRanks ? fb->rank : 0;
  const uint32_t tilesX = (fb->sizeX + 63) / 64;
  for (;;) {
    const int y = nextRow->fetch_add(1);
    if (y >= fb->sizeY) break;
    for (int x = 0; x < fb->sizeX; x++) {
      // image-tile partition of the multidevice path (MultiDevice.cpp:403-416: tile % N == rank)
      const uint32_t tileID = (uint32_t)(y / 64) * tilesX + (uint32_t)(x / 64);
      if (tileID % numRanks != rank) continue;
      // adaptive accumulation: only render tasks above the error threshold (LocalFB.cpp:238-261)
      if (fb->taskError && !(fb->taskError[(size_t)(y / 8) * ((fb->sizeX + 7) / 8) + (x / 8)] > fb->errorThreshold)) continue;
      const ScreenSample ss = ren->type == OC_RENDERER_PATHTRACER ? PathTracer_renderPixel(c, x, y) : defaultRenderPixel(c, x, y);
      accumulateSample(*fb, x, y, ss);
    }
  }
  *out = c.cnt;
}

inline uint32_t cvt_uint32(float f) { return (uint32_t)rintf(255.f * fminf(fmaxf(f, 0.f), 1.f)); } // rkcommon
inline float linear_to_srgb(float f) { const float c = fmaxf(f, 0.f); return c <= 0.0031308f ? 12.92f * c : powf(c, 1.f / 2.4f) * 1.055f - 0.055f; }

} // namespace

// ================================ C entry points ======================================================
ORACLE_API int oracle_render_frame(const OcWorld *w, const OcCamera *cam, const OcRenderer *ren, const OcFrameBuffer *fb, int numThreads)
{
  if (numThreads <= 0) numThreads = (int)std::max(1u, std::thread::hardware_concurrency());
  numThreads = std::min(numThreads, std::max(1, fb->sizeY));
  std::atomic<int> nextRow{0};
  std::vector<Counters> cnt(numThreads);
  std::vector<std::thread> th;
  for (int i = 1; i < numThreads; i++) th.emplace_back(renderRows, w, cam, ren, fb, &nextRow, &cnt[i]);
  renderRows(w, cam, ren, fb, &nextRow, &cnt[0]);
  for (auto &t : th) t.join();
  if (fb->rayCounters) {
    unsigned long long p = 0, s = 0, h = 0;
    for (auto &c : cnt) { p += c.primary; s += c.secondary; h += c.shadow; }
    fb->rayCounters[0] += p; fb->rayCounters[1] += s; fb->rayCounters[2] += h;
  }
  return 0;
}

ORACLE_API int oracle_trace(const OcWorld *w, OcRayHit *rays, uint32_t n, int occlusionOnly)
{
  for (uint32_t i = 0; i < n; i++) {
    OcRayHit &rh = rays[i];
    Ray r;
    setRay(r, v3(rh.org[0], rh.org[1], rh.org[2]), v3(rh.dir[0], rh.dir[1], rh.dir[2]), rh.t0, rh.t);
    if (occlusionOnly) { if (traceWorld(w, r, true)) r.t = -INF; }
    else traceWorld(w, r, false);
    rh.t = r.t; rh.u = r.u; rh.v = r.v; rh.Ng[0] = r.Ng.x; rh.Ng[1] = r.Ng.y; rh.Ng[2] = r.Ng.z;
    rh.primID = r.primID; rh.geomID = r.geomID; rh.instID = r.instID;
  }
  return 0;
}

// primary camera rays for given pixels with the pathtracer (renderer 0) or default-task (1) sample streams
ORACLE_API void oracle_camera_rays(const OcCamera *cam, const OcRenderer *ren, const OcFrameBuffer *fb, const uint32_t *px,
    const uint32_t *py, uint32_t n, int sampleIndex, OcRayHit *out)
{
  const V2 rcpSize = {1.f / (float)fb->sizeX, 1.f / (float)fb->sizeY};
  for (uint32_t i = 0; i < n; i++) {
    V2 pixelSample;
    if (ren->type == OC_RENDERER_PATHTRACER) {
      const uint32_t seed = 0x49c5f37eu ^ ((uint32_t)fb->sizeX * py[i] + px[i]);
      LDSampler ld;
      LDSampler_init(ld, seed, (uint32_t)sampleIndex, (uint32_t)sampleIndex, 16, 16);
      uint32_t ss[3];
      pixelSample = LDSampler_getNext5Samples(ld, ss);
    } else {
      pixelSample = {Halton_sample2(sampleIndex + 1), CranleyPattersonRotation(Halton_sample3(sampleIndex + 1), 1.f / 6.f)};
    }
    const V2 pf = pixelFilterSample(*ren, pixelSample);
    const V2 pixelCenter = V2{((float)px[i] + 0.5f) * rcpSize.x, ((float)py[i] + 0.5f) * rcpSize.y};
    const V2 screen = pixelCenter + pf * rcpSize;
    Ray r;
    cameraInitRay(*cam, r, screen, {0.5f, 0.5f}, pixelCenter);
    OcRayHit &o = out[i];
    o.org[0] = r.org.x; o.org[1] = r.org.y; o.org[2] = r.org.z; o.t0 = r.t0;
    o.dir[0] = r.dir.x; o.dir[1] = r.dir.y; o.dir[2] = r.dir.z; o.t = r.t;
    o.Ng[0] = o.Ng[1] = o.Ng[2] = 0.f; o.u = o.v = 0.f; o.primID = o.geomID = o.instID = INVALID_ID;
  }
}

// ---- sampler known-answer streams ---------------------------------------------------------------------
// kind 0: LDSampler: per index i -> [px.x, px.y, lens.x, lens.y, time, light.x, light.y, light.ss, bsdf.x, bsdf.y, bsdf.ss, rr]
//         (getNext5Samples+finalize, nextGroup, get3LightSamples(0,false), getNext4Samples+finalizeDim3)
// kind 1: Halton2, CP(Halton3,1/6), PermHalton5, PermHalton7 (4 floats)     kind 2: PCG32 init(seed,i) then 4 floats
// kind 3: tea8(seed, i) -> 2 floats                                        kind 4: OwenScramble4(interleaveZOrder16(i,seed),0x49c5f37e) as float bits
ORACLE_API void oracle_sampler_kat(int kind, uint32_t seed, uint32_t first, uint32_t count, float *out)
{
  for (uint32_t k = 0; k < count; k++) {
    const uint32_t i = first + k;
    if (kind == 0) {
      float *o = out + 12 * (size_t)k;
      LDSampler ld;
      LDSampler_init(ld, seed, i, i, 16, 16);
      uint32_t ss[3];
      const V2 p = LDSampler_getNext5Samples(ld, ss);
      const V2 l = LDSampler_finalizeDim23(ld, ss);
      o[0] = p.x; o[1] = p.y; o[2] = l.x; o[3] = l.y; o[4] = LDSamplerScramble(ss[2], ld.scramble ^ SCRAMBLE4);
      LDSampler_nextGroup(ld);
      float s1;
      const V2 ls = LDSampler_get3LightSamples(ld, 0, false, s1);
      o[5] = ls.x; o[6] = ls.y; o[7] = s1;
      float s2; uint32_t u3;
      const V2 b = LDSampler_getNext4Samples(ld, s2, u3);
      o[8] = b.x; o[9] = b.y; o[10] = s2; o[11] = LDSampler_finalizeDim3(ld, u3);
    } else if (kind == 1) {
      float *o = out + 4 * (size_t)k;
      o[0] = Halton_sample2(i); o[1] = CranleyPattersonRotation(Halton_sample3(i), 1.f / 6.f); o[2] = PermHalton_sample5(i); o[3] = PermHalton_sample7(i);
    } else if (kind == 2) {
      float *o = out + 4 * (size_t)k;
      RandomSampler rs;
      RandomSampler_init(rs, seed, i);
      for (int j = 0; j < 4; j++) o[j] = RandomSampler_getFloat(rs);
    } else if (kind == 3) {
      uint32_t v0 = seed, v1 = i;
      tea8(v0, v1);
      out[2 * (size_t)k] = to_float_unorm(v0); out[2 * (size_t)k + 1] = to_float_unorm(v1);
    } else {
      const uint32_t v = OwenScramble4(interleaveZOrder16(i, seed), 0x49c5f37eu);
      memcpy(out + k, &v, 4);
    }
  }
}

// camera/Camera.cpp:33-126 + OrthographicCamera.cpp:25-49 (no transform / motion blur)
ORACLE_API void oracle_commit_ortho_camera(const float pos[3], const float dir[3], const float up[3], float height, float aspect, OcCamera *out)
{
  const V3 d = normalize(v3(dir[0], dir[1], dir[2])), u = normalize(v3(up[0], up[1], up[2]));
  const V3 dn = normalize(d);
  V3 du = normalize(cross(dn, u));
  const V3 dv = cross(du, dn) * height;
  du = du * (height * aspect);
  const V3 org = v3(pos[0], pos[1], pos[2]) - 0.5f * du - 0.5f * dv;
  memset(out, 0, sizeof(*out));
  out->type = OC_CAMERA_ORTHOGRAPHIC;
  out->org = {org.x, org.y, org.z}; out->dir_00 = {dn.x, dn.y, dn.z}; out->du_size = {du.x, du.y, du.z};
  out->dv_up = {dv.x, dv.y, dv.z}; out->direction = {d.x, d.y, d.z};
  out->nearClip = 1e-6f; out->subImage[2] = 1.f; out->subImage[3] = 1.f;
  out->aspect = aspect; out->imgPlaneSizeY = height;
}

// ---- host-side commit restatements ---------------------------------------------------------------------
// pf/PixelFilter.cpp:23-40 + LUTPixelFilter.ispc:17-142 (Gaussian only: the default filter)
ORACLE_API void oracle_gauss_pixelfilter_lut(float width, int lut, float *cdfX, float *cdfY)
{
  for (int y = 0; y < lut; y++) {
    const float fy = (y + 0.5f) * (1.0f / (float)lut);
    for (int x = 0; x < lut; x++) {
      const float fx = (x + 0.5f) * (1.0f / (float)lut);
      const float cx = (fx - 0.5f) * width, cy = (fy - 0.5f) * width;
      const float sigma = (0.5f * width) / 3.0f;
      cdfX[y * lut + x] = fmaxf(0.0f, expf(-0.5f * ((cx * cx + cy * cy) / (sigma * sigma))));
    }
    cdfY[y] = Distribution1D_create(lut, cdfX + y * lut);
  }
  Distribution1D_create(lut, cdfY);
}
// camera/Camera.cpp:33-126 + PerspectiveCamera.cpp:26-117 (no transform / stereo / motion blur)
ORACLE_API void oracle_commit_camera_stereo(const float pos[3], const float dir[3], const float up[3], float fovy, float aspect,
    float apertureRadius, float focusDistance, int stereoMode, float interpupillaryDistance, OcCamera *out);
ORACLE_API void oracle_commit_camera(const float pos[3], const float dir[3], const float up[3], float fovy, float aspect,
    float apertureRadius, float focusDistance, OcCamera *out)
{
  oracle_commit_camera_stereo(pos, dir, up, fovy, aspect, apertureRadius, focusDistance, 0, 0.0635f, out);
}
// camera/PanoramicCamera.cpp:29-53
ORACLE_API void oracle_commit_pano_camera(const float pos[3], const float dir[3], const float up[3], int stereoMode, float interpupillaryDistance,
    OcCamera *out)
{
  const V3 d = normalize(v3(dir[0], dir[1], dir[2])), u = normalize(v3(up[0], up[1], up[2]));
  const V3 vz = -1.f * normalize(d);
  const V3 vx = normalize(cross(u, vz));
  const V3 vy = cross(vz, vx);
  memset(out, 0, sizeof(*out));
  out->type = OC_CAMERA_PANORAMIC;
  out->org = {pos[0], pos[1], pos[2]};
  out->du_size = {vx.x, vx.y, vx.z}; out->dv_up = {vy.x, vy.y, vy.z}; out->dir_00 = {vz.x, vz.y, vz.z}; out->direction = {d.x, d.y, d.z};
  out->nearClip = 1e-6f; out->subImage[2] = 1.f; out->subImage[3] = 1.f;
  out->stereoMode = (uint32_t)stereoMode;
  float ipd = 0.5f * interpupillaryDistance;
  if (stereoMode == 4) ipd = -ipd;
  out->ipdOffset = {ipd, 0.f, 0.f};
}
// + PerspectiveCamera.cpp:35-100: stereo modes (aspect of the halves, eye offset along du, LEFT / RIGHT folded into org)
ORACLE_API void oracle_commit_camera_stereo(const float pos[3], const float dir[3], const float up[3], float fovy, float aspect,
    float apertureRadius, float focusDistance, int stereoMode, float interpupillaryDistance, OcCamera *out)
{
  if (stereoMode == 3) aspect *= 0.5f;
  else if (stereoMode == 4) aspect *= 2.f;
  const V3 d = normalize(v3(dir[0], dir[1], dir[2])), u = normalize(v3(up[0], up[1], up[2]));
  const float imgY = 2.f * tanf((0.5f * fovy) * (float)(M_PI / 180.0)), imgX = imgY * aspect;
  const float scaledAperture = apertureRadius ? apertureRadius / (imgX * focusDistance) : 0.0f;
  V3 dir_00 = normalize(d);
  V3 du = normalize(cross(dir_00, u));
  V3 dv = cross(du, dir_00);
  V3 ipd = (0.5f * interpupillaryDistance) * du;
  V3 org = v3(pos[0], pos[1], pos[2]);
  if (stereoMode == 1) org = org - ipd;
  else if (stereoMode == 2) org = org + ipd;
  else if (stereoMode == 4) ipd = -1.f * ipd;
  du = du * imgX; dv = dv * imgY;
  dir_00 = dir_00 - 0.5f * du - 0.5f * dv;
  if (scaledAperture > 0.f) { du = du * focusDistance; dv = dv * focusDistance; dir_00 = dir_00 * focusDistance; }
  memset(out, 0, sizeof(*out));
  out->stereoMode = (uint32_t)stereoMode; out->ipdOffset = {ipd.x, ipd.y, ipd.z};
  out->org = {org.x, org.y, org.z}; out->dir_00 = {dir_00.x, dir_00.y, dir_00.z}; out->du_size = {du.x, du.y, du.z};
  out->dv_up = {dv.x, dv.y, dv.z}; out->direction = {d.x, d.y, d.z};
  out->nearClip = 1e-6f; out->subImage[0] = 0.f; out->subImage[1] = 0.f; out->subImage[2] = 1.f; out->subImage[3] = 1.f;
  out->scaledAperture = scaledAperture; out->aspect = aspect; out->imgPlaneSizeY = imgY; out->needLensSample = scaledAperture > 0.f;
}
// a light of a (static) transformed instance: QuadLight.ispc:14-33, PointLight.ispc:17-29, SpotLight.ispc:14-24 (Transform),
// DirectionalLight.cpp:36-44 (frame = xfm.l * frame); xfm = 12 floats (l.vx, l.vy, l.vz, p)
ORACLE_API void oracle_transform_light(OcLight *l, const float *x)
{
  A3 xfm;
  xfm.l.vx = v3(x[0], x[1], x[2]); xfm.l.vy = v3(x[3], x[4], x[5]); xfm.l.vz = v3(x[6], x[7], x[8]); xfm.p = v3(x[9], x[10], x[11]);
  auto oc = [](V3 v) { return OcVec3f{v.x, v.y, v.z}; };
  switch (l->type) {
  case OC_LIGHT_QUAD: {
    const V3 position = xfmPoint(xfm, v3(l->position)), e1 = xfmVector(xfm.l, v3(l->edge1)), e2 = xfmVector(xfm.l, v3(l->edge2));
    const V3 ndirection = cross(e2, e1);
    const float ppdf = 1.f / length(ndirection);
    l->position = oc(position); l->edge1 = oc(e1); l->edge2 = oc(e2); l->ppdf = ppdf; l->nnormal = oc(ndirection * ppdf);
    break;
  }
  case OC_LIGHT_SPHERE: l->position = oc(xfmPoint(xfm, v3(l->position))); break;
  case OC_LIGHT_CYLINDER: l->position = oc(xfmPoint(xfm, v3(l->position))); l->edge1 = oc(xfmPoint(xfm, v3(l->edge1))); break; // CylinderLight.ispc:12-19
  case OC_LIGHT_SPOT: { // fx / fy / fz = c0 / c90 / direction
    const V3 direction = normalize(xfmVector(xfm.l, v3(l->fz)));
    const V3 c90 = normalize(cross(xfmVector(xfm.l, v3(l->fx)), direction));
    l->position = oc(xfmPoint(xfm, v3(l->position)));
    l->fz = oc(direction); l->fy = oc(c90); l->fx = oc(cross(direction, c90));
    break;
  }
  case OC_LIGHT_DISTANT:
  case OC_LIGHT_HDRI: // HDRILight.cpp:52-55: light2world = xfm.l * light2world (world2light = its inverse, taken by the user of the struct)
    l->fx = oc(xfmVector(xfm.l, v3(l->fx))); l->fy = oc(xfmVector(xfm.l, v3(l->fy))); l->fz = oc(xfmVector(xfm.l, v3(l->fz)));
    break;
  default: break; // ambient
  }
}
// lights: Light.cpp:20-25, QuadLight.cpp:68-108 + QuadLight.ispc:15-30, AmbientLight.cpp, DirectionalLight.cpp
ORACLE_API void oracle_commit_light(int type, const float color[3], float intensity, int visible, const float position[3],
    const float edge1[3], const float edge2[3], const float direction[3], float angularDiameter, OcLight *out)
{
  memset(out, 0, sizeof(*out));
  const V3 ci = v3(color[0], color[1], color[2]) * intensity;
  out->isVisible = visible;
  out->type = type;
  if (type == OC_LIGHT_QUAD) {
    const V3 e1 = v3(edge1[0], edge1[1], edge1[2]), e2 = v3(edge2[0], edge2[1], edge2[2]);
    const V3 nd = cross(e2, e1);
    const float ppdf = 1.f / length(nd);
    const V3 nn = nd * ppdf;
    out->radiance = {ci.x, ci.y, ci.z}; // intensityQuantity RADIANCE (default)
    out->position = {position[0], position[1], position[2]};
    out->edge1 = {e1.x, e1.y, e1.z}; out->edge2 = {e2.x, e2.y, e2.z}; out->nnormal = {nn.x, nn.y, nn.z}; out->ppdf = ppdf;
  } else if (type == OC_LIGHT_AMBIENT) {
    out->radiance = {ci.x, ci.y, ci.z};
  } else {
    const V3 d = -normalize(v3(direction[0], direction[1], direction[2]));
    const float ad = fminf(fmaxf(angularDiameter, 0.f), 180.f);
    const float cosAngle = cosf((0.5f * ad) * (float)(M_PI / 180.0));
    const L3 f = frame(d);
    out->radiance = {ci.x, ci.y, ci.z}; // irradiance
    out->fx = {f.vx.x, f.vx.y, f.vx.z}; out->fy = {f.vy.x, f.vy.y, f.vy.z}; out->fz = {f.vz.x, f.vz.y, f.vz.z};
    out->cosAngle = cosAngle;
    out->pdf = cosAngle < 0.99999988f ? 1.f / (TWO_PI * (1.0f - cosAngle)) : INF;
  }
}

// lights/PointLight.cpp:77-124 ("sphere" light, default intensityQuantity INTENSITY, no intensity distribution)
ORACLE_API void oracle_commit_sphere_light(const float color[3], float intensity, int visible, const float position[3], float radius, OcLight *out)
{
  memset(out, 0, sizeof(*out));
  const V3 ci = v3(color[0], color[1], color[2]) * intensity;
  out->type = OC_LIGHT_SPHERE;
  out->isVisible = visible;
  out->position = {position[0], position[1], position[2]};
  out->radius = radius;
  out->intensity = {ci.x, ci.y, ci.z};
  if (radius > 0.0f) {
    const float sphereArea = 4.0f * M_PI * radius * radius; // double product, rounded once (as written in the reference)
    const V3 rad = ci / (sphereArea / 4.0f);
    out->radiance = {rad.x, rad.y, rad.z};
  }
}

// lights/SpotLight.cpp:22-128 (default intensityQuantity INTENSITY, no intensity distribution, no instance transform)
ORACLE_API void oracle_commit_spot_light(const float color[3], float intensity, int visible, const float position[3], const float direction[3],
    float openingAngle, float penumbraAngle, float radius, float innerRadius, OcLight *out)
{
  memset(out, 0, sizeof(*out));
  const V3 ci = v3(color[0], color[1], color[2]) * intensity;
  out->type = OC_LIGHT_SPOT;
  out->isVisible = visible;
  out->position = {position[0], position[1], position[2]};
  radius = fmaxf(0.0f, radius);
  innerRadius = fminf(fmaxf(innerRadius, 0.0f), 0.999f * radius);
  const V3 dp = v3(direction[0], direction[1], direction[2]);
  const V3 c0i = fabsf(dp.x) < fabsf(dp.y) ? v3(0.0f, dp.z, dp.y) : v3(dp.z, 0.0f, dp.x);
  const V3 d = normalize(dp);
  const V3 c90 = normalize(cross(c0i, d));
  const V3 c0 = cross(d, c90);
  out->fx = {c0.x, c0.y, c0.z}; out->fy = {c90.x, c90.y, c90.z}; out->fz = {d.x, d.y, d.z};
  openingAngle = fminf(fmaxf(openingAngle, 0.f), 360.f);
  penumbraAngle = fminf(fmaxf(penumbraAngle, 0.f), 0.5f * openingAngle);
  const float deg = 1.74532925199432957692e-2f; // rkcommon deg2rad
  out->cosAngle = cosf((0.5f * openingAngle) * deg);
  const float cosAngleMin = cosf((0.5f * openingAngle - penumbraAngle) * deg);
  out->cosAngleScale = 1.0f / (cosAngleMin - out->cosAngle);
  out->radius = radius; out->innerRadius = innerRadius;
  out->ppdf = 1.f / (PI * (sqr(radius) - sqr(innerRadius))); // uniformSampleRingPDF
  out->intensity = {ci.x, ci.y, ci.z};
  if (radius > 0.0f) {
    const float ringDiskArea = M_PI * (sqr(radius) - sqr(innerRadius));
    const V3 rad = ci / ringDiskArea;
    out->radiance = {rad.x, rad.y, rad.z};
  }
}

// ---- framebuffer ops ---------------------------------------------------------------------------------------
// LocalFB_accumulateSample applied to an externally produced sample image
ORACLE_API void oracle_fb_accumulate(const OcFrameBuffer *fb, const OcVec4f *sc, const float *sz)
{
  const size_t n = (size_t)fb->sizeX * fb->sizeY;
  const int accumID = fb->frameID;
  const float f = 1.f / (float)(accumID + 1), fv = 1.f / (float)(accumID / 2 + 1);
  for (size_t i = 0; i < n; i++) {
    const V4 s = {sc[i].x, sc[i].y, sc[i].z, sc[i].w};
    if (fb->color) accum4(fb->color, i, s, f);
    if (fb->accumulateVariance && fb->variance) accum4(fb->variance, i, s, fv);
    if (fb->depth && sz) if (accumID == 0 || fabsf(sz[i]) < fabsf(fb->depth[i])) fb->depth[i] = sz[i];
  }
}
// fb/frame_ops/Variance.ispc:11-59
ORACLE_API void oracle_fb_variance(const OcVec4f *color, const OcVec4f *variance, int sizeX, int sizeY, int taskSize, float *taskError)
{
  const int tasksX = (sizeX + taskSize - 1) / taskSize, tasksY = (sizeY + taskSize - 1) / taskSize;
  for (int ty = 0; ty < tasksY; ty++)
    for (int tx = 0; tx < tasksX; tx++) {
      float v = 0.f; int cnt = 0;
      for (int ry = 0; ry < taskSize; ry++) {
        const int y = ty * taskSize + ry;
        if (y >= sizeY) continue;
        for (int rx = 0; rx < taskSize; rx++) {
          const int x = tx * taskSize + rx;
          if (x >= sizeX) continue;
          const OcVec4f c = color[(size_t)y * sizeX + x], w = variance[(size_t)y * sizeX + x];
          const float den2 = (c.x + c.y + c.z) + (1.f - c.w);
          if (den2 > 0.f) { cnt++; v += (fabsf(c.x - w.x) + fabsf(c.y - w.y) + fabsf(c.z - w.z) + fabsf(c.w - w.w)) * (1.f / sqrtf(den2)); }
        }
      }
      if (cnt) v *= 1.f / (float)cnt;
      taskError[ty * tasksX + tx] = v * 64.f;
    }
}
// fb/frame_ops/ToneMapper.ispc:12-41
ORACLE_API void oracle_fb_tonemap(const OcVec4f *in, OcVec4f *out, uint32_t n, const OcToneMapper *tm)
{
  for (uint32_t i = 0; i < n; i++) {
    V3 x = v3(in[i].x, in[i].y, in[i].z) * tm->exposure;
    if (tm->acesColor)
      x = {0.5972782409f * x.x + 0.3545713181f * x.y + 0.0482176639f * x.z, 0.0760130499f * x.x + 0.9083220973f * x.y + 0.0156579968f * x.z,
          0.0284085382f * x.x + 0.1338243154f * x.y + 0.8375684636f * x.z};
    x = {powf(x.x, tm->a) / (powf(x.x, tm->a * tm->d) * tm->b + tm->c), powf(x.y, tm->a) / (powf(x.y, tm->a * tm->d) * tm->b + tm->c),
        powf(x.z, tm->a) / (powf(x.z, tm->a * tm->d) * tm->b + tm->c)};
    if (tm->acesColor)
      x = {1.6047539945f * x.x - 0.5310794927f * x.y - 0.0736720338f * x.z, -0.1020831870f * x.x + 1.1081322801f * x.y - 0.0060518756f * x.z,
          -0.0032670420f * x.x - 0.0727552477f * x.y + 1.0760219533f * x.z};
    out[i] = {clampf(x.x, 0.f, 1.f), clampf(x.y, 0.f, 1.f), clampf(x.z, 0.f, 1.f), in[i].w};
  }
}
// fb/frame_ops/ColorConversion.ispc:11-36
ORACLE_API void oracle_fb_color_convert(const OcVec4f *in, void *out, uint32_t n, uint32_t format)
{
  for (uint32_t i = 0; i < n; i++) {
    const OcVec4f v = in[i];
    if (format == 1) ((uint32_t *)out)[i] = cvt_uint32(v.x) | (cvt_uint32(v.y) << 8) | (cvt_uint32(v.z) << 16) | (cvt_uint32(v.w) << 24);
    else if (format == 2)
      ((uint32_t *)out)[i] = cvt_uint32(linear_to_srgb(v.x)) | (cvt_uint32(linear_to_srgb(v.y)) << 8) | (cvt_uint32(linear_to_srgb(v.z)) << 16)
          | (cvt_uint32(v.w) << 24);
    else if (format == 3) ((OcVec4f *)out)[i] = {fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fminf(fmaxf(v.w, 0.f), 1.f)};
  }
}
// LocalFrameBuffer::writeTiles (fb/LocalFB.cpp:286-353): copy the 64x64 tiles owned by srcRank
ORACLE_API void oracle_fb_gather_tiles(void *dst, const void *src, int sizeX, int sizeY, uint32_t bpp, uint32_t numRanks, uint32_t srcRank)
{
  const uint32_t tilesX = (sizeX + 63) / 64;
  for (int y = 0; y < sizeY; y++)
    for (int x = 0; x < sizeX; x++) {
      const uint32_t tileID = (uint32_t)(y / 64) * tilesX + (uint32_t)(x / 64);
      if (tileID % numRanks != srcRank) continue;
      const size_t off = ((size_t)y * sizeX + x) * bpp;
      memcpy((char *)dst + off, (const char *)src + off, bpp);
    }
}
ORACLE_API int oracle_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }
// tile-major staging of the multi-device gather (pack != 0: image -> staging, else staging -> image);
// layout contract of ospcuda_fb_pack_tiles / ospcuda_fb_unpack_tiles (include/ospcuda.h)
ORACLE_API void oracle_fb_pack_tiles(void *packed, void *image, int sizeX, int sizeY, uint32_t bpp, uint32_t numRanks, uint32_t rank, int pack)
{
  const uint32_t tilesX = (sizeX + 63) / 64, tilesY = (sizeY + 63) / 64, total = tilesX * tilesY;
  uint32_t k = 0;
  for (uint32_t tileID = rank; tileID < total; tileID += numRanks, k++) {
    const int x0 = (tileID % tilesX) * 64, y0 = (tileID / tilesX) * 64;
    const int w = std::min(64, sizeX - x0), h = std::min(64, sizeY - y0);
    for (int r = 0; r < h; r++) {
      char *a = (char *)packed + ((size_t)k * 64 + r) * 64 * bpp;
      char *b = (char *)image + ((size_t)(y0 + r) * sizeX + x0) * bpp;
      if (pack) memcpy(a, b, (size_t)w * bpp); else memcpy(b, a, (size_t)w * bpp);
    }
  }
}

// ---- 2D textures: texel formats and the MIP chain ------------------------------------------------------------------
// texture/Texture2DFormats.ispc (getTexel_* :216-400, setTexel_* :46-214) and texture/MipMapGeneration.ispc:8-95 restated;
// the chain layout is Texture2D::commit's (texture/Texture2D.cpp:103-141: level l+1 = max(size_l / 2, 1) down to 1x1).
// The checker builds the chain with this code and hands it to the reference's own Texture2D_get (oracle/_ref); the device
// builds its own at ospCommit. sRGB transfer = rkcommon's (un-vendored, see the header of this file).
namespace {
struct TexFmt { int chan; int kind; bool srgb; }; // kind: 0 u8, 1 u16, 2 f32, 3 f16; chan 1 R, 2 RA, 3 RGB, 4 RGBA
bool texFmt(uint32_t f, TexFmt &t)
{
  // OSPTextureFormat order (OSPEnums.h): RGBA8 SRGBA RGBA32F RGB8 SRGB RGB32F R8 R32F L8 RA8 LA8 RGBA16 RGB16 RA16 R16 RA32F RGBA16F RGB16F RA16F R16F
  static const TexFmt tab[20] = {{4, 0, false}, {4, 0, true}, {4, 2, false}, {3, 0, false}, {3, 0, true}, {3, 2, false}, {1, 0, false}, {1, 2, false},
      {1, 0, true}, {2, 0, false}, {2, 0, true}, {4, 1, false}, {3, 1, false}, {2, 1, false}, {1, 1, false}, {2, 2, false}, {4, 3, false},
      {3, 3, false}, {2, 3, false}, {1, 3, false}};
  if (f >= 20) return false;
  t = tab[f];
  return true;
}
inline size_t texChanBytes(int kind) { return kind == 0 ? 1 : kind == 2 ? 4 : 2; }
inline float texSrgbDecode(float x) { return x <= 0.04045f ? x * (1.f / 12.92f) : powf((x + 0.055f) * (1.f / 1.055f), 2.4f); }
inline float texSrgbEncode(float f) { const float c = std::max(f, 0.f); return c <= 0.0031308f ? 12.92f * c : powf(c, 1.f / 2.4f) * 1.055f - 0.055f; }
inline float texLoadChan(const TexFmt &t, const void *data, size_t i)
{
  switch (t.kind) {
  case 0: return (float)((const uint8_t *)data)[i] * (1.f / 255.f);
  case 1: return (float)((const uint16_t *)data)[i] * (1.f / 65535.f);
  case 2: return ((const float *)data)[i];
  default: { _Float16 h; memcpy(&h, (const uint16_t *)data + i, 2); return (float)h; }
  }
}
inline void texStoreChan(const TexFmt &t, void *data, size_t i, float v)
{
  switch (t.kind) {
  case 0: ((uint8_t *)data)[i] = (uint8_t)rintf(255.f * std::min(std::max(v, 0.f), 1.f)); break;
  case 1: ((uint16_t *)data)[i] = (uint16_t)rintf(65535.f * std::min(std::max(v, 0.f), 1.f)); break;
  case 2: ((float *)data)[i] = v; break;
  default: { const _Float16 h = (_Float16)v; memcpy((uint16_t *)data + i, &h, 2); }
  }
}
// vec4f of texel idx: missing channels read (r, 0, 0, 1) / (r, 0, 0, a) / (r, g, b, 1); luminance formats replicate
V4 texGet(uint32_t f, const TexFmt &t, const void *data, uint32_t idx)
{
  const size_t b = (size_t)idx * t.chan;
  V4 v{0.f, 0.f, 0.f, 1.f};
  v.x = texLoadChan(t, data, b);
  if (t.chan == 2) v.w = texLoadChan(t, data, b + 1);
  if (t.chan >= 3) { v.y = texLoadChan(t, data, b + 1); v.z = texLoadChan(t, data, b + 2); }
  if (t.chan == 4) v.w = texLoadChan(t, data, b + 3);
  if (t.srgb) {
    v.x = texSrgbDecode(v.x);
    if (t.chan >= 3) { v.y = texSrgbDecode(v.y); v.z = texSrgbDecode(v.z); }
    else v.y = v.z = v.x; // L8 / LA8
  }
  (void)f;
  return v;
}
void texSet(const TexFmt &t, void *data, uint32_t idx, V4 v)
{
  const size_t b = (size_t)idx * t.chan;
  if (t.srgb) { v.x = texSrgbEncode(v.x); v.y = texSrgbEncode(v.y); v.z = texSrgbEncode(v.z); if (t.chan == 4) v.w = std::max(v.w, 0.f); }
  texStoreChan(t, data, b, v.x);
  if (t.chan == 2) texStoreChan(t, data, b + 1, v.w);
  if (t.chan >= 3) { texStoreChan(t, data, b + 1, v.y); texStoreChan(t, data, b + 2, v.z); }
  if (t.chan == 4) texStoreChan(t, data, b + 3, v.w);
}
inline V4 tex3(float a, V4 x, float b, V4 y, float c, V4 z)
{
  return {a * x.x + b * y.x + c * z.x, a * x.y + b * y.y + c * z.y, a * x.z + b * y.z + c * z.z, a * x.w + b * y.w + c * z.w};
}
} // namespace

// bytes of one texel, 0 for an unknown format
ORACLE_API uint32_t oracle_texture_texel_bytes(uint32_t format)
{
  TexFmt t;
  return texFmt(format, t) ? (uint32_t)(t.chan * texChanBytes(t.kind)) : 0u;
}
// one MIP level from the finer one (MipMap_generate)
ORACLE_API void oracle_texture_mip_level(uint32_t format, const void *src, int sx, int sy, void *dst, int dx, int dy)
{
  TexFmt t;
  if (!texFmt(format, t)) return;
  if (sx > 1 && sy > 1) {
    for (int y = 0; y < dy; y++) { // MipMap_generateRows, one task per destination row
      const int dsty = y * dx, srcy = 2 * y * sx;
      float rx = 1.f / (float)sx, ry = 1.f / (float)sy, hx = (float)dx * rx, hy = (float)dy * ry;
      int ox = 2, oy = srcy + 2 * sx;
      if (!(sx & 1)) { hx = 0.5f; rx = 0.f; ox = 0; }
      if (!(sy & 1)) { hy = 0.5f; ry = 0.f; oy = srcy; }
      const float wy0 = hy - y * ry, wy2 = ry + y * ry;
      for (int x = 0; x < dx; x++) {
        const float tx = x * rx;
        const int s = 2 * x;
        V4 v[3];
        const int rows[3] = {srcy, srcy + sx, oy};
        for (int k = 0; k < 3; k++)
          v[k] = tex3(hx - tx, texGet(format, t, src, rows[k] + s), hx, texGet(format, t, src, rows[k] + s + 1), rx + tx, texGet(format, t, src, rows[k] + s + ox));
        texSet(t, dst, dsty + x, tex3(wy0, v[0], hy, v[1], wy2, v[2]));
      }
    }
  } else if (sx > 1 || sy > 1) { // MipMap_generateRow
    const int sn = std::max(sx, sy), dn = std::max(dx, dy);
    float r = 1.f / (float)sn, h = (float)dn * r;
    int o = 2;
    if (!(sn & 1)) { h = 0.5f; r = 0.f; o = 0; }
    for (int x = 0; x < dn; x++) {
      const float tx = x * r;
      texSet(t, dst, x, tex3(h - tx, texGet(format, t, src, 2 * x), h, texGet(format, t, src, 2 * x + 1), r + tx, texGet(format, t, src, 2 * x + o)));
    }
  }
}
