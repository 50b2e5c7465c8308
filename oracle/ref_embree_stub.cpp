// ref_embree_stub.cpp -- brute-force stand-in for the Embree 4.3.3 scene queries the reference's kernels call
// (rtcIntersect1 / rtcOccluded1 / rtcGetGeometryTransformFromScene; call sites modules/cpu/common/World.ih:29-52,
// :311-335, Instance.ih:26-40).  TEST INFRASTRUCTURE ONLY: part of oracle/_ref/libospref.so, never linked by
// the product.
//
// Embree is NOT vendored in /root/reference.  What is restated here from Embree's published kernels is the same
// arithmetic oracle.cpp documents (Moeller-Trumbore with e1 = v0-v1, e2 = v2-v0 and fused multiply-adds, the quad
// split (v0,v1,v3)+(v2,v3,v1) with (1-u,1-v), the sphere test with projected centre).  User geometries (boxes,
// volumes) are NOT restated: like Embree, the stub calls back into the reference's own intersect kernels through
// RTCIntersectArguments::intersect / RTCOccludedArguments::occluded with the ray in instance space.
//
// Scene description: an RTCScene wraps the OcWorld the test passes in (include/ospcuda.h PODs, HOST pointers)
// plus, per (instance, model|volume), the geometryUserPtr the reference's callbacks expect.
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include "ref_embree_stub.h"
#include "curve_capsule.h"

namespace {

struct V3 { float x, y, z; };
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dotF(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline V3 crossF(V3 a, V3 b) { return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))}; }
inline float xorSign(float f, uint32_t sgn) { uint32_t b; memcpy(&b, &f, 4); b ^= sgn; memcpy(&f, &b, 4); return f; }
inline V3 get3f(const OcData1D &d, uint32_t i) { const float *p = (const float *)(d.addr + (int64_t)i * d.byteStride); return {p[0], p[1], p[2]}; }
inline uint32_t get1ui(const OcData1D &d, uint32_t i, int c) { return ((const uint32_t *)(d.addr + (int64_t)i * d.byteStride))[c]; }

// Embree 4 TriangleM / Moeller-Trumbore (kernels/geometry/triangle_intersector_moeller.h)
inline bool intersectTriangle(const RTCRay &ray, V3 v0, V3 e1, V3 e2, float &t, float &u, float &v, V3 &Ng)
{
  const V3 org = {ray.org_x, ray.org_y, ray.org_z}, dir = {ray.dir_x, ray.dir_y, ray.dir_z};
  const V3 triNg = crossF(e2, e1);
  const V3 C = v0 - org;
  const V3 R = crossF(C, dir);
  const float den = dotF(triNg, dir);
  const float absDen = fabsf(den);
  uint32_t sgn; memcpy(&sgn, &den, 4); sgn &= 0x80000000u;
  const float U = xorSign(dotF(R, e2), sgn), V = xorSign(dotF(R, e1), sgn);
  if (!(den != 0.f && U >= 0.f && V >= 0.f && U + V <= absDen)) return false;
  const float T = xorSign(dotF(triNg, C), sgn);
  if (!(absDen * ray.tnear < T && T <= absDen * ray.tfar)) return false;
  const float rcpAbsDen = 1.f / absDen;
  u = U * rcpAbsDen; v = V * rcpAbsDen; t = T * rcpAbsDen; Ng = triNg;
  return true;
}
// Embree 4 sphere points (kernels/geometry/sphere_intersector.h)
inline bool intersectSphere(const RTCRay &ray, V3 c, float r, float &t, V3 &Ng)
{
  const V3 org = {ray.org_x, ray.org_y, ray.org_z}, dir = {ray.dir_x, ray.dir_y, ray.dir_z};
  const float rd2 = 1.f / dotF(dir, dir);
  const V3 c0 = c - org;
  const float projC0 = dotF(c0, dir) * rd2;
  const V3 perp = {fmaf(-projC0, dir.x, c0.x), fmaf(-projC0, dir.y, c0.y), fmaf(-projC0, dir.z, c0.z)};
  const float l2 = dotF(perp, perp), r2 = r * r;
  if (!(l2 <= r2)) return false;
  const float td = sqrtf((r2 - l2) * rd2);
  const float tFront = projC0 - td, tBack = projC0 + td;
  float tdSel;
  if (ray.tnear <= tFront && tFront <= ray.tfar) { t = tFront; tdSel = -td; }
  else if (ray.tnear <= tBack && tBack <= ray.tfar) { t = tBack; tdSel = td; }
  else return false;
  Ng = {fmaf(tdSel, dir.x, -perp.x), fmaf(tdSel, dir.y, -perp.y), fmaf(tdSel, dir.z, -perp.z)};
  return true;
}

inline bool hitTri(RTCRayHit &rh, V3 a, V3 b, V3 c, bool second, uint32_t primID, uint32_t geomID, uint32_t instID, bool occl)
{
  float t, u, v; V3 Ng;
  if (!intersectTriangle(rh.ray, a, a - b, c - a, t, u, v, Ng)) return false;
  rh.ray.tfar = t;
  if (!occl) {
    rh.hit.u = second ? 1.f - u : u; rh.hit.v = second ? 1.f - v : v;
    rh.hit.Ng_x = Ng.x; rh.hit.Ng_y = Ng.y; rh.hit.Ng_z = Ng.z;
    rh.hit.primID = primID; rh.hit.geomID = geomID; rh.hit.instID[0] = instID;
  }
  return true;
}

struct Callbacks { RTCIntersectFunctionN intersect; RTCOccludedFunctionN occluded; RTCRayQueryContext *context; };

// one user primitive: Embree hands the callback the ray in instance space, valid = -1, and expects the callback to
// shorten ray.tfar / fill the hit itself (RTCIntersectFunctionNArguments)
inline bool userPrim(RTCRayHit &rh, void *userPtr, uint32_t primID, uint32_t geomID, const Callbacks &cb, bool occl)
{
  int valid = -1;
  if (occl) {
    if (!cb.occluded) return false;
    RTCOccludedFunctionNArguments a;
    a.valid = &valid; a.geometryUserPtr = userPtr; a.primID = primID; a.context = cb.context;
    a.ray = (RTCRayN *)&rh; a.N = 1; a.geomID = geomID;
    const float before = rh.ray.tfar;
    cb.occluded(&a);
    return rh.ray.tfar < before && rh.ray.tfar < rh.ray.tnear; // occlusion = tfar set to -inf
  }
  if (!cb.intersect) return false;
  RTCIntersectFunctionNArguments a;
  a.valid = &valid; a.geometryUserPtr = userPtr; a.primID = primID; a.context = cb.context;
  a.rayhit = (RTCRayHitN *)&rh; a.N = 1; a.geomID = geomID;
  const float before = rh.ray.tfar;
  cb.intersect(&a);
  return rh.ray.tfar != before;
}

// all primitives of one group in (model, prim) order; ray in instance space
bool traceGroup(const RTCSceneTy *s, uint32_t instIdx, RTCRayHit &rh, const Callbacks &cb, bool occl)
{
  const OcGroup *g = s->world->instances[instIdx].group;
  bool any = false;
  if (s->volumes) {
    for (uint32_t v = 0; v < g->numVolumes; v++)
      if (userPrim(rh, s->userPtr[instIdx][v], 0, v, cb, occl)) { any = true; if (occl) return true; }
    return any;
  }
  for (uint32_t m = 0; m < g->numModels; m++) {
    const OcGeometry *geo = g->models[m].geom;
    for (uint32_t p = 0; p < geo->numPrims; p++) {
      bool h = false;
      if (geo->type == OC_GEOM_TRIANGLES) {
        const V3 a = get3f(geo->vertex, get1ui(geo->index, p, 0)), b = get3f(geo->vertex, get1ui(geo->index, p, 1)),
                 c = get3f(geo->vertex, get1ui(geo->index, p, 2));
        h = hitTri(rh, a, b, c, false, p, m, instIdx, occl);
      } else if (geo->type == OC_GEOM_QUADS) {
        const V3 q0 = get3f(geo->vertex, get1ui(geo->index, p, 0)), q1 = get3f(geo->vertex, get1ui(geo->index, p, 1)),
                 q2 = get3f(geo->vertex, get1ui(geo->index, p, 2)), q3 = get3f(geo->vertex, get1ui(geo->index, p, 3));
        h = hitTri(rh, q0, q1, q3, false, p, m, instIdx, occl);
        if (h && occl) return true;
        h |= hitTri(rh, q2, q3, q1, true, p, m, instIdx, occl);
      } else if (geo->type == OC_GEOM_SPHERES) {
        const float *sp = (const float *)(geo->sphere.addr + (int64_t)p * geo->sphere.byteStride);
        float t; V3 Ng;
        if (intersectSphere(rh.ray, {sp[0], sp[1], sp[2]}, sp[3], t, Ng)) {
          h = true; rh.ray.tfar = t;
          if (!occl) {
            rh.hit.u = rh.hit.v = 0.f; rh.hit.Ng_x = Ng.x; rh.hit.Ng_y = Ng.y; rh.hit.Ng_z = Ng.z;
            rh.hit.primID = p; rh.hit.geomID = m; rh.hit.instID[0] = instIdx;
          }
        }
      } else if (geo->type == OC_GEOM_CURVES) { // round curve = 8 capsules per segment (curve_capsule.h)
        const uint32_t first = get1ui(geo->index, p, 0);
        curvecap::F4 cv[4] = {};
        for (int k = 0; k < (geo->curveBasis == OC_CURVE_BASIS_LINEAR ? 2 : 4); k++) {
          const float *c = (const float *)(geo->sphere.addr + (int64_t)(first + k) * geo->sphere.byteStride);
          cv[k] = {c[0], c[1], c[2], c[3]};
        }
        const int basis = (int)geo->curveBasis;
        for (uint32_t k = 0; k < curvecap::CURVE_SUBDIV; k++) {
          const float u0 = (float)k * (1.f / curvecap::CURVE_SUBDIV), u1 = (float)(k + 1) * (1.f / curvecap::CURVE_SUBDIV);
          const curvecap::F4 a4 = curvecap::curveEval(basis, cv[0], cv[1], cv[2], cv[3], u0), b4 = curvecap::curveEval(basis, cv[0], cv[1], cv[2], cv[3], u1);
          const float rad = fmaxf(0.f, 0.5f * (a4.w + b4.w));
          float t, sPar; curvecap::F3 N;
          if (curvecap::intersectCapsule({rh.ray.org_x, rh.ray.org_y, rh.ray.org_z}, {rh.ray.dir_x, rh.ray.dir_y, rh.ray.dir_z}, rh.ray.tnear,
                  rh.ray.tfar, {a4.x, a4.y, a4.z}, {b4.x, b4.y, b4.z}, rad, t, sPar, N)) {
            h = true; rh.ray.tfar = t;
            if (occl) break;
            rh.hit.u = u0 + (1.f / curvecap::CURVE_SUBDIV) * sPar; rh.hit.v = 0.f; rh.hit.Ng_x = N.x; rh.hit.Ng_y = N.y; rh.hit.Ng_z = N.z;
            rh.hit.primID = p; rh.hit.geomID = m; rh.hit.instID[0] = instIdx;
          }
        }
      } else {
        h = userPrim(rh, s->userPtr[instIdx][m], p, m, cb, occl);
      }
      if (h) { any = true; if (occl) return true; }
    }
  }
  return any;
}

bool traceScene(const RTCSceneTy *s, RTCRayHit &rh, const Callbacks &cb, bool occl)
{
  const OcWorld *w = s->world;
  bool any = false;
  for (uint32_t i = 0; i < w->numInstances; i++) {
    const OcInstance *inst = w->instances + i;
    if ((s->volumes ? inst->group->numVolumes : inst->group->numModels) == 0) continue;
    cb.context->instID[0] = i;
    RTCRayHit local = rh;
    if (!inst->isIdentity) {
      const OcAffine3f &x = inst->rcp_xfm;
      const float ox = rh.ray.org_x, oy = rh.ray.org_y, oz = rh.ray.org_z, dx = rh.ray.dir_x, dy = rh.ray.dir_y, dz = rh.ray.dir_z;
      // xfmPoint / xfmVector in the oracle's operation order: (x*vx + y*vy + z*vz) + p
      local.ray.org_x = (ox * x.vx.x + oy * x.vy.x + oz * x.vz.x) + x.p.x;
      local.ray.org_y = (ox * x.vx.y + oy * x.vy.y + oz * x.vz.y) + x.p.y;
      local.ray.org_z = (ox * x.vx.z + oy * x.vy.z + oz * x.vz.z) + x.p.z;
      local.ray.dir_x = dx * x.vx.x + dy * x.vy.x + dz * x.vz.x;
      local.ray.dir_y = dx * x.vx.y + dy * x.vy.y + dz * x.vz.y;
      local.ray.dir_z = dx * x.vx.z + dy * x.vy.z + dz * x.vz.z;
    }
    if (traceGroup(s, i, local, cb, occl)) {
      any = true;
      rh.ray.tfar = local.ray.tfar;
      if (occl) break;
      rh.hit = local.hit;
    }
    cb.context->instID[0] = RTC_INVALID_GEOMETRY_ID;
  }
  cb.context->instID[0] = RTC_INVALID_GEOMETRY_ID;
  return any;
}

struct ThreadCounters { std::atomic<uint64_t> intersect{0}, occluded{0}; };
std::mutex gCountersMutex;
std::vector<ThreadCounters *> gCounters;
ThreadCounters *myCounters()
{
  thread_local ThreadCounters *c = nullptr;
  if (!c) {
    c = new ThreadCounters;
    std::lock_guard<std::mutex> l(gCountersMutex);
    gCounters.push_back(c);
  }
  return c;
}

} // namespace

extern "C" void rtcIntersect1(RTCScene scene, RTCRayHit *rayhit, RTCIntersectArguments *args)
{
  if (!scene->volumes) myCounters()->intersect.fetch_add(1, std::memory_order_relaxed); // volume-box queries are not BVH rays
  RTCRayQueryContext dummy; rtcInitRayQueryContext(&dummy);
  Callbacks cb{args ? args->intersect : nullptr, nullptr, (args && args->context) ? args->context : &dummy};
  traceScene(scene, *rayhit, cb, false);
}

extern "C" void rtcOccluded1(RTCScene scene, RTCRay *ray, RTCOccludedArguments *args)
{
  if (!scene->volumes) myCounters()->occluded.fetch_add(1, std::memory_order_relaxed);
  RTCRayQueryContext dummy; rtcInitRayQueryContext(&dummy);
  Callbacks cb{nullptr, args ? args->occluded : nullptr, (args && args->context) ? args->context : &dummy};
  RTCRayHit rh; rh.ray = *ray; memset(&rh.hit, 0, sizeof(rh.hit));
  rh.hit.geomID = rh.hit.primID = rh.hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
  if (traceScene(scene, rh, cb, true)) ray->tfar = -INFINITY;
}

// common/Instance.ih:26-40: only ever asked for the instance transform; the reference stores xfm on the host side
extern "C" void rtcGetGeometryTransformFromScene(RTCScene scene, unsigned int geomID, float, RTCFormat, void *xfm)
{
  memcpy(xfm, &scene->world->instances[geomID].xfm, sizeof(OcAffine3f));
}
extern "C" void rtcInterpolate0(RTCGeometry, unsigned int, float, float, RTCBufferType, unsigned int, float *, unsigned int) {}
extern "C" void rtcInterpolate1(RTCGeometry, unsigned int, float, float, RTCBufferType, unsigned int, float *, float *, float *, unsigned int) {}

extern "C" void ospref_ray_counters(uint64_t out[2], int reset)
{
  std::lock_guard<std::mutex> l(gCountersMutex);
  out[0] = out[1] = 0;
  for (auto *c : gCounters) {
    out[0] += c->intersect.load(); out[1] += c->occluded.load();
    if (reset) { c->intersect = 0; c->occluded = 0; }
  }
}
