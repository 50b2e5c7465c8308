// Stand-in for Embree 4.3.3's <embree4/rtcore.h> (NOT vendored in /root/reference): only the declarations the
// reference's kernel sources touch (SURVEY.md 8c lists the call sites).  The scene queries are implemented
// brute force in oracle/ref_embree_stub.cpp from Embree's published intersectors (the same restatement
// oracle.cpp documents); user-geometry primitives call back into the REFERENCE's own intersect kernels through
// RTCIntersectArguments::intersect exactly as Embree does.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>

#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1
#define RTC_SYCL_INDIRECTLY_CALLABLE
#define RTC_NAMESPACE_USE

enum RTCFeatureFlags : unsigned int
{
  RTC_FEATURE_FLAG_NONE = 0,
  RTC_FEATURE_FLAG_MOTION_BLUR = 1 << 0,
  RTC_FEATURE_FLAG_TRIANGLE = 1 << 1,
  RTC_FEATURE_FLAG_QUAD = 1 << 2,
  RTC_FEATURE_FLAG_GRID = 1 << 3,
  RTC_FEATURE_FLAG_SUBDIVISION = 1 << 4,
  RTC_FEATURE_FLAG_CONE_LINEAR_CURVE = 1 << 5,
  RTC_FEATURE_FLAG_ROUND_LINEAR_CURVE = 1 << 6,
  RTC_FEATURE_FLAG_FLAT_LINEAR_CURVE = 1 << 7,
  RTC_FEATURE_FLAG_ROUND_BEZIER_CURVE = 1 << 8,
  RTC_FEATURE_FLAG_FLAT_BEZIER_CURVE = 1 << 9,
  RTC_FEATURE_FLAG_NORMAL_ORIENTED_BEZIER_CURVE = 1 << 10,
  RTC_FEATURE_FLAG_ROUND_BSPLINE_CURVE = 1 << 11,
  RTC_FEATURE_FLAG_FLAT_BSPLINE_CURVE = 1 << 12,
  RTC_FEATURE_FLAG_NORMAL_ORIENTED_BSPLINE_CURVE = 1 << 13,
  RTC_FEATURE_FLAG_ROUND_HERMITE_CURVE = 1 << 14,
  RTC_FEATURE_FLAG_FLAT_HERMITE_CURVE = 1 << 15,
  RTC_FEATURE_FLAG_NORMAL_ORIENTED_HERMITE_CURVE = 1 << 16,
  RTC_FEATURE_FLAG_ROUND_CATMULL_ROM_CURVE = 1 << 17,
  RTC_FEATURE_FLAG_FLAT_CATMULL_ROM_CURVE = 1 << 18,
  RTC_FEATURE_FLAG_NORMAL_ORIENTED_CATMULL_ROM_CURVE = 1 << 19,
  RTC_FEATURE_FLAG_LINEAR_CURVES = RTC_FEATURE_FLAG_CONE_LINEAR_CURVE | RTC_FEATURE_FLAG_ROUND_LINEAR_CURVE | RTC_FEATURE_FLAG_FLAT_LINEAR_CURVE,
  RTC_FEATURE_FLAG_BEZIER_CURVES = RTC_FEATURE_FLAG_ROUND_BEZIER_CURVE | RTC_FEATURE_FLAG_FLAT_BEZIER_CURVE | RTC_FEATURE_FLAG_NORMAL_ORIENTED_BEZIER_CURVE,
  RTC_FEATURE_FLAG_BSPLINE_CURVES = RTC_FEATURE_FLAG_ROUND_BSPLINE_CURVE | RTC_FEATURE_FLAG_FLAT_BSPLINE_CURVE | RTC_FEATURE_FLAG_NORMAL_ORIENTED_BSPLINE_CURVE,
  RTC_FEATURE_FLAG_HERMITE_CURVES = RTC_FEATURE_FLAG_ROUND_HERMITE_CURVE | RTC_FEATURE_FLAG_FLAT_HERMITE_CURVE | RTC_FEATURE_FLAG_NORMAL_ORIENTED_HERMITE_CURVE,
  RTC_FEATURE_FLAG_CATMULL_ROM_CURVES = RTC_FEATURE_FLAG_ROUND_CATMULL_ROM_CURVE | RTC_FEATURE_FLAG_FLAT_CATMULL_ROM_CURVE | RTC_FEATURE_FLAG_NORMAL_ORIENTED_CATMULL_ROM_CURVE,
  RTC_FEATURE_FLAG_CURVES = RTC_FEATURE_FLAG_LINEAR_CURVES | RTC_FEATURE_FLAG_BEZIER_CURVES | RTC_FEATURE_FLAG_BSPLINE_CURVES | RTC_FEATURE_FLAG_HERMITE_CURVES | RTC_FEATURE_FLAG_CATMULL_ROM_CURVES,
  RTC_FEATURE_FLAG_SPHERE_POINT = 1 << 20,
  RTC_FEATURE_FLAG_DISC_POINT = 1 << 21,
  RTC_FEATURE_FLAG_ORIENTED_DISC_POINT = 1 << 22,
  RTC_FEATURE_FLAG_POINT = RTC_FEATURE_FLAG_SPHERE_POINT | RTC_FEATURE_FLAG_DISC_POINT | RTC_FEATURE_FLAG_ORIENTED_DISC_POINT,
  RTC_FEATURE_FLAG_INSTANCE = 1 << 23,
  RTC_FEATURE_FLAG_FILTER_FUNCTION_IN_ARGUMENTS = 1 << 24,
  RTC_FEATURE_FLAG_FILTER_FUNCTION_IN_GEOMETRY = 1 << 25,
  RTC_FEATURE_FLAG_FILTER_FUNCTION = RTC_FEATURE_FLAG_FILTER_FUNCTION_IN_ARGUMENTS | RTC_FEATURE_FLAG_FILTER_FUNCTION_IN_GEOMETRY,
  RTC_FEATURE_FLAG_USER_GEOMETRY_CALLBACK_IN_ARGUMENTS = 1 << 26,
  RTC_FEATURE_FLAG_USER_GEOMETRY_CALLBACK_IN_GEOMETRY = 1 << 27,
  RTC_FEATURE_FLAG_USER_GEOMETRY = RTC_FEATURE_FLAG_USER_GEOMETRY_CALLBACK_IN_ARGUMENTS | RTC_FEATURE_FLAG_USER_GEOMETRY_CALLBACK_IN_GEOMETRY,
  RTC_FEATURE_FLAG_32_BIT_RAY_MASK = 1 << 28,
  RTC_FEATURE_FLAG_INSTANCE_ARRAY = 1 << 29,
  RTC_FEATURE_FLAG_ALL = 0xffffffff,
};
enum RTCRayQueryFlags { RTC_RAY_QUERY_FLAG_NONE = 0, RTC_RAY_QUERY_FLAG_INVOKE_ARGUMENT_FILTER = 1 << 1, RTC_RAY_QUERY_FLAG_INCOHERENT = 0, RTC_RAY_QUERY_FLAG_COHERENT = 1 << 16 };
enum RTCBufferType { RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1, RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE = 2 };
enum RTCFormat { RTC_FORMAT_UNDEFINED = 0, RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR = 0x9244 };

struct RTCRay { float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar; unsigned int mask, id, flags; };
struct RTCHit { float Ng_x, Ng_y, Ng_z, u, v; unsigned int primID, geomID, instID[RTC_MAX_INSTANCE_LEVEL_COUNT]; };
struct RTCRayHit { RTCRay ray; RTCHit hit; };
struct RTCRayN;
struct RTCHitN;
struct RTCRayQueryContext { unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT]; };
inline void rtcInitRayQueryContext(RTCRayQueryContext *c) { c->instID[0] = RTC_INVALID_GEOMETRY_ID; }

typedef struct RTCSceneTy *RTCScene;
typedef struct RTCGeometryTy *RTCGeometry;

struct RTCBounds { float lower_x, lower_y, lower_z, align0, upper_x, upper_y, upper_z, align1; };
struct RTCBoundsFunctionArguments { void *geometryUserPtr; unsigned int primID; unsigned int timeStep; RTCBounds *bounds_o; };
struct RTCRayHitN;
struct RTCIntersectFunctionNArguments
{
  int *valid; void *geometryUserPtr; unsigned int primID; RTCRayQueryContext *context; RTCRayHitN *rayhit;
  unsigned int N; unsigned int geomID;
};
struct RTCOccludedFunctionNArguments
{
  int *valid; void *geometryUserPtr; unsigned int primID; RTCRayQueryContext *context; RTCRayN *ray;
  unsigned int N; unsigned int geomID;
};
struct RTCFilterFunctionNArguments
{
  int *valid; void *geometryUserPtr; RTCRayQueryContext *context; RTCRayN *ray; RTCHitN *hit; unsigned int N;
};
typedef void (*RTCIntersectFunctionN)(const RTCIntersectFunctionNArguments *args);
typedef void (*RTCOccludedFunctionN)(const RTCOccludedFunctionNArguments *args);
typedef void (*RTCFilterFunctionN)(const RTCFilterFunctionNArguments *args);

struct RTCIntersectArguments
{
  RTCRayQueryFlags flags; RTCFeatureFlags feature_mask; RTCRayQueryContext *context; RTCFilterFunctionN filter;
  RTCIntersectFunctionN intersect;
};
struct RTCOccludedArguments
{
  RTCRayQueryFlags flags; RTCFeatureFlags feature_mask; RTCRayQueryContext *context; RTCFilterFunctionN filter;
  RTCOccludedFunctionN occluded;
};
inline void rtcInitIntersectArguments(RTCIntersectArguments *a)
{
  a->flags = RTC_RAY_QUERY_FLAG_INCOHERENT; a->feature_mask = RTC_FEATURE_FLAG_ALL; a->context = nullptr; a->filter = nullptr; a->intersect = nullptr;
}
inline void rtcInitOccludedArguments(RTCOccludedArguments *a)
{
  a->flags = RTC_RAY_QUERY_FLAG_INCOHERENT; a->feature_mask = RTC_FEATURE_FLAG_ALL; a->context = nullptr; a->filter = nullptr; a->occluded = nullptr;
}

// implemented in oracle/ref_embree_stub.cpp
extern "C" {
void rtcIntersect1(RTCScene scene, RTCRayHit *rayhit, RTCIntersectArguments *args);
void rtcOccluded1(RTCScene scene, RTCRay *ray, RTCOccludedArguments *args);
void rtcGetGeometryTransformFromScene(RTCScene scene, unsigned int geomID, float time, RTCFormat format, void *xfm);
void rtcInterpolate0(RTCGeometry g, unsigned int primID, float u, float v, RTCBufferType t, unsigned int slot, float *P, unsigned int n);
void rtcInterpolate1(RTCGeometry g, unsigned int primID, float u, float v, RTCBufferType t, unsigned int slot, float *P, float *dPdu, float *dPdv, unsigned int n);
}
#define rtcInterpolateV0 rtcInterpolate0
