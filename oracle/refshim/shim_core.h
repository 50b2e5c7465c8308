// shim_core.h -- stand-ins for the headers the reference's kernel sources expect from rkcommon 1.14.2
// (rkcommon/math/*.ih in their C++ mode) and from the ISPC standard library, so that the reference's own
// .ih/.ispc files compile as scalar C++ the way the reference's `ospray_module_gpu_kernels` target does
// (modules/cpu/CMakeLists.txt:455-480: uniform= varying= unmasked= export= programCount=1 programIndex=0,
// OSPRAY_TARGET_SYCL, OSPRAY_GLOBAL=static, namespace ispc).  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// rkcommon is NOT vendored in /root/reference; the bodies below are a restatement of its small, well known
// math helpers (vec/linear/affine/box, to_float_unorm, frame, lerp(factor,a,b), rcp_safe, luminance ...).
// They are the same restatement oracle.cpp uses; what this build pins is everything ABOVE them: the
// reference's own samplers, sampling routines, lights, BSDFs, integrators and frame ops, unmodified.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <cassert>
#ifndef SYCL_EXTERNAL
#define SYCL_EXTERNAL
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ispc {

using uint64 = uint64_t;
using uint32 = uint32_t;
using uint16 = uint16_t;
using uint8 = uint8_t;
using int64 = int64_t;
using int32 = int32_t;
using int16 = int16_t;
using int8 = int8_t;
using uint = unsigned int;

// ---- constants (rkcommon/math/math.ih) -------------------------------------------------------------------
static const float inf = INFINITY, neg_inf = -INFINITY, pos_inf = INFINITY, nan = NAN;
static const float flt_min = 1.17549435e-38f, flt_max = 3.40282347e+38f, ulp = 5.9604644775390625e-8f, one_over_255 = 1.f / 255.f;
static const double pi = 3.14159265358979323846, two_pi = 6.283185307179586232, four_pi = 12.566370614359172464,
                    half_pi = 1.57079632679489661923, one_over_pi = 0.31830988618379069122,
                    one_over_two_pi = 0.15915494309189534561, one_over_four_pi = 0.079577471545947672804,
                    one_over_two_pi_sqr = 0.050660591821168885722, lntwo_over_two = 0.346573590279972654709,
                    two_over_pi = 0.636619772367581343076, sqrt_two = 1.41421356237309504880;
struct OneTy {}; struct ZeroTy {}; struct EmptyTy {};
static const OneTy one{}; static const ZeroTy zero{}; static const EmptyTy empty{};

// ---- ISPC stdlib scalars ---------------------------------------------------------------------------------
inline float abs(float x) { return fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float absf(float x) { return fabsf(x); }
inline float sqrt(float x) { return sqrtf(x); }
inline float rcp(float x) { return 1.0f / x; }
inline float rcpf(float x) { return 1.0f / x; }
inline float rsqrt(float x) { return 1.0f / sqrtf(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float rcp_safe(float f) { return 1.f / (fabsf(f) < 1e-8f ? (f < 0.f ? -1e-8f : 1e-8f) : f); }
inline float sqrt_safe(float x) { return sqrtf(x > 0.f ? x : 0.f); }
inline int exclusive_scan_add(int) { return 0; }
inline float expf(float x) { return ::expf(x); }
inline float logf(float x) { return ::logf(x); }
inline float sqr(float x) { return x * x; }
inline float sin(float x) { return sinf(x); }
inline float cos(float x) { return cosf(x); }
inline float tan(float x) { return tanf(x); }
inline float acos(float x) { return acosf(x); }
inline float asin(float x) { return asinf(x); }
inline float atan(float x) { return atanf(x); }
inline float atan2(float y, float x) { return atan2f(y, x); }
inline float exp(float x) { return expf(x); }
inline float log(float x) { return logf(x); }
inline float pow(float a, float b) { return powf(a, b); }
inline float floor(float x) { return floorf(x); }
inline float ceil(float x) { return ceilf(x); }
inline float round(float x) { return roundf(x); }
inline float frac(float x) { return x - floorf(x); }
inline void sincos(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
inline float min(float a, float b) { return a < b ? a : b; }
inline float max(float a, float b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint32 min(uint32 a, uint32 b) { return a < b ? a : b; }
inline uint32 max(uint32 a, uint32 b) { return a > b ? a : b; }
inline int64 min(int64 a, int64 b) { return a < b ? a : b; }
inline int64 max(int64 a, int64 b) { return a > b ? a : b; }
inline uint64 min(uint64 a, uint64 b) { return a < b ? a : b; }
inline uint64 max(uint64 a, uint64 b) { return a > b ? a : b; }
inline float min(float a, float b, float c) { return min(min(a, b), c); }
inline float max(float a, float b, float c) { return max(max(a, b), c); }
inline float min(float a, float b, float c, float d) { return min(min(a, b), min(c, d)); }
inline float max(float a, float b, float c, float d) { return max(max(a, b), max(c, d)); }
inline float clamp(float v, float lo, float hi) { return max(lo, min(v, hi)); }
inline float clamp(float v) { return clamp(v, 0.f, 1.f); }
inline int clamp(int v, int lo, int hi) { return max(lo, min(v, hi)); }
inline uint32 clamp(uint32 v, uint32 lo, uint32 hi) { return max(lo, min(v, hi)); }
inline bool isnan(float x) { return std::isnan(x); }
inline bool isinf(float x) { return std::isinf(x); }
inline float floatbits(uint32 u) { float f; memcpy(&f, &u, 4); return f; }
inline float floatbits(int u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32 intbits(float f) { uint32 u; memcpy(&u, &f, 4); return u; }
inline float select(bool c, float a, float b) { return c ? a : b; }
inline int select(bool c, int a, int b) { return c ? a : b; }
inline bool any(bool b) { return b; }
inline bool all(bool b) { return b; }
inline bool none(bool b) { return !b; }
inline float reduce_max(float x) { return x; }
inline float reduce_min(float x) { return x; }
inline float reduce_add(float x) { return x; }
inline int reduce_max(int x) { return x; }
inline int reduce_min(int x) { return x; }
inline uint32 reduce_max(uint32 x) { return x; }
inline float extract(float x, int) { return x; }
inline int extract(int x, int) { return x; }
inline uint32 extract(uint32 x, int) { return x; }
inline int popcnt(uint32 x) { return __builtin_popcount(x); }
inline int count_leading_zeros(uint32 x) { return x ? __builtin_clz(x) : 32; }
inline int count_trailing_zeros(uint32 x) { return x ? __builtin_ctz(x) : 32; }
inline uint32 lanemask() { return 1u; }
inline float lerp(float factor, float a, float b) { return (1.f - factor) * a + factor * b; }
inline float deg2rad(float x) { return x * 1.74532925199432957692e-2f; }
inline float rad2deg(float x) { return x * 5.72957795130823208768e1f; }
inline float cos2sin(float f) { return sqrt(max(0.f, 1.f - sqr(f))); }
inline float sin2cos(float f) { return cos2sin(f); }
inline float roundToPow2(float x) { return x; }
inline bool similar(float a, float b) { return fabsf(a - b) <= 1e-5f * max(fabsf(a), fabsf(b)); }
inline float divide_safe(float f) { return 1.f / ((fabsf(f) < flt_min) ? (f >= 0.f ? flt_min : -flt_min) : f); }
inline uint32 signbits(float x) { return intbits(x) & 0x80000000u; }
inline float nextafter(float a, float b) { return ::nextafterf(a, b); }
// rkcommon/math/math.ih: u32 -> [0,1]: one int->float conversion then an exact power-of-two scale
inline float to_float_unorm(uint32 a) { return (float)a * 0x1.0p-32f; }

// ---- short vectors ---------------------------------------------------------------------------------------
template <typename T> struct V2 { T x, y; V2() = default; V2(T x_, T y_) : x(x_), y(y_) {} V2(T s) : x(s), y(s) {} };
template <typename T> struct V3 { T x, y, z; V3() = default; V3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {} V3(T s) : x(s), y(s), z(s) {} };
template <typename T> struct V4 { T x, y, z, w; V4() = default; V4(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {} V4(T s) : x(s), y(s), z(s), w(s) {} };
using vec2f = V2<float>; using vec3f = V3<float>; using vec4f = V4<float>;
using vec2i = V2<int32>; using vec3i = V3<int32>; using vec4i = V4<int32>;
using vec2ui = V2<uint32>; using vec3ui = V3<uint32>; using vec4ui = V4<uint32>;
using vec2uc = V2<uint8>; using vec3uc = V3<uint8>; using vec4uc = V4<uint8>;
using vec2l = V2<int64>; using vec3l = V3<int64>; using vec3ul = V3<uint64>; using vec2ul = V2<uint64>;
using vec3fa = vec3f;
struct vec3b { bool x, y, z; };
struct vec2b { bool x, y; };

#define SHIM_MAKE(T, S)                                                                                      \
  inline T##2##S make_vec2##S(T x, T y) { return {x, y}; }                                                   \
  inline T##2##S make_vec2##S(T x) { return {x, x}; }                                                        \
  inline T##3##S make_vec3##S(T x, T y, T z) { return {x, y, z}; }                                           \
  inline T##3##S make_vec3##S(T x) { return {x, x, x}; }                                                     \
  inline T##4##S make_vec4##S(T x, T y, T z, T w) { return {x, y, z, w}; }                                   \
  inline T##4##S make_vec4##S(T x) { return {x, x, x, x}; }
#undef SHIM_MAKE
#define SHIM_MAKE(S, T)                                                                                      \
  inline V2<T> make_vec2##S(T x, T y) { return {x, y}; }                                                     \
  inline V2<T> make_vec2##S(T x) { return {x, x}; }                                                          \
  inline V3<T> make_vec3##S(T x, T y, T z) { return {x, y, z}; }                                             \
  inline V3<T> make_vec3##S(T x) { return {x, x, x}; }                                                       \
  inline V4<T> make_vec4##S(T x, T y, T z, T w) { return {x, y, z, w}; }                                     \
  inline V4<T> make_vec4##S(T x) { return {x, x, x, x}; }                                                    \
  template <typename U> inline V2<T> make_vec2##S(const V2<U> &v) { return {(T)v.x, (T)v.y}; }               \
  template <typename U> inline V2<T> make_vec2##S(const V3<U> &v) { return {(T)v.x, (T)v.y}; }               \
  template <typename U> inline V3<T> make_vec3##S(const V3<U> &v) { return {(T)v.x, (T)v.y, (T)v.z}; }       \
  template <typename U> inline V3<T> make_vec3##S(const V4<U> &v) { return {(T)v.x, (T)v.y, (T)v.z}; }       \
  template <typename U> inline V3<T> make_vec3##S(const V2<U> &v, T z) { return {(T)v.x, (T)v.y, z}; }       \
  template <typename U> inline V4<T> make_vec4##S(const V4<U> &v) { return {(T)v.x, (T)v.y, (T)v.z, (T)v.w}; } \
  template <typename U> inline V4<T> make_vec4##S(const V3<U> &v, T w) { return {(T)v.x, (T)v.y, (T)v.z, w}; } \
  template <typename U> inline V4<T> make_vec4##S(const V3<U> &v) { return {(T)v.x, (T)v.y, (T)v.z, (T)0}; }
SHIM_MAKE(f, float)
SHIM_MAKE(i, int32)
SHIM_MAKE(ui, uint32)
SHIM_MAKE(uc, uint8)
SHIM_MAKE(l, int64)
SHIM_MAKE(ul, uint64)
#undef SHIM_MAKE
inline vec3f make_vec3fa(float x, float y, float z) { return {x, y, z}; }
inline vec3f make_vec3fa(float x) { return {x, x, x}; }

#define SHIM_BINOP(op)                                                                                       \
  template <typename T> inline V2<T> operator op(const V2<T> &a, const V2<T> &b) { return {(T)(a.x op b.x), (T)(a.y op b.y)}; } \
  template <typename T> inline V3<T> operator op(const V3<T> &a, const V3<T> &b) { return {(T)(a.x op b.x), (T)(a.y op b.y), (T)(a.z op b.z)}; } \
  template <typename T> inline V4<T> operator op(const V4<T> &a, const V4<T> &b) { return {(T)(a.x op b.x), (T)(a.y op b.y), (T)(a.z op b.z), (T)(a.w op b.w)}; } \
  template <typename T> inline V2<T> operator op(const V2<T> &a, T b) { return {(T)(a.x op b), (T)(a.y op b)}; } \
  template <typename T> inline V3<T> operator op(const V3<T> &a, T b) { return {(T)(a.x op b), (T)(a.y op b), (T)(a.z op b)}; } \
  template <typename T> inline V4<T> operator op(const V4<T> &a, T b) { return {(T)(a.x op b), (T)(a.y op b), (T)(a.z op b), (T)(a.w op b)}; } \
  template <typename T> inline V2<T> operator op(T a, const V2<T> &b) { return {(T)(a op b.x), (T)(a op b.y)}; } \
  template <typename T> inline V3<T> operator op(T a, const V3<T> &b) { return {(T)(a op b.x), (T)(a op b.y), (T)(a op b.z)}; } \
  template <typename T> inline V4<T> operator op(T a, const V4<T> &b) { return {(T)(a op b.x), (T)(a op b.y), (T)(a op b.z), (T)(a op b.w)}; }
SHIM_BINOP(+)
SHIM_BINOP(-)
SHIM_BINOP(*)
SHIM_BINOP(/)
#undef SHIM_BINOP
// mixed scalar types the ISPC sources rely on (int literal * vec3f etc.)
inline vec3f operator*(double a, const vec3f &b) { return (float)a * b; }
inline vec3f operator*(const vec3f &a, double b) { return a * (float)b; }
inline vec3f operator*(int a, const vec3f &b) { return (float)a * b; }
inline vec3f operator*(const vec3f &a, int b) { return a * (float)b; }
inline vec2f operator*(double a, const vec2f &b) { return (float)a * b; }
inline vec2f operator*(const vec2f &a, double b) { return a * (float)b; }
inline vec2f operator*(const vec2f &a, int b) { return a * (float)b; }
inline vec2f operator*(int a, const vec2f &b) { return (float)a * b; }
inline vec4f operator*(double a, const vec4f &b) { return (float)a * b; }
inline vec4f operator*(const vec4f &a, double b) { return a * (float)b; }
inline vec3f operator/(const vec3f &a, double b) { return a / (float)b; }
inline vec3f operator+(const vec3f &a, double b) { return a + (float)b; }
inline vec3f operator-(const vec3f &a, double b) { return a - (float)b; }
inline vec3f operator-(double a, const vec3f &b) { return (float)a - b; }
inline vec2f operator-(const vec2f &a, double b) { return a - (float)b; }
inline vec2f operator+(const vec2f &a, double b) { return a + (float)b; }
inline vec2f operator/(const vec2f &a, double b) { return a / (float)b; }
inline vec2ui operator-(const vec2ui &a, int b) { return {a.x - (uint32)b, a.y - (uint32)b}; }
inline vec2ui operator+(const vec2ui &a, int b) { return {a.x + (uint32)b, a.y + (uint32)b}; }
inline vec2i operator*(const vec2i &a, uint32 b) { return {(int)(a.x * b), (int)(a.y * b)}; }

template <typename T> inline V2<T> operator-(const V2<T> &a) { return {(T)-a.x, (T)-a.y}; }
template <typename T> inline V3<T> operator-(const V3<T> &a) { return {(T)-a.x, (T)-a.y, (T)-a.z}; }
template <typename T> inline V4<T> operator-(const V4<T> &a) { return {(T)-a.x, (T)-a.y, (T)-a.z, (T)-a.w}; }
template <typename T> inline V2<T> neg(const V2<T> &a) { return -a; }
template <typename T> inline V3<T> neg(const V3<T> &a) { return -a; }
template <typename T> inline V4<T> neg(const V4<T> &a) { return -a; }
template <typename T> inline V3<T> operator>>(const V3<T> &a, int s) { return {(T)(a.x >> s), (T)(a.y >> s), (T)(a.z >> s)}; }
template <typename T> inline V3<T> operator<<(const V3<T> &a, int s) { return {(T)(a.x << s), (T)(a.y << s), (T)(a.z << s)}; }
template <typename T> inline V2<T> operator>>(const V2<T> &a, int s) { return {(T)(a.x >> s), (T)(a.y >> s)}; }
template <typename T> inline V2<T> operator<<(const V2<T> &a, int s) { return {(T)(a.x << s), (T)(a.y << s)}; }
template <typename T> inline V3<T> operator&(const V3<T> &a, T m) { return {(T)(a.x & m), (T)(a.y & m), (T)(a.z & m)}; }

template <typename T> inline bool eq(const V2<T> &a, const V2<T> &b) { return a.x == b.x && a.y == b.y; }
template <typename T> inline bool eq(const V3<T> &a, const V3<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> inline bool eq(const V4<T> &a, const V4<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
template <typename T> inline bool eq(const V2<T> &a, T b) { return a.x == b && a.y == b; }
template <typename T> inline bool eq(const V3<T> &a, T b) { return a.x == b && a.y == b && a.z == b; }
template <typename T> inline bool ne(const V2<T> &a, const V2<T> &b) { return !eq(a, b); }
template <typename T> inline bool ne(const V3<T> &a, const V3<T> &b) { return !eq(a, b); }
template <typename T> inline bool ne(const V3<T> &a, T b) { return !eq(a, b); }
template <typename T> inline bool ne(const V4<T> &a, const V4<T> &b) { return !eq(a, b); }
template <typename T> inline bool operator==(const V2<T> &a, const V2<T> &b) { return eq(a, b); }
template <typename T> inline bool operator==(const V3<T> &a, const V3<T> &b) { return eq(a, b); }
template <typename T> inline bool operator!=(const V2<T> &a, const V2<T> &b) { return !eq(a, b); }
template <typename T> inline bool operator!=(const V3<T> &a, const V3<T> &b) { return !eq(a, b); }
template <typename T> inline bool anyLessThan(const V2<T> &a, const V2<T> &b) { return a.x < b.x || a.y < b.y; }
template <typename T> inline bool anyLessThan(const V3<T> &a, const V3<T> &b) { return a.x < b.x || a.y < b.y || a.z < b.z; }
inline vec3b lt(const vec3f &a, const vec3f &b) { return {a.x < b.x, a.y < b.y, a.z < b.z}; }
inline vec3b le(const vec3f &a, const vec3f &b) { return {a.x <= b.x, a.y <= b.y, a.z <= b.z}; }
inline vec3b gt(const vec3f &a, const vec3f &b) { return {a.x > b.x, a.y > b.y, a.z > b.z}; }
inline vec3b ge(const vec3f &a, const vec3f &b) { return {a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline vec3b lt(const vec3f &a, float b) { return {a.x < b, a.y < b, a.z < b}; }
inline vec3b gt(const vec3f &a, float b) { return {a.x > b, a.y > b, a.z > b}; }
inline bool any(const vec3b &b) { return b.x || b.y || b.z; }
inline bool all(const vec3b &b) { return b.x && b.y && b.z; }
inline vec3ui intbits(const vec3f &v) { return {intbits(v.x), intbits(v.y), intbits(v.z)}; }
inline bool isnan(const vec3f &v) { return isnan(v.x + v.y + v.z); }

#define SHIM_UNARY(fn)                                                                                       \
  inline vec2f fn(const vec2f &a) { return {fn(a.x), fn(a.y)}; }                                             \
  inline vec3f fn(const vec3f &a) { return {fn(a.x), fn(a.y), fn(a.z)}; }                                    \
  inline vec4f fn(const vec4f &a) { return {fn(a.x), fn(a.y), fn(a.z), fn(a.w)}; }
SHIM_UNARY(expf) SHIM_UNARY(logf) SHIM_UNARY(sqrt_safe) SHIM_UNARY(abs) SHIM_UNARY(absf) SHIM_UNARY(sqrt) SHIM_UNARY(rcp) SHIM_UNARY(rcp_safe) SHIM_UNARY(exp) SHIM_UNARY(log) SHIM_UNARY(floor)
SHIM_UNARY(ceil) SHIM_UNARY(frac) SHIM_UNARY(sqr) SHIM_UNARY(rsqrt) SHIM_UNARY(sin) SHIM_UNARY(cos) SHIM_UNARY(divide_safe)
#undef SHIM_UNARY
inline vec3f rcpf(const vec3f &a) { return rcp(a); }
#define SHIM_BINARY(fn)                                                                                      \
  template <typename T> inline V2<T> fn(const V2<T> &a, const V2<T> &b) { return {fn(a.x, b.x), fn(a.y, b.y)}; } \
  template <typename T> inline V3<T> fn(const V3<T> &a, const V3<T> &b) { return {fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z)}; } \
  template <typename T> inline V4<T> fn(const V4<T> &a, const V4<T> &b) { return {fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z), fn(a.w, b.w)}; } \
  template <typename T> inline V2<T> fn(const V2<T> &a, T b) { return {fn(a.x, b), fn(a.y, b)}; }           \
  template <typename T> inline V3<T> fn(const V3<T> &a, T b) { return {fn(a.x, b), fn(a.y, b), fn(a.z, b)}; } \
  template <typename T> inline V4<T> fn(const V4<T> &a, T b) { return {fn(a.x, b), fn(a.y, b), fn(a.z, b), fn(a.w, b)}; } \
  template <typename T> inline V3<T> fn(T a, const V3<T> &b) { return {fn(a, b.x), fn(a, b.y), fn(a, b.z)}; }
SHIM_BINARY(min) SHIM_BINARY(max)
#undef SHIM_BINARY
inline vec3f pow(const vec3f &a, float b) { return {pow(a.x, b), pow(a.y, b), pow(a.z, b)}; }
inline vec3f pow(const vec3f &a, const vec3f &b) { return {pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)}; }
inline vec4f pow(const vec4f &a, float b) { return {pow(a.x, b), pow(a.y, b), pow(a.z, b), pow(a.w, b)}; }
template <typename T> inline V2<T> clamp(const V2<T> &v, const V2<T> &lo, const V2<T> &hi) { return max(lo, min(v, hi)); }
template <typename T> inline V3<T> clamp(const V3<T> &v, const V3<T> &lo, const V3<T> &hi) { return max(lo, min(v, hi)); }
template <typename T> inline V4<T> clamp(const V4<T> &v, const V4<T> &lo, const V4<T> &hi) { return max(lo, min(v, hi)); }
inline vec3f clamp(const vec3f &v) { return {clamp(v.x), clamp(v.y), clamp(v.z)}; }
inline vec3f clamp(const vec3f &v, float lo, float hi) { return {clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)}; }

inline float dot(const vec2f &a, const vec2f &b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3f &a, const vec3f &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4f &a, const vec4f &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float absdot(const vec3f &a, const vec3f &b) { return abs(dot(a, b)); }
inline vec3f cross(const vec3f &a, const vec3f &b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(const vec2f &a) { return sqrt(dot(a, a)); }
inline float length(const vec3f &a) { return sqrt(dot(a, a)); }
inline float distance(const vec3f &a, const vec3f &b) { return length(a - b); }
inline vec3f normalize(const vec3f &v) { return v * (1.f / sqrt(dot(v, v))); }
inline vec3f normalize(const vec3f &v, float &len) { len = sqrt(dot(v, v)); return v * rcpf(len); }
inline vec3f safe_normalize(const vec3f &v) { return v * (1.f / sqrt(max(flt_min, dot(v, v)))); }
inline vec2f normalize(const vec2f &v) { return v * (1.f / sqrt(dot(v, v))); }
inline float reduce_max(const vec3f &v) { return max(max(v.x, v.y), v.z); }
inline float reduce_min(const vec3f &v) { return min(min(v.x, v.y), v.z); }
inline float reduce_add(const vec3f &v) { return v.x + v.y + v.z; }
inline float reduce_mul(const vec3f &v) { return v.x * v.y * v.z; }
inline float reduce_avg(const vec3f &v) { return (v.x + v.y + v.z) * (1.0f / 3.0f); }
inline float reduce_max(const vec4f &v) { return max(max(v.x, v.y), max(v.z, v.w)); }
inline float reduce_min(const vec4f &v) { return min(min(v.x, v.y), min(v.z, v.w)); }
inline float reduce_add(const vec4f &v) { return v.x + v.y + v.z + v.w; }
inline float reduce_max(const vec2f &v) { return max(v.x, v.y); }
inline float reduce_min(const vec2f &v) { return min(v.x, v.y); }
inline int reduce_mul(const vec3i &v) { return v.x * v.y * v.z; }
inline int reduce_mul(const vec2i &v) { return v.x * v.y; }
inline int reduce_max(const vec3i &v) { return max(max(v.x, v.y), v.z); }
inline int reduce_min(const vec3i &v) { return min(min(v.x, v.y), v.z); }
inline float luminance(const vec3f &c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }
inline vec2f lerp(float f, const vec2f &a, const vec2f &b) { return (1.f - f) * a + f * b; }
inline vec3f lerp(float f, const vec3f &a, const vec3f &b) { return (1.f - f) * a + f * b; }
inline vec4f lerp(float f, const vec4f &a, const vec4f &b) { return (1.f - f) * a + f * b; }
inline vec2f lerp(const vec2f &f, const vec2f &a, const vec2f &b) { return (make_vec2f(1.f) - f) * a + f * b; }
inline vec3f lerp(const vec3f &f, const vec3f &a, const vec3f &b) { return (make_vec3f(1.f) - f) * a + f * b; }
inline vec3f interpolate(const vec3f &f, const vec3f &a, const vec3f &b, const vec3f &c) { return f.x * a + f.y * b + f.z * c; }
inline vec4f interpolate(const vec3f &f, const vec4f &a, const vec4f &b, const vec4f &c) { return f.x * a + f.y * b + f.z * c; }
inline vec2f interpolate(const vec3f &f, const vec2f &a, const vec2f &b, const vec2f &c) { return f.x * a + f.y * b + f.z * c; }
inline vec2f to_float_unorm(const vec2ui &a) { return {to_float_unorm(a.x), to_float_unorm(a.y)}; }
inline vec3f to_float(const vec3i &a) { return {(float)a.x, (float)a.y, (float)a.z}; }
inline vec2f to_float(const vec2i &a) { return {(float)a.x, (float)a.y}; }
inline vec3i to_int(const vec3f &a) { return {(int)a.x, (int)a.y, (int)a.z}; }
inline vec2i to_int(const vec2f &a) { return {(int)a.x, (int)a.y}; }
inline vec3f select(bool c, const vec3f &a, const vec3f &b) { return c ? a : b; }
inline vec3f madd(const vec3f &a, const vec3f &b, const vec3f &c) { return a * b + c; }
inline vec3f madd(float a, const vec3f &b, const vec3f &c) { return a * b + c; }
// orthonormal frame (rkcommon/math/LinearSpace.ih frame()): restated, see oracle.cpp frame()
inline vec3f rotate(const vec3f &v, const vec3f &axis, float theta) // Rodrigues (rkcommon/math/vec.ih, restated)
{
  return v * cos(theta) + cross(axis, v) * sin(theta) + axis * dot(axis, v) * (1.f - cos(theta));
}
inline int maxDim(const vec3f &v) { return v.x > v.y ? (v.x > v.z ? 0 : 2) : (v.y > v.z ? 1 : 2); }
inline float component(const vec3f &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// ---- linear / affine space, boxes ------------------------------------------------------------------------
struct linear3f { vec3f vx, vy, vz; linear3f() = default; linear3f(const vec3f &x, const vec3f &y, const vec3f &z) : vx(x), vy(y), vz(z) {} linear3f(OneTy) : vx(1, 0, 0), vy(0, 1, 0), vz(0, 0, 1) {} };
using LinearSpace3f = linear3f;
struct LinearSpace2f { vec2f vx, vy; LinearSpace2f() = default; LinearSpace2f(const vec2f &x, const vec2f &y) : vx(x), vy(y) {} LinearSpace2f(OneTy) : vx(1, 0), vy(0, 1) {} };
using linear2f = LinearSpace2f;
struct affine3f { linear3f l; vec3f p; affine3f() = default; affine3f(const linear3f &l_, const vec3f &p_) : l(l_), p(p_) {} affine3f(OneTy) : l(one), p(0, 0, 0) {} };
using AffineSpace3f = affine3f;
struct AffineSpace2f { LinearSpace2f l; vec2f p; AffineSpace2f() = default; AffineSpace2f(const LinearSpace2f &l_, const vec2f &p_) : l(l_), p(p_) {} AffineSpace2f(OneTy) : l(one), p(0, 0) {} };
using affine2f = AffineSpace2f;
inline linear3f make_LinearSpace3f(const vec3f &x, const vec3f &y, const vec3f &z) { return {x, y, z}; }
inline linear3f make_LinearSpace3f_identity() { return {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; }
inline linear3f make_LinearSpace3f(const affine3f &a) { return a.l; }
inline affine3f make_AffineSpace3f(const linear3f &l, const vec3f &p) { return {l, p}; }
inline affine3f make_AffineSpace3f(const linear3f &l) { return {l, {0, 0, 0}}; }
inline affine3f make_AffineSpace3f_identity() { return {make_LinearSpace3f_identity(), {0, 0, 0}}; }
inline vec3f operator*(const linear3f &l, const vec3f &v) { return v.x * l.vx + v.y * l.vy + v.z * l.vz; }
inline vec3f xfmVector(const linear3f &l, const vec3f &v) { return v.x * l.vx + v.y * l.vy + v.z * l.vz; }
inline vec3f xfmPoint(const affine3f &a, const vec3f &v) { return a.p + v.x * a.l.vx + v.y * a.l.vy + v.z * a.l.vz; }
inline vec3f xfmVector(const affine3f &a, const vec3f &v) { return v.x * a.l.vx + v.y * a.l.vy + v.z * a.l.vz; }
inline vec3f operator*(const affine3f &a, const vec3f &v) { return xfmPoint(a, v); }
inline linear3f operator*(const linear3f &a, const linear3f &b) { return {a * b.vx, a * b.vy, a * b.vz}; }
inline linear3f operator*(float s, const linear3f &a) { return {s * a.vx, s * a.vy, s * a.vz}; }
inline affine3f operator*(const affine3f &a, const affine3f &b) { return {a.l * b.l, a.l * b.p + a.p}; }
inline linear3f transposed(const linear3f &l) { return {{l.vx.x, l.vy.x, l.vz.x}, {l.vx.y, l.vy.y, l.vz.y}, {l.vx.z, l.vy.z, l.vz.z}}; }
inline float det(const linear3f &l) { return dot(l.vx, cross(l.vy, l.vz)); }
inline linear3f adjoint(const linear3f &l) { return transposed(linear3f{cross(l.vy, l.vz), cross(l.vz, l.vx), cross(l.vx, l.vy)}); }
inline linear3f rcp(const linear3f &l) { return rcp(det(l)) * adjoint(l); }
inline affine3f rcp(const affine3f &a) { const linear3f il = rcp(a.l); return {il, -(il * a.p)}; }
inline linear3f frame(const vec3f &N)
{
  const vec3f dx0 = make_vec3f(0.0f, N.z, -N.y);
  const vec3f dx1 = make_vec3f(-N.z, 0.0f, N.x);
  const vec3f dx = normalize(abs(N.x) < abs(N.y) ? dx0 : dx1);
  const vec3f dy = cross(N, dx);
  return make_LinearSpace3f(dx, dy, N);
}
inline vec2f xfmPoint(const AffineSpace2f &a, const vec2f &v) { return a.p + v.x * a.l.vx + v.y * a.l.vy; }
inline vec2f xfmVector(const AffineSpace2f &a, const vec2f &v) { return v.x * a.l.vx + v.y * a.l.vy; }
inline vec2f operator*(const LinearSpace2f &l, const vec2f &v) { return v.x * l.vx + v.y * l.vy; }
inline vec2f operator*(const AffineSpace2f &a, const vec2f &v) { return a.p + v.x * a.l.vx + v.y * a.l.vy; }

struct box1f { float lower, upper; box1f() = default; box1f(float l, float u) : lower(l), upper(u) {} explicit box1f(float v) : lower(v), upper(v) {} box1f(EmptyTy) : lower(INFINITY), upper(-INFINITY) {} };
using range1f = box1f;
struct box2f { vec2f lower, upper; box2f() = default; box2f(const vec2f &l, const vec2f &u) : lower(l), upper(u) {} explicit box2f(float v) : lower(v), upper(v) {} box2f(EmptyTy) : lower(INFINITY), upper(-INFINITY) {} };
struct box2i { vec2i lower, upper; box2i() = default; box2i(const vec2i &l, const vec2i &u) : lower(l), upper(u) {} box2i(EmptyTy) : lower(INT32_MAX), upper(INT32_MIN) {} };
using range2i = box2i;
struct box3f { vec3f lower, upper; box3f() = default; box3f(const vec3f &l, const vec3f &u) : lower(l), upper(u) {} explicit box3f(float v) : lower(v), upper(v) {} box3f(EmptyTy) : lower(INFINITY), upper(-INFINITY) {} };
using box3fa = box3f;
struct box3i { vec3i lower, upper; box3i() = default; box3i(const vec3i &l, const vec3i &u) : lower(l), upper(u) {} box3i(EmptyTy) : lower(INT32_MAX), upper(INT32_MIN) {} };
struct box4f { vec4f lower, upper; };
inline box1f make_box1f(float lo, float hi) { return {lo, hi}; }
inline box1f make_box1f(const vec2f &v) { return {v.x, v.y}; }
inline box2f make_box2f(const vec2f &lo, const vec2f &hi) { return {lo, hi}; }
inline box2i make_box2i(const vec2i &lo, const vec2i &hi) { return {lo, hi}; }
inline box3f make_box3f(const vec3f &lo, const vec3f &hi) { return {lo, hi}; }
inline box3f make_box3fa(const vec3f &lo, const vec3f &hi) { return {lo, hi}; }
inline box3i make_box3i(const vec3i &lo, const vec3i &hi) { return {lo, hi}; }
inline box3f make_box3fa(const box3f &b) { return b; }
inline box3f make_box3f_empty() { return {make_vec3f(inf), make_vec3f(neg_inf)}; }
inline box3f make_box3fa_empty() { return make_box3f_empty(); }
inline box1f make_box1f_empty() { return {inf, neg_inf}; }
inline bool isEmpty(const box1f &b) { return b.upper <= b.lower; }
inline bool isEmpty(const box2i &b) { return b.upper.x <= b.lower.x || b.upper.y <= b.lower.y; }
inline bool isEmpty(const box3f &b) { return anyLessThan(b.upper, b.lower); }
inline vec2f box_size(const box2f &b) { return b.upper - b.lower; }
inline vec3f box_size(const box3f &b) { return b.upper - b.lower; }
inline vec2i box_size(const box2i &b) { return b.upper - b.lower; }
inline float box_size(const box1f &b) { return b.upper - b.lower; }
inline box3f box_extend(const box3f &b, const vec3f &v) { return {min(b.lower, v), max(b.upper, v)}; }
inline box3f box_extend(const box3f &b, const box3f &o) { return {min(b.lower, o.lower), max(b.upper, o.upper)}; }
inline void extend(box3f &b, const vec3f &v) { b.lower = min(b.lower, v); b.upper = max(b.upper, v); }
inline void extend(box3f &b, const box3f &o) { b.lower = min(b.lower, o.lower); b.upper = max(b.upper, o.upper); }
inline void extend(box1f &b, float v) { b.lower = min(b.lower, v); b.upper = max(b.upper, v); }
inline bool box_contains(const box3f &b, const vec3f &p) { return !anyLessThan(p, b.lower) && !anyLessThan(b.upper, p); }
inline bool contains(const box3f &b, const vec3f &p) { return box_contains(b, p); }
inline box3f xfmBox(const affine3f &m, const box3f &b)
{
  box3f r = make_box3f_empty();
  for (int i = 0; i < 8; i++)
    extend(r, xfmPoint(m, make_vec3f(i & 1 ? b.upper.x : b.lower.x, i & 2 ? b.upper.y : b.lower.y, i & 4 ? b.upper.z : b.lower.z)));
  return r;
}
inline box3f intersectionOf(const box3f &a, const box3f &b) { return {max(a.lower, b.lower), min(a.upper, b.upper)}; }
inline box2f intersect(const box2f &a, const box2f &b) { return {max(a.lower, b.lower), min(a.upper, b.upper)}; }

// rkcommon colour helpers used by the frame ops (restated exactly as in oracle.cpp: IEC 61966-2-1 transfer, ISPC round =
// round-half-even)
inline float linear_to_srgb(float f) { const float c = max(f, 0.f); return c <= 0.0031308f ? 12.92f * c : powf(c, 1.f / 2.4f) * 1.055f - 0.055f; }
inline vec4f linear_to_srgba(const vec4f &c) { return {linear_to_srgb(c.x), linear_to_srgb(c.y), linear_to_srgb(c.z), max(c.w, 0.f)}; }
inline uint32 cvt_uint32(float f) { return (uint32)rintf(255.f * min(max(f, 0.f), 1.f)); }
inline uint32 cvt_uint32(const vec4f &v) { return (cvt_uint32(v.x) << 0) | (cvt_uint32(v.y) << 8) | (cvt_uint32(v.z) << 16) | (cvt_uint32(v.w) << 24); }
inline uint32 cvt_uint32(const vec3f &v) { return (cvt_uint32(v.x) << 0) | (cvt_uint32(v.y) << 8) | (cvt_uint32(v.z) << 16); }
inline uint32 linear_to_srgba8(const vec4f &c) { return cvt_uint32(linear_to_srgba(c)); }
inline float srgb_to_linear_(float x) { return x <= 0.04045f ? x * (1.f / 12.92f) : powf((x + 0.055f) * (1.f / 1.055f), 2.4f); }
inline vec4f srgba_to_linear(const vec4f &c) { return {srgb_to_linear_(c.x), srgb_to_linear_(c.y), srgb_to_linear_(c.z), c.w}; }
inline vec3f srgb_to_linear(const vec3f &c) { return {srgb_to_linear_(c.x), srgb_to_linear_(c.y), srgb_to_linear_(c.z)}; }
// rkcommon/utility/random.ih makeRandomColor (debug renderer only; restated)
inline vec3f makeRandomColor(uint32 i)
{
  const uint32 mx = 13 * 17 * 43, my = 11 * 29, mz = 7 * 23 * 63, g = (i * (3 * 5 * 127) + 12312314);
  if (i == (uint32)-1) return make_vec3f(0.f);
  return make_vec3f((g % mx) * (1.f / (mx - 1)), (g % my) * (1.f / (my - 1)), (g % mz) * (1.f / (mz - 1)));
}
inline float srgb_to_linear(float x) { return x <= 0.04045f ? x * (1.f / 12.92f) : pow((x + 0.055f) * (1.f / 1.055f), 2.4f); }

} // namespace ispc
