// ospray-cuda: tiny vector/affine library shared by the host module (g++) and the
// CUDA kernels (nvcc). Replaces the subset of rkcommon 1.14.2 math the reference
// leans on (vec*, box*, linear3f, affine3f; reference uses: rkcommon/math/*.h,
// un-vendored). Everything is POD so the same structs cross the C-ABI in ospcuda.h.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define OSP_HD __host__ __device__ __forceinline__
#else
#define OSP_HD inline
#endif

namespace ospm {

struct vec2f { float x, y; };
struct vec3f { float x, y, z; };
struct vec4f { float x, y, z, w; };
struct vec2i { int x, y; };
struct vec3i { int x, y, z; };
struct vec2ui { uint32_t x, y; };
struct vec3ui { uint32_t x, y, z; };
struct linear2f { vec2f vx, vy; };
struct vec4ui { uint32_t x, y, z, w; };
struct vec3ul { uint64_t x, y, z; };
struct vec3l { int64_t x, y, z; };
struct box1f { float lower, upper; };
struct box3f { vec3f lower, upper; };
struct linear3f { vec3f vx, vy, vz; };           // columns
struct affine3f { linear3f l; vec3f p; };       // 3x4 column-major, as OSP_AFFINE3F

OSP_HD vec2f make_vec2f(float x, float y) { vec2f v; v.x = x; v.y = y; return v; }
OSP_HD vec3f make_vec3f(float x, float y, float z) { vec3f v; v.x = x; v.y = y; v.z = z; return v; }
OSP_HD vec3f make_vec3f(float s) { return make_vec3f(s, s, s); }
OSP_HD vec4f make_vec4f(float x, float y, float z, float w) { vec4f v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
OSP_HD vec4f make_vec4f(vec3f a, float w) { return make_vec4f(a.x, a.y, a.z, w); }
OSP_HD vec3f make_vec3f(vec4f a) { return make_vec3f(a.x, a.y, a.z); }

OSP_HD vec3f operator+(vec3f a, vec3f b) { return make_vec3f(a.x + b.x, a.y + b.y, a.z + b.z); }
OSP_HD vec3f operator-(vec3f a, vec3f b) { return make_vec3f(a.x - b.x, a.y - b.y, a.z - b.z); }
OSP_HD vec3f operator*(vec3f a, vec3f b) { return make_vec3f(a.x * b.x, a.y * b.y, a.z * b.z); }
OSP_HD vec3f operator/(vec3f a, vec3f b) { return make_vec3f(a.x / b.x, a.y / b.y, a.z / b.z); }
OSP_HD vec3f operator*(float s, vec3f a) { return make_vec3f(s * a.x, s * a.y, s * a.z); }
OSP_HD vec3f operator*(vec3f a, float s) { return make_vec3f(a.x * s, a.y * s, a.z * s); }
OSP_HD vec3f operator/(vec3f a, float s) { return make_vec3f(a.x / s, a.y / s, a.z / s); }
OSP_HD vec3f operator-(vec3f a) { return make_vec3f(-a.x, -a.y, -a.z); }
OSP_HD vec3f neg(vec3f a) { return -a; }
OSP_HD vec2f operator+(vec2f a, vec2f b) { return make_vec2f(a.x + b.x, a.y + b.y); }
OSP_HD vec2f operator-(vec2f a, vec2f b) { return make_vec2f(a.x - b.x, a.y - b.y); }
OSP_HD vec2f operator*(vec2f a, vec2f b) { return make_vec2f(a.x * b.x, a.y * b.y); }
OSP_HD vec2f operator*(vec2f a, float s) { return make_vec2f(a.x * s, a.y * s); }
OSP_HD vec2f operator*(float s, vec2f a) { return make_vec2f(a.x * s, a.y * s); }
OSP_HD vec4f operator+(vec4f a, vec4f b) { return make_vec4f(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
OSP_HD vec4f operator-(vec4f a, vec4f b) { return make_vec4f(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
OSP_HD vec4f operator*(float s, vec4f a) { return make_vec4f(s * a.x, s * a.y, s * a.z, s * a.w); }
OSP_HD vec4f operator*(vec4f a, float s) { return make_vec4f(s * a.x, s * a.y, s * a.z, s * a.w); }

OSP_HD float dot(vec3f a, vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OSP_HD vec3f cross(vec3f a, vec3f b)
{
  return make_vec3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
OSP_HD float length(vec3f a) { return sqrtf(dot(a, a)); }
OSP_HD vec3f normalize(vec3f a) { return a * (1.f / sqrtf(dot(a, a))); }
OSP_HD float reduce_max(vec3f a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
OSP_HD float reduce_min(vec3f a) { return fminf(a.x, fminf(a.y, a.z)); }
OSP_HD float reduce_add(vec3f a) { return a.x + a.y + a.z; }
OSP_HD vec3f vabs(vec3f a) { return make_vec3f(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
OSP_HD vec3f vmin(vec3f a, vec3f b) { return make_vec3f(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
OSP_HD vec3f vmax(vec3f a, vec3f b) { return make_vec3f(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
OSP_HD float sqr(float x) { return x * x; }
OSP_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
// rkcommon convention: lerp(factor, a, b) = (1-f)*a + f*b
OSP_HD float lerpf(float f, float a, float b) { return (1.f - f) * a + f * b; }
OSP_HD vec3f lerp3(float f, vec3f a, vec3f b) { return (1.f - f) * a + f * b; }
OSP_HD vec4f lerp4(float f, vec4f a, vec4f b) { return (1.f - f) * a + f * b; }
OSP_HD float luminance(vec3f c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }
OSP_HD bool eq3(vec3f a, vec3f b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// linear / affine ---------------------------------------------------------------
OSP_HD vec3f xfmVector(const linear3f &l, vec3f v) { return v.x * l.vx + v.y * l.vy + v.z * l.vz; }
OSP_HD vec3f operator*(const linear3f &l, vec3f v) { return xfmVector(l, v); }
OSP_HD vec3f xfmVector(const affine3f &a, vec3f v) { return xfmVector(a.l, v); }
OSP_HD vec3f xfmPoint(const affine3f &a, vec3f v) { return xfmVector(a.l, v) + a.p; }
OSP_HD linear3f transposed(const linear3f &l)
{
  linear3f t;
  t.vx = make_vec3f(l.vx.x, l.vy.x, l.vz.x);
  t.vy = make_vec3f(l.vx.y, l.vy.y, l.vz.y);
  t.vz = make_vec3f(l.vx.z, l.vy.z, l.vz.z);
  return t;
}
OSP_HD linear3f operator*(const linear3f &a, const linear3f &b)
{
  linear3f r; r.vx = a * b.vx; r.vy = a * b.vy; r.vz = a * b.vz; return r;
}
OSP_HD float det(const linear3f &l) { return dot(l.vx, cross(l.vy, l.vz)); }
OSP_HD linear3f rcp(const linear3f &l)
{
  // adjoint / determinant (rkcommon LinearSpace3::inverse)
  const float d = det(l);
  linear3f adj;
  const vec3f c0 = cross(l.vy, l.vz), c1 = cross(l.vz, l.vx), c2 = cross(l.vx, l.vy);
  adj.vx = make_vec3f(c0.x, c1.x, c2.x);
  adj.vy = make_vec3f(c0.y, c1.y, c2.y);
  adj.vz = make_vec3f(c0.z, c1.z, c2.z);
  const float r = 1.f / d;
  adj.vx = adj.vx * r; adj.vy = adj.vy * r; adj.vz = adj.vz * r;
  return adj;
}
OSP_HD affine3f rcp(const affine3f &a)
{
  affine3f r; r.l = rcp(a.l); r.p = -(r.l * a.p); return r;
}
OSP_HD affine3f affine_identity()
{
  affine3f a;
  a.l.vx = make_vec3f(1, 0, 0); a.l.vy = make_vec3f(0, 1, 0); a.l.vz = make_vec3f(0, 0, 1);
  a.p = make_vec3f(0.f);
  return a;
}
OSP_HD affine3f operator*(const affine3f &a, const affine3f &b)
{
  affine3f r; r.l = a.l * b.l; r.p = xfmPoint(a, b.p); return r;
}
// build an orthonormal frame around N (rkcommon math/LinearSpace: frame())
OSP_HD linear3f frame(vec3f N)
{
  const vec3f dx0 = make_vec3f(0.f, N.z, -N.y);
  const vec3f dx1 = make_vec3f(-N.z, 0.f, N.x);
  const vec3f dx = normalize(dot(dx0, dx0) > dot(dx1, dx1) ? dx0 : dx1);
  const vec3f dy = normalize(cross(N, dx));
  linear3f f; f.vx = dx; f.vy = dy; f.vz = N; return f;
}

OSP_HD box3f box_empty()
{
  box3f b; b.lower = make_vec3f(INFINITY); b.upper = make_vec3f(-INFINITY); return b;
}
OSP_HD void box_extend(box3f &b, vec3f p) { b.lower = vmin(b.lower, p); b.upper = vmax(b.upper, p); }
OSP_HD void box_extend(box3f &b, const box3f &o) { b.lower = vmin(b.lower, o.lower); b.upper = vmax(b.upper, o.upper); }
OSP_HD bool box_is_empty(const box3f &b) { return b.lower.x > b.upper.x || b.lower.y > b.upper.y || b.lower.z > b.upper.z; }
OSP_HD box3f xfmBounds(const affine3f &m, const box3f &b)
{
  box3f dst = box_empty();
  for (int i = 0; i < 8; i++) {
    const vec3f c = make_vec3f((i & 1) ? b.upper.x : b.lower.x, (i & 2) ? b.upper.y : b.lower.y,
        (i & 4) ? b.upper.z : b.lower.z);
    box_extend(dst, xfmPoint(m, c));
  }
  return dst;
}

} // namespace ospm
