// ref_vkl_stub.cpp -- see ref_vkl_stub.h.  TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libospref.so).
#include <cmath>
#include <cstddef>
#include "ref_vkl_stub.h"

namespace {
float sampleAt(const OcVolume &v, float px, float py, float pz)
{
  const float lx = (px - v.gridOrigin.x) / v.gridSpacing.x, ly = (py - v.gridOrigin.y) / v.gridSpacing.y, lz = (pz - v.gridOrigin.z) / v.gridSpacing.z;
  const float ux = (float)(v.dims[0] - 1), uy = (float)(v.dims[1] - 1), uz = (float)(v.dims[2] - 1);
  if (!(lx >= 0.f && ly >= 0.f && lz >= 0.f && lx <= ux && ly <= uy && lz <= uz)) return NAN;
  const float cx = fminf(lx, nextafterf(ux, 0.f)), cy = fminf(ly, nextafterf(uy, 0.f)), cz = fminf(lz, nextafterf(uz, 0.f));
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
struct IterState { vkl_range1f tRange; float nominalDeltaT; int done; };
} // namespace

float vklComputeSample(const VKLSampler *s, const vkl_vec3f *p, unsigned int, float, VKLFeatureFlags)
{
  return sampleAt(*((const RefVklSampler *)s->host)->v, p->x, p->y, p->z);
}

vkl_vec3f vklComputeGradient(const VKLSampler *s, const vkl_vec3f *p, unsigned int, float, VKLFeatureFlags)
{
  const OcVolume &v = *((const RefVklSampler *)s->host)->v;
  const float c = sampleAt(v, p->x, p->y, p->z);
  vkl_vec3f g = {sampleAt(v, p->x + v.gridSpacing.x, p->y, p->z) - c, sampleAt(v, p->x, p->y + v.gridSpacing.y, p->z) - c,
      sampleAt(v, p->x, p->y, p->z + v.gridSpacing.z) - c};
  g.x /= v.gridSpacing.x; g.y /= v.gridSpacing.y; g.z /= v.gridSpacing.z;
  return g;
}

VKLIntervalIterator vklInitIntervalIterator(const VKLIntervalIteratorContext *ctx, const vkl_vec3f *, const vkl_vec3f *dir,
    const vkl_range1f *tRange, float, void *buffer, VKLFeatureFlags)
{
  const OcVolume &v = *((const RefVklIntervalContext *)ctx->host)->v;
  IterState *st = (IterState *)buffer;
  st->tRange = *tRange;
  float n = INFINITY;
  if (fabsf(dir->x) > 0.f) n = fminf(n, v.gridSpacing.x / fabsf(dir->x));
  if (fabsf(dir->y) > 0.f) n = fminf(n, v.gridSpacing.y / fabsf(dir->y));
  if (fabsf(dir->z) > 0.f) n = fminf(n, v.gridSpacing.z / fabsf(dir->z));
  st->nominalDeltaT = n;
  st->done = !(tRange->lower < tRange->upper);
  return st;
}

int vklIterateInterval(VKLIntervalIterator it, VKLInterval *interval, VKLFeatureFlags)
{
  IterState *st = (IterState *)it;
  if (st->done) return 0;
  st->done = 1;
  interval->tRange = st->tRange;
  interval->valueRange = {-INFINITY, INFINITY};
  interval->nominalDeltaT = st->nominalDeltaT;
  return 1;
}
