// Structured-regular volume sampling, transfer function and the SciVis / AO ray marchers.
//
// Stands in for Open VKL 2.0.1 (un-vendored): vklComputeSampleV (vertex-centred trilinear,
// NaN outside the grid), vklComputeGradientV (forward differences, step = gridSpacing) and the
// structured interval iterator (DDA over macrocells of 16^3 cells; a macrocell is reported when
// its value range overlaps a positive-opacity range of the transfer function; nominalDeltaT =
// min_i gridSpacing_i / |dir_i|) -- call sites modules/cpu/volume/Volume.ih:26-56 and
// render/scivis/volumes.ispc:302-439, render/ao/volumes.ispc (single-volume form).
// Transfer function: volume/transferFunction/LinearTransferFunction.ispc:11-43.
#pragma once
#include "shading.cuh"

namespace oc {

#define OC_MACROCELL_WIDTH 16

// nextafterf(u, 0.f) for the positive, finite u this file hands it (a grid dimension minus one, >= 1): the float below is the
// bit pattern minus one -- one instruction instead of the library routine, per sample (it was 6 % of the marcher's instructions)
__device__ __forceinline__ float floatBelow(float u) { return u > 0.f ? __uint_as_float(__float_as_uint(u) - 1u) : nextafterf(u, 0.f); }
__device__ __forceinline__ float volumeSample(const OcVolume &v, vec3f p)
{
  const vec3f lc = make_vec3f((p.x - v.gridOrigin.x) / v.gridSpacing.x, (p.y - v.gridOrigin.y) / v.gridSpacing.y,
      (p.z - v.gridOrigin.z) / v.gridSpacing.z);
  const float ux = (float)(v.dims[0] - 1), uy = (float)(v.dims[1] - 1), uz = (float)(v.dims[2] - 1);
  if (!(lc.x >= 0.f && lc.y >= 0.f && lc.z >= 0.f && lc.x <= ux && lc.y <= uy && lc.z <= uz))
    return NAN;
  const float cx = fminf(lc.x, floatBelow(ux)), cy = fminf(lc.y, floatBelow(uy)), cz = fminf(lc.z, floatBelow(uz));
  const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
  const uint32_t ix = (uint32_t)fx0, iy = (uint32_t)fy0, iz = (uint32_t)fz0;
  const float fx = cx - fx0, fy = cy - fy0, fz = cz - fz0;
  const size_t sx = 1, sy = v.dims[0], sz = (size_t)v.dims[0] * v.dims[1];
  const float *b = v.voxels + ix * sx + iy * sy + iz * sz;
  const float v000 = __ldg(b), v001 = __ldg(b + sx), v010 = __ldg(b + sy), v011 = __ldg(b + sy + sx);
  const float v100 = __ldg(b + sz), v101 = __ldg(b + sz + sx), v110 = __ldg(b + sz + sy), v111 = __ldg(b + sz + sy + sx);
  const float v00 = v000 + fx * (v001 - v000), v01 = v010 + fx * (v011 - v010);
  const float v10 = v100 + fx * (v101 - v100), v11 = v110 + fx * (v111 - v110);
  const float v0 = v00 + fy * (v01 - v00), v1 = v10 + fy * (v11 - v10);
  return v0 + fz * (v1 - v0);
}

__device__ __forceinline__ vec3f volumeGradient(const OcVolume &v, vec3f p)
{
  const float s = volumeSample(v, p);
  vec3f g;
  g.x = volumeSample(v, make_vec3f(p.x + v.gridSpacing.x, p.y, p.z)) - s;
  g.y = volumeSample(v, make_vec3f(p.x, p.y + v.gridSpacing.y, p.z)) - s;
  g.z = volumeSample(v, make_vec3f(p.x, p.y, p.z + v.gridSpacing.z)) - s;
  g = make_vec3f(g.x / v.gridSpacing.x, g.y / v.gridSpacing.y, g.z / v.gridSpacing.z);
  // Volume.ih:49-54: NaN guard
  if (isnan(g.x)) g.x = 1.f;
  if (isnan(g.y)) g.y = 1.f;
  if (isnan(g.z)) g.z = 1.f;
  return g;
}

__device__ __forceinline__ vec4f transferFunctionGet(const OcTransferFunction &tf, float value)
{
  value = (value - tf.valueLower) * (1.f / (tf.valueUpper - tf.valueLower));
  value = clampf(value, 0.0f, 0x1.fffffep-1f);
  const int maxIdxC = (int)tf.numColors - 1;
  const float idxCf = value * maxIdxC;
  const int idxC = (int)idxCf;
  const float fracC = idxCf - floorf(idxCf);
  const OcVec3f c0 = tf.color[idxC], c1 = tf.color[min(maxIdxC, idxC + 1)];
  const vec3f col = lerp3(fracC, make_vec3f(c0.x, c0.y, c0.z), make_vec3f(c1.x, c1.y, c1.z));
  const int maxIdxO = (int)tf.numOpacities - 1;
  const float idxOf = value * maxIdxO;
  const int idxO = (int)idxOf;
  const float fracO = idxOf - floorf(idxOf);
  const float opacity = lerpf(fracO, __ldg(tf.opacity + idxO), __ldg(tf.opacity + min(maxIdxO, idxO + 1)));
  return make_vec4f(col, opacity);
}

// macrocell interval iterator --------------------------------------------------------------
// The per-axis DDA state lives in named scalars, never in arrays indexed by the stepping axis: a dynamically indexed
// array is placed in local memory, and the marcher's loop then streams its iterator through L1/L2 into DRAM (C2 at 512^3:
// 1.77 GB of DRAM writes per frame against 83 MB of framebuffer writes, profiles/r2_c2_512_scivis_volume_ncu.txt).
struct VolumeIterator
{
  vec3f org, dir;          // volume-local ray
  float tEnd, tCur;
  int cellX, cellY, cellZ, stepX, stepY, stepZ;
  float tNextX, tNextY, tNextZ, tDeltaX, tDeltaY, tDeltaZ;
  float nominalDeltaT;
  bool done;
};

__device__ __forceinline__ bool rangesOverlap(const OcVolume &v, float lo, float hi)
{
  for (uint32_t i = 0; i < v.numRanges; i++)
    if (hi >= v.ranges[2 * i] && lo <= v.ranges[2 * i + 1]) return true;
  return false;
}

__device__ __forceinline__ void volumeIteratorAxis(float o, float d, float gs, float go, float tStart, int mcDim, float &nominalDeltaT,
    int &cell, int &step, float &tNext, float &tDelta)
{
  const float ad = fabsf(d);
  if (ad > 0.f) nominalDeltaT = fminf(nominalDeltaT, gs / ad);
  const float cw = gs * OC_MACROCELL_WIDTH;
  const float pos = (o + tStart * d - go) / cw;
  int c = (int)floorf(pos);
  c = max(0, min(c, mcDim - 1));
  cell = c;
  if (d > 0.f) {
    step = 1;
    tDelta = cw / d;
    tNext = (go + (c + 1) * cw - o) / d;
  } else if (d < 0.f) {
    step = -1;
    tDelta = -cw / d;
    tNext = (go + c * cw - o) / d;
  } else {
    step = 0;
    tDelta = INFINITY;
    tNext = INFINITY;
  }
}

__device__ __forceinline__ void volumeIteratorInit(VolumeIterator &it, const OcVolume &v, vec3f org, vec3f dir, float t0, float t1)
{
  it.org = org; it.dir = dir; it.tCur = t0; it.tEnd = t1; it.done = !(t0 < t1);
  it.nominalDeltaT = INFINITY;
  const float tStart = t0 + 1e-6f * fmaxf(1.f, fabsf(t0)); // nudge inside the first cell
  volumeIteratorAxis(org.x, dir.x, v.gridSpacing.x, v.gridOrigin.x, tStart, (int)v.mcDims[0], it.nominalDeltaT, it.cellX, it.stepX, it.tNextX, it.tDeltaX);
  volumeIteratorAxis(org.y, dir.y, v.gridSpacing.y, v.gridOrigin.y, tStart, (int)v.mcDims[1], it.nominalDeltaT, it.cellY, it.stepY, it.tNextY, it.tDeltaY);
  volumeIteratorAxis(org.z, dir.z, v.gridSpacing.z, v.gridOrigin.z, tStart, (int)v.mcDims[2], it.nominalDeltaT, it.cellZ, it.stepZ, it.tNextZ, it.tDeltaZ);
}

// returns the next macrocell interval [lower, upper] whose value range is of interest
__device__ __forceinline__ bool volumeIteratorNext(VolumeIterator &it, const OcVolume &v, float &lower, float &upper)
{
  while (!it.done) {
    const int a = (it.tNextX <= it.tNextY) ? ((it.tNextX <= it.tNextZ) ? 0 : 2) : ((it.tNextY <= it.tNextZ) ? 1 : 2);
    const float tNextA = a == 0 ? it.tNextX : (a == 1 ? it.tNextY : it.tNextZ);
    const float tExit = fminf(tNextA, it.tEnd);
    const size_t mc = it.cellX + (size_t)v.mcDims[0] * (it.cellY + (size_t)v.mcDims[1] * it.cellZ);
    const float lo = __ldg(v.mcRange + 2 * mc), hi = __ldg(v.mcRange + 2 * mc + 1);
    const float tEnter = it.tCur;
    // advance
    it.tCur = tExit;
    int cellA, dimA;
    if (a == 0) { it.cellX += it.stepX; it.tNextX += it.tDeltaX; cellA = it.cellX; dimA = (int)v.mcDims[0]; }
    else if (a == 1) { it.cellY += it.stepY; it.tNextY += it.tDeltaY; cellA = it.cellY; dimA = (int)v.mcDims[1]; }
    else { it.cellZ += it.stepZ; it.tNextZ += it.tDeltaZ; cellA = it.cellZ; dimA = (int)v.mcDims[2]; }
    if (tExit >= it.tEnd || cellA < 0 || cellA >= dimA) it.done = true;
    if (tExit > tEnter && rangesOverlap(v, lo, hi)) {
      lower = tEnter; upper = tExit;
      return true;
    }
  }
  return false;
}

// bounding-box interval of the first volume instance hit (Volume.ispc:22-73 + traceVolumeRay);
// single volume interval, as on the reference's GPU code path.
struct VolumeHit { const OcVolume *vol; const OcInstance *inst; uint32_t instID, geomID; float t0, t1; };

__device__ __forceinline__ bool traceVolumeRay(const OcWorld *world, const Ray &ray, VolumeHit &vh)
{
  vh.vol = nullptr;
  float best = INFINITY;
  for (uint32_t i = 0; i < world->numInstances; i++) {
    const OcInstance *inst = world->instances + i;
    const OcGroup *g = inst->group;
    if (g->numVolumes == 0) continue;
    vec3f org = ray.org, dir = ray.dir;
    if (!inst->isIdentity) {
      const affine3f rx = toA(inst->rcp_xfm);
      org = xfmPoint(rx, ray.org); dir = xfmVector(rx, ray.dir);
    }
    for (uint32_t k = 0; k < g->numVolumes; k++) {
      const OcVolume *v = g->volumes + k;
      const vec3f rd = make_vec3f(rcp_safe(dir.x), rcp_safe(dir.y), rcp_safe(dir.z));
      const vec3f mins = make_vec3f((v->bounds[0] - org.x) * rd.x, (v->bounds[1] - org.y) * rd.y, (v->bounds[2] - org.z) * rd.z);
      const vec3f maxs = make_vec3f((v->bounds[3] - org.x) * rd.x, (v->bounds[4] - org.y) * rd.y, (v->bounds[5] - org.z) * rd.z);
      const float tEntry = reduce_max(vmin(mins, maxs)), tExit = reduce_min(vmax(mins, maxs));
      const float tmin = fmaxf(ray.t0, tEntry), tmax = fminf(ray.t, tExit);
      if (tmin < tmax && tmin < best) {
        best = tmin;
        vh.vol = v; vh.inst = inst; vh.instID = i; vh.geomID = k; vh.t0 = tmin; vh.t1 = tmax;
      }
    }
  }
  return vh.vol != nullptr;
}

} // namespace oc
