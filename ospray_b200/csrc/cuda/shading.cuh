// Hit -> DifferentialGeometry, OBJ material (3-lobe MultiBSDF evaluated in registers), lights.
// reference: common/World.ih:389-465 (postIntersect), common/Instance.ih:44-105,
// geometry/GeometricModel.ih:21-76, geometry/Mesh.ispc:65-268, Spheres.ispc:24-42,
// Boxes.ispc:65-71, common/DifferentialGeometry.ih:46-58, render/Material.ih:25-37,
// render/materials/OBJ.ispc:29-144, render/bsdfs/{Lambert,Specular,Transmission,MultiBSDF}.ih,
// lights/QuadLight.ispc:15-215, lights/SphericalQuadSampling.ih:10-204,
// lights/AmbientLight.ispc:17-57, lights/DirectionalLight.ispc:14-80, math/sampling.ih.
#pragma once
#include "intersect.cuh"
#include "samplers.cuh"
#include "texture.cuh"
#include <cfloat>

namespace oc {

#define OC_PI 3.14159265358979323846f
#define OC_TWO_PI 6.28318530717958647692f
#define OC_ONE_OVER_PI 0.31830988618379067154f
#define OC_ONE_OVER_TWO_PI 0.15915494309189533577f
#define OC_ONE_OVER_FOUR_PI 0.079577471545947667884f
#define OC_FOUR_PI 12.566370614359172954f
#define OC_LNTWO_OVER_TWO 0.34657359027997265471f
#define OC_ULP_EPSILON 0x1.0fp-21f

enum {
  DG_FACEFORWARD = 1 << 0, DG_NORMALIZE = 1 << 1, DG_NG = 1 << 2, DG_NS = 1 << 3,
  DG_COLOR = 1 << 4, DG_TEXCOORD = 1 << 5, DG_TANGENTS = 1 << 6, DG_MOTIONBLUR = 1 << 7
};

struct DifferentialGeometry
{
  vec3f P, Ng, Ns, dPds;
  vec4f color;
  float epsilon;
  uint32_t primID, objID, instID;
  const OcMaterial *material;
  // only filled by postIntersect<true> (frames whose world / renderer reference textures)
  vec3f dPdt;
  vec2f st;
  float pixelFootprint;
  float areaPDF; // geometry lights: 1 / emissive area of the model (GeometricModel.ih:58)
};
// common/RayCone.ih: width at the current vertex, spread per unit distance
struct RayCone { float width, dwdt; };

__device__ __forceinline__ float calcEpsilon(vec3f P, float t)
{
  return fmaxf(fmaxf(fabsf(P.x), fabsf(P.y)), fmaxf(fabsf(P.z), t)) * OC_ULP_EPSILON;
}
__device__ __forceinline__ float calcEpsilon(vec3f P, vec3f dir, float t)
{
  return calcEpsilon(P, reduce_max(vabs(dir)) * t);
}

// strided fetches (reference: common/Data.ih get_* accessors)
__device__ __forceinline__ vec3f get3f(const OcData1D &d, uint32_t i)
{
  const float *p = reinterpret_cast<const float *>(d.addr + (int64_t)i * d.byteStride);
  return make_vec3f(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
__device__ __forceinline__ vec4f get4f(const OcData1D &d, uint32_t i)
{
  const float *p = reinterpret_cast<const float *>(d.addr + (int64_t)i * d.byteStride);
  return make_vec4f(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
__device__ __forceinline__ uint32_t get1ui(const OcData1D &d, uint32_t i, int c)
{
  return __ldg(reinterpret_cast<const uint32_t *>(d.addr + (int64_t)i * d.byteStride) + c);
}
__device__ __forceinline__ vec4f getColor(const OcGeometry *g, uint32_t i)
{
  if (g->flags & OC_GEOMFLAG_COLOR_IS_VEC4) return get4f(g->color, i);
  return make_vec4f(get3f(g->color, i), 1.f);
}

__device__ __forceinline__ const OcMaterial *getMaterial(const OcWorld *world, const OcRenderer &ren,
    const OcModel *m, uint32_t primID)
{
  if (m->materialMode == OC_MAT_NONE || m->numMaterials == 0)
    return nullptr;
  uint32_t idx = primID;
  if (m->index.addr)
    idx = __ldg(m->index.addr + (int64_t)min(idx, m->index.numItems - 1) * m->index.byteStride);
  uint32_t matIdx = __ldg(m->material + min(idx, m->numMaterials - 1));
  if (m->materialMode == OC_MAT_FROM_RENDERER) {
    if (!ren.materials || ren.numMaterials == 0) return nullptr;
    matIdx = __ldg(ren.materials + min(matIdx, ren.numMaterials - 1));
  }
  return world->materials + matIdx;
}

__device__ __forceinline__ vec2f get2f(const OcData1D &d, uint32_t i)
{
  const float *p = reinterpret_cast<const float *>(d.addr + (int64_t)i * d.byteStride);
  return make_vec2f(__ldg(p), __ldg(p + 1));
}
// texture-space tangents of a triangle / of a quad's (a, b, d) corner (Mesh.ispc:152-168, :253-268); false: degenerate
__device__ __forceinline__ bool meshTexTangents(DifferentialGeometry &dg, vec2f a, vec2f b, vec2f c, vec3f va, vec3f vb, vec3f vc)
{
  const vec2f dst02 = a - c, dst12 = b - c;
  const float det = dst02.x * dst12.y - dst02.y * dst12.x;
  if (det == 0.f) return false;
  const float invDet = 1.f / det;
  const vec3f dp02 = va - vc, dp12 = vb - vc;
  dg.dPds = (dst12.y * dp02 - dst02.y * dp12) * invDet;
  dg.dPdt = (dst02.x * dp12 - dst12.x * dp02) * invDet;
  return true;
}

template <bool TEX = false>
__device__ __forceinline__ void geometryPostIntersect(const OcGeometry *g, DifferentialGeometry &dg,
    const Ray &ray, uint32_t flags)
{
  dg.Ng = dg.Ns = ray.Ng;
  if (g->type == OC_GEOM_QUADS) {
    const uint32_t i0 = get1ui(g->index, ray.primID, 0), i1 = get1ui(g->index, ray.primID, 1),
                   i2 = get1ui(g->index, ray.primID, 2), i3 = get1ui(g->index, ray.primID, 3);
    const uint32_t fo = ray.primID * 4;
    const float u = ray.u, v = ray.v;
    vec4f uv;
    if (i2 == i3) uv = make_vec4f(1.f - u - v, u, v, 0.f);
    else uv = make_vec4f((1 - v) * (1 - u), (1 - v) * u, v * u, v * (1 - u));
    if ((flags & DG_NS) && g->normal.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_NORMAL_FACEVARYING;
      const vec3f a = get3f(g->normal, fv ? fo : i0), b = get3f(g->normal, fv ? fo + 1 : i1),
                  c = get3f(g->normal, fv ? fo + 2 : i2), d = get3f(g->normal, fv ? fo + 3 : i3);
      dg.Ns = uv.x * a + uv.y * b + uv.z * c + uv.w * d;
    }
    if ((flags & DG_COLOR) && g->color.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_COLOR_FACEVARYING;
      const vec4f a = getColor(g, fv ? fo : i0), b = getColor(g, fv ? fo + 1 : i1),
                  c = getColor(g, fv ? fo + 2 : i2), d = getColor(g, fv ? fo + 3 : i3);
      dg.color = uv.x * a + uv.y * b + uv.z * c + uv.w * d;
      if (!(g->flags & OC_GEOMFLAG_HAS_ALPHA)) dg.color.w = 1.f;
    }
    const vec3f va = get3f(g->vertex, i0), vb = get3f(g->vertex, i1), vc = get3f(g->vertex, i2),
                vd = get3f(g->vertex, i3);
    dg.epsilon = reduce_max(vmax(vmax(vabs(va), vabs(vb)), vmax(vabs(vc), vabs(vd)))) * OC_ULP_EPSILON;
    { // texture coordinates (TEX only) and tangents: along texture space when the mesh has texcoords, else along its edges
      bool tangents = flags & DG_TANGENTS;
      if ((tangents || (TEX && (flags & DG_TEXCOORD))) && g->texcoord.addr) {
        const bool fv = g->flags & OC_GEOMFLAG_TEXCOORD_FACEVARYING;
        const vec2f a = get2f(g->texcoord, fv ? fo : i0), b = get2f(g->texcoord, fv ? fo + 1 : i1), d = get2f(g->texcoord, fv ? fo + 3 : i3);
        if (TEX && (flags & DG_TEXCOORD)) {
          const vec2f c = get2f(g->texcoord, fv ? fo + 2 : i2);
          dg.st = uv.x * a + uv.y * b + uv.z * c + uv.w * d;
        }
        if (tangents && meshTexTangents(dg, a, b, d, va, vb, vd)) tangents = false;
      }
      if (tangents) { dg.dPds = vb - va; dg.dPdt = vd - va; } // (Mesh.ispc:170-173)
    }
  } else if (g->type == OC_GEOM_TRIANGLES) {
    const uint32_t i0 = get1ui(g->index, ray.primID, 0), i1 = get1ui(g->index, ray.primID, 1),
                   i2 = get1ui(g->index, ray.primID, 2);
    const uint32_t fo = ray.primID * 3;
    const vec3f uv = make_vec3f(1.0f - ray.u - ray.v, ray.u, ray.v);
    if ((flags & DG_NS) && g->normal.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_NORMAL_FACEVARYING;
      const vec3f a = get3f(g->normal, fv ? fo : i0), b = get3f(g->normal, fv ? fo + 1 : i1),
                  c = get3f(g->normal, fv ? fo + 2 : i2);
      dg.Ns = uv.x * a + uv.y * b + uv.z * c;
    }
    if ((flags & DG_COLOR) && g->color.addr) {
      const bool fv = g->flags & OC_GEOMFLAG_COLOR_FACEVARYING;
      const vec4f a = getColor(g, fv ? fo : i0), b = getColor(g, fv ? fo + 1 : i1), c = getColor(g, fv ? fo + 2 : i2);
      dg.color = uv.x * a + uv.y * b + uv.z * c;
      if (!(g->flags & OC_GEOMFLAG_HAS_ALPHA)) dg.color.w = 1.f;
    }
    const vec3f va = get3f(g->vertex, i0), vb = get3f(g->vertex, i1), vc = get3f(g->vertex, i2);
    dg.epsilon = reduce_max(vmax(vmax(vabs(va), vabs(vb)), vabs(vc))) * OC_ULP_EPSILON;
    {
      bool tangents = flags & DG_TANGENTS;
      if ((tangents || (TEX && (flags & DG_TEXCOORD))) && g->texcoord.addr) {
        const bool fv = g->flags & OC_GEOMFLAG_TEXCOORD_FACEVARYING;
        const vec2f a = get2f(g->texcoord, fv ? fo : i0), b = get2f(g->texcoord, fv ? fo + 1 : i1), c = get2f(g->texcoord, fv ? fo + 2 : i2);
        if (TEX && (flags & DG_TEXCOORD)) dg.st = uv.x * a + uv.y * b + uv.z * c;
        if (tangents && meshTexTangents(dg, a, b, c, va, vb, vc)) tangents = false;
      }
      if (tangents) { dg.dPds = vb - va; dg.dPdt = vc - va; }
    }
  } else if (g->type == OC_GEOM_SPHERES) {
    dg.epsilon = get4f(g->sphere, ray.primID).w * OC_ULP_EPSILON;
    if (TEX && (flags & DG_TEXCOORD) && g->texcoord.addr) dg.st = get2f(g->texcoord, ray.primID); // Spheres.ispc:39-41
  } else if (g->type == OC_GEOM_CURVES) {
    // Curves_postIntersect + interpolate<vec4f> (geometry/Curves.ispc:17-94): vertex colours through the curve basis at ray.u
    if ((flags & DG_COLOR) && g->color.addr) {
      const uint32_t index = get1ui(g->index, ray.primID, 0);
      const float t = ray.u;
      const vec4f v0 = get4f(g->color, index + 0), v1 = get4f(g->color, index + 1);
      if (g->curveBasis == OC_CURVE_BASIS_LINEAR)
        dg.color = (1.f - t) * v0 + t * v1;
      else {
        const vec4f v2 = get4f(g->color, index + 2), v3 = get4f(g->color, index + 3);
        const float s = 1.0f - t, s2 = s * s, t2 = t * t, s2t = s2 * t, t2s = t2 * s, s3 = s2 * s, t3 = t2 * t;
        float b0, b1, b2, b3, cf;
        if (g->curveBasis == OC_CURVE_BASIS_BEZIER) { b0 = s3; b1 = 3.0f * s2t; b2 = 3.0f * t2s; b3 = t3; cf = 1.0f; }
        else if (g->curveBasis == OC_CURVE_BASIS_BSPLINE) {
          b0 = s3; b1 = (4.0f * s3 + t3) + (12.0f * s2t + 6.0f * t2s); b2 = (4.0f * t3 + s3) + (12.0f * t2s + 6.0f * s2t); b3 = t3; cf = 1.0f / 6.0f;
        } else { b0 = -s2t; b1 = 2.0f + t2 * (3.0f * t - 5.0f); b2 = 2.0f + s2 * (3.0f * s - 5.0f); b3 = -t2s; cf = 0.5f; }
        dg.color = cf * (b0 * v0 + b1 * v1 + b2 * v2 + b3 * v3);
      }
    }
  }
  // boxes: nothing else (Boxes.ispc:65-71)
}

// World.ih:389-465. TEX: also texture coordinates, both tangents and the texture-space footprint of the ray cone
// (World.ih:459-464), which is propagated to the hit
template <bool TEX = false>
__device__ __forceinline__ void postIntersect(const OcWorld *world, const OcRenderer &ren,
    DifferentialGeometry &dg, const Ray &ray, uint32_t flags, RayCone *rayCone = nullptr)
{
  dg.primID = ray.primID;
  dg.material = nullptr;
  if (flags & DG_COLOR) dg.color = make_vec4f(1.f, 1.f, 1.f, 1.f);
  if (flags & DG_TANGENTS) dg.dPds = make_vec3f(1.f, 0.f, 0.f);
  if (TEX) {
    dg.st = make_vec2f(ray.u, ray.v);
    if (flags & DG_TANGENTS) dg.dPdt = make_vec3f(0.f, 1.f, 0.f);
  }
  dg.P = ray.org + ray.t * ray.dir;
  dg.epsilon = 0.f;

  if (ray.instID != OC_INVALID_ID) {
    const OcInstance *inst = world->instances + ray.instID;
    const OcModel *m = inst->group->models + ray.geomID;
    geometryPostIntersect<TEX>(m->geom, dg, ray, flags);
    dg.objID = (m->userID == OC_INVALID_ID) ? ray.geomID : m->userID;
    if ((flags & DG_COLOR) && m->color.addr) {
      uint32_t idx = ray.primID;
      if (m->index.addr)
        idx = __ldg(m->index.addr + (int64_t)min(idx, m->index.numItems - 1) * m->index.byteStride);
      dg.color = get4f(m->color, min(idx, m->color.numItems - 1));
    }
    dg.material = getMaterial(world, ren, m, ray.primID);
    if (TEX) dg.areaPDF = inst->modelAreaPDF ? __ldg(inst->modelAreaPDF + ray.geomID) : 0.f;
    if ((flags & DG_NS) && m->invertedNormals) dg.Ns = neg(dg.Ns);
    dg.instID = (inst->userID == OC_INVALID_ID) ? ray.instID : inst->userID;
    if (!inst->isIdentity) {
      const affine3f xfm = toA(inst->xfm), rxfm = toA(inst->rcp_xfm);
      const linear3f rt = transposed(rxfm.l);
      dg.Ns = xfmVector(rt, dg.Ns);
      dg.Ng = xfmVector(rt, dg.Ng);
      dg.epsilon *= fmaxf(fabsf(xfm.l.vx.x), fmaxf(fabsf(xfm.l.vy.y), fabsf(xfm.l.vz.z)));
      if (flags & DG_TANGENTS) {
        dg.dPds = xfmVector(xfm, dg.dPds);
        if (TEX) dg.dPdt = xfmVector(xfm, dg.dPdt);
      }
    }
  } else {
    dg.Ns = dg.Ng = ray.Ng;
    dg.objID = OC_INVALID_ID; dg.instID = OC_INVALID_ID;
  }

  dg.epsilon = fmaxf(dg.epsilon, calcEpsilon(dg.P, ray.dir, ray.t));

  vec3f ffnng = normalize(dg.Ng);
  if ((flags & (DG_NG | DG_NORMALIZE)) == (DG_NG | DG_NORMALIZE)) dg.Ng = ffnng;
  if ((flags & (DG_NS | DG_NORMALIZE)) == (DG_NS | DG_NORMALIZE)) dg.Ns = normalize(dg.Ns);
  const bool flip = dot(ray.dir, dg.Ng) >= 0.f;
  if (flip) ffnng = neg(ffnng);
  if ((flags & (DG_NG | DG_FACEFORWARD)) == (DG_NG | DG_FACEFORWARD) && flip) dg.Ng = neg(dg.Ng);
  if ((flags & (DG_NS | DG_FACEFORWARD)) == (DG_NS | DG_FACEFORWARD)) {
    if (dot(dg.Ng, dg.Ns) < 0.f) dg.Ns = neg(dg.Ns);
  }
  dg.P = dg.P + dg.epsilon * ffnng;
  if (TEX) {
    float width = 0.f;
    if (rayCone) { rayCone->width += rayCone->dwdt * ray.t; width = rayCone->width; }
    const float texArea = length(cross(dg.dPds, dg.dPdt));
    const float world2tex = 1.0f / sqrtf(texArea);
    const float projConeWidth = width / -dot(ray.dir, dg.Ng);
    dg.pixelFootprint = projConeWidth * world2tex * ren.mipBiasFactor;
  }
}
template <bool TEX = false>
static __device__ __noinline__ void postIntersectNI(const OcWorld *world, const OcRenderer &ren, DifferentialGeometry &dg,
    const Ray &ray, uint32_t flags, RayCone *rayCone = nullptr)
{
  postIntersect<TEX>(world, ren, dg, ray, flags, rayCone);
}

// ---- sampling helpers (math/sampling.ih) ---------------------------------------------------
__device__ __forceinline__ vec3f cartesian(float phi, float sinTheta, float cosTheta)
{
  float s, c;
  sincosf(phi, &s, &c);
  return make_vec3f(c * sinTheta, s * sinTheta, cosTheta);
}
__device__ __forceinline__ float cos2sin(float c) { return sqrtf(fmaxf(0.f, 1.f - c * c)); }
__device__ __forceinline__ vec3f cosineSampleHemisphere(vec2f s)
{
  return cartesian(OC_TWO_PI * s.x, sqrtf(1.0f - s.y), sqrtf(s.y));
}
__device__ __forceinline__ vec3f powerCosineSampleHemisphere(float n, vec2f s)
{
  const float cosTheta = powf(s.y, 1.0f / (n + 1.0f));
  return cartesian(OC_TWO_PI * s.x, cos2sin(cosTheta), cosTheta);
}
__device__ __forceinline__ float powerCosineSampleHemispherePDF(float cosTheta, float n)
{
  return (n + 1.0f) * OC_ONE_OVER_TWO_PI * powf(cosTheta, n);
}
__device__ __forceinline__ vec3f uniformSampleCone(float cosAngle, vec2f s)
{
  const float cosTheta = 1.0f - s.y * (1.0f - cosAngle);
  return cartesian(OC_TWO_PI * s.x, cos2sin(cosTheta), cosTheta);
}
__device__ __forceinline__ vec3f uniformSampleSphere(float radius, vec2f s)
{
  return cartesian(OC_TWO_PI * s.x, 2.f * radius * sqrtf(s.y * (1.f - s.y)), radius * (1.f - 2.f * s.y));
}

// ---- OBJ material: Lambert + Transmission + Specular MultiBSDF, in registers ----------------
enum {
  SC_SPECULAR_REFLECTION = 1 << 0, SC_GLOSSY_REFLECTION = 1 << 1, SC_DIFFUSE_REFLECTION = 1 << 2,
  SC_SPECULAR_TRANSMISSION = 1 << 3, SC_GLOSSY_TRANSMISSION = 1 << 4, SC_DIFFUSE_TRANSMISSION = 1 << 5,
  SC_DIFFUSE = SC_DIFFUSE_REFLECTION | SC_DIFFUSE_TRANSMISSION,
  SC_GLOSSY = SC_GLOSSY_REFLECTION | SC_GLOSSY_TRANSMISSION,
  SC_SPECULAR = SC_SPECULAR_REFLECTION | SC_SPECULAR_TRANSMISSION,
  SC_REFLECTION = SC_DIFFUSE_REFLECTION | SC_GLOSSY_REFLECTION | SC_SPECULAR_REFLECTION,
  SC_TRANSMISSION = SC_DIFFUSE_TRANSMISSION | SC_GLOSSY_TRANSMISSION | SC_SPECULAR_TRANSMISSION,
  SC_SMOOTH = SC_DIFFUSE | SC_GLOSSY
};

struct EvalRes { vec3f value; float pdf; };
struct SampleRes { vec3f weight; vec3f wi; float pdf; uint32_t type; };

struct ObjBSDF
{
  linear3f frame;          // shading frame (vz = N)
  vec3f Kd, T, ks, specWeight;
  float ns;                // Blinn exponent
  float impL, impT, impS, importanceSum;
  vec3f albedo;
  uint32_t scatteringType;
  bool valid;
  bool thinGlass;          // RobustThinDielectric (materials/ThinGlass.ispc): eta in `ns`, attenuation in `Kd`
  bool pureTransmission;   // a lone Transmission BSDF (materials/Luminous.ispc:18-35): T in `T`
  uint32_t kind;           // BSDF_KIND_*: 0 = the cases above; the others only exist in the extended-material kernels
};
enum { BSDF_KIND_OBJ = 0, BSDF_KIND_CONDUCTOR, BSDF_KIND_MICROFACET_CONDUCTOR, BSDF_KIND_PLASTIC_SMOOTH, BSDF_KIND_PLASTIC_ROUGH,
       BSDF_KIND_PRINCIPLED, BSDF_KIND_GLASS, BSDF_KIND_VELVET, BSDF_KIND_METALLICPAINT, BSDF_KIND_CARPAINT };

// render/Material.ih:25-37
__device__ __forceinline__ linear3f makeShadingFrame(const DifferentialGeometry &dg)
{
  const vec3f N = dg.Ns;
  vec3f U = normalize(dg.dPds);
  vec3f V = cross(N, U);
  if (dot(V, V) > 0.f) {
    V = normalize(V);
    U = cross(V, N);
    linear3f f; f.vx = U; f.vy = V; f.vz = N; return f;
  }
  return frame(N);
}

// render/Material.ih:39-81: shading frame from the bump (normal) map, rotated in tangent space by map_bump's 2D rotation
__device__ __forceinline__ linear3f makeShadingFrameBump(const DifferentialGeometry &dg, const OcMaterial *m)
{
  linear3f f; f.vx = normalize(dg.dPds); f.vy = normalize(dg.dPdt); f.vz = dg.Ns;
  const vec4f t = texParamGet4f(m->mapBump, dg.st, dg.pixelFootprint);
  vec3f localN = make_vec3f(t.x, t.y, t.z) * 2.f - make_vec3f(1.f); // getNormal
  const vec2f rotN = make_vec2f(m->bumpRot[0] * localN.x + m->bumpRot[2] * localN.y, m->bumpRot[1] * localN.x + m->bumpRot[3] * localN.y);
  localN.x = rotN.x; localN.y = rotN.y;
  const vec3f N = normalize(f * localN);
  const vec3f U0 = f.vx;
  vec3f V = cross(N, U0);
  if (dot(V, V) > 0.f) {
    V = normalize(V);
    linear3f r; r.vx = cross(V, N); r.vy = V; r.vz = N; return r;
  }
  return frame(N);
}

// ---- thin glass: bsdfs/Fresnel.ih:19-47, Optics.ih:13-28, RobustThinDielectric.ih, materials/ThinGlass.ispc ----
__device__ __forceinline__ float sqrCosT(float cosI, float eta) { return 1.0f - (eta * eta) * (1.0f - cosI * cosI); }
__device__ __forceinline__ float fresnelDielectric3(float cosI, float cosT, float eta)
{
  const float Rper = (eta * cosI - cosT) * (1.f / (eta * cosI + cosT));
  const float Rpar = (cosI - eta * cosT) * (1.f / (cosI + eta * cosT));
  return 0.5f * (Rpar * Rpar + Rper * Rper);
}
__device__ __forceinline__ float fresnelDielectric(float cosI, float eta)
{
  const float sqrcost = sqrCosT(cosI, eta);
  if (sqrcost < 0.0f) return 1.0f;
  return fresnelDielectric3(cosI, sqrtf(sqrcost), eta);
}
__device__ __forceinline__ float fresnelDielectricEx(float cosI, float &cosT, float eta)
{
  const float sqrcost = sqrCosT(cosI, eta);
  if (sqrcost < 0.0f) { cosT = 0.0f; return 1.0f; }
  cosT = sqrtf(sqrcost);
  return fresnelDielectric3(cosI, cosT, eta);
}
__device__ __forceinline__ vec3f expf3(vec3f v) { return make_vec3f(expf(v.x), expf(v.y), expf(v.z)); }
__device__ __forceinline__ vec3f logf3(vec3f v) { return make_vec3f(logf(v.x), logf(v.y), logf(v.z)); }
__device__ __forceinline__ vec3f sqr3(vec3f v) { return v * v; }
__device__ __forceinline__ vec3f rcp3(vec3f v) { return make_vec3f(1.f / v.x, 1.f / v.y, 1.f / v.z); }
} // namespace oc
#include "microfacet.cuh"
#include "principled.cuh"
namespace oc {
// state of the microfacet-family BSDFs (metal, plastic, principled): kept apart from ObjBSDF so that the OBJ path's registers
// stay as they are
struct ExtBSDF
{
  ConductorBSDF cond; LayerParams layer; LambertSub lambert; PrBSDF pr;
  vec3f glassNg, glassNs; float glassEta; // RobustDielectric (bsdfs/RobustDielectric.ih:14-25)
  VelvetBSDF velvet;
  MetallicPaintSub paint;                 // under `layer`
  CarPaintBSDF car;
};
// the medium a path travels in (render/materials/MediumShared.h:10-18; vacuum = no attenuation, ior 1)
struct Medium { vec3f attenuation; float ior; };
__device__ __forceinline__ Medium mediumVacuum() { Medium m; m.attenuation = make_vec3f(0.f); m.ior = 1.f; return m; }
__device__ __forceinline__ bool mediumIsOutside(const Medium &a, const OcPrincipled &q) // eq(currentMedium, self->outsideMedium)
{
  return a.ior == q.outsideIor && a.attenuation.x == q.outsideAttenuation.x && a.attenuation.y == q.outsideAttenuation.y
      && a.attenuation.z == q.outsideAttenuation.z;
}

// ThinGlass.ispc:17-30
template <bool TEX = false>
__device__ __forceinline__ vec3f ThinGlass_getAttenuation(const OcMaterial *m, const DifferentialGeometry &dg)
{
  vec3f attenuation = toV(m->attenuation) + logf3(make_vec3f(dg.color)) * m->attenuationScale;
  if (TEX && m->mapKd.tex) { // map_attenuationColor
    const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint);
    attenuation = attenuation + logf3(make_vec3f(t.x, t.y, t.z)) * m->attenuationScale;
  }
  return attenuation;
}
// RobustThinDielectric.ih:99-117
__device__ __forceinline__ vec3f RobustThinDielectric_getTransparency(float cosThetaO, float eta, vec3f attenuation)
{
  float cosThetaT;
  const float Fo = fresnelDielectricEx(cosThetaO, cosThetaT, eta);
  const float Ft = fresnelDielectric(cosThetaT, 1.f / eta);
  const float length = 1.f / cosThetaT;
  const vec3f A = expf3(attenuation * length);
  return (1.f - Fo) * A * (1.f - Ft) * (make_vec3f(1.f) + sqr3(Ft * A) * rcp3(make_vec3f(1.f) - sqr3(Ft * A)));
}

// materials/OBJ.ispc:29-91, materials/ThinGlass.ispc:32-50. A NULL / unsupported material yields no BSDF (path
// terminates).
template <bool TEX = false>
__device__ __forceinline__ void OBJ_getBSDF(ObjBSDF &b, const DifferentialGeometry &dg)
{
  const OcMaterial *m = dg.material;
  b.valid = false;
  b.thinGlass = false;
  b.pureTransmission = false;
  b.kind = BSDF_KIND_OBJ;
  if (TEX && m && m->type == OC_MATERIAL_LUMINOUS) { // Luminous_getBSDF: no BSDF unless the emitter is transparent
    if (m->transparency > 0.0f) {
      b.valid = true;
      b.pureTransmission = true;
      b.frame = frame(dg.Ns);
      b.T = make_vec3f(m->transparency);
      b.Kd = make_vec3f(0.f);
      b.albedo = b.T;
      b.scatteringType = SC_SPECULAR_TRANSMISSION;
      b.importanceSum = 1.f;
      b.impL = b.impT = b.impS = 0.f;
      b.ns = 0.f;
    }
    return;
  }
  if (m && m->type == OC_MATERIAL_THINGLASS) {
    b.valid = true;
    b.thinGlass = true;
    b.frame = frame(dg.Ns);
    b.Kd = ThinGlass_getAttenuation<TEX>(m, dg);
    b.ns = m->eta;
    b.albedo = expf3(b.Kd);
    b.scatteringType = SC_SPECULAR;
    b.importanceSum = 1.f;
    b.impL = b.impT = b.impS = 0.f;
    return;
  }
  if (!m || m->type != OC_MATERIAL_OBJ)
    return;
  b.valid = true;
  const bool tex = TEX && m->hasTextures;
  b.frame = (tex && m->mapBump.tex) ? makeShadingFrameBump(dg, m) : makeShadingFrame(dg);
  b.albedo = make_vec3f(0.f);
  b.scatteringType = 0;
  b.importanceSum = 0.f;
  b.impL = b.impT = b.impS = 0.f;
  float d = m->d;
  if (tex && m->mapD.tex) d = d * texParamGet4f(m->mapD, dg.st, dg.pixelFootprint).x;
  d = d * dg.color.w;
  vec3f Kd = toV(m->Kd);
  if (tex && m->mapKd.tex) {
    const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint);
    Kd = Kd * make_vec3f(t.x, t.y, t.z);
    d *= t.w;
  }
  Kd = Kd * d * make_vec3f(dg.color);
  b.Kd = Kd;
  if (reduce_max(Kd) > 0.0f) {
    b.albedo = b.albedo + Kd;
    b.scatteringType |= SC_DIFFUSE_REFLECTION;
    b.impL = luminance(Kd);
    b.importanceSum += b.impL;
  }
  const vec3f T = toV(m->Tf) * d + make_vec3f(1.f - d);
  b.T = T;
  if (reduce_max(T) > 0.0f) {
    b.albedo = b.albedo + T;
    b.scatteringType |= SC_SPECULAR_TRANSMISSION;
    b.impT = luminance(T);
    b.importanceSum += b.impT;
  }
  float Ns = m->Ns;
  vec3f Ks = d * toV(m->Ks);
  if (tex && m->mapNs.tex) Ns = Ns * texParamGet4f(m->mapNs, dg.st, dg.pixelFootprint).x;
  if (tex && m->mapKs.tex) {
    const vec4f t = texParamGet4f(m->mapKs, dg.st, dg.pixelFootprint);
    Ks = Ks * make_vec3f(t.x, t.y, t.z);
  }
  if (reduce_max(Ks) > 0.0f) {
    b.albedo = b.albedo + Ks;
    b.scatteringType |= SC_GLOSSY_REFLECTION;
    b.impS = luminance(Ks);
    b.importanceSum += b.impS;
    b.ns = Ns * 4.f; // Phong -> Blinn (Specular.ih:71-79)
    const float tmp = (b.ns + 2.f) * (1.f / (2.f - expf(b.ns * -OC_LNTWO_OVER_TWO)));
    b.ks = Ks * (OC_ONE_OVER_FOUR_PI * tmp);
    b.specWeight = Ks * (2.f * tmp * (1.f / (b.ns + 1.f)));
  }
}

__device__ __forceinline__ EvalRes Lambert_eval(const ObjBSDF &b, vec3f wi)
{
  EvalRes r;
  const float cosThetaI = fmaxf(dot(wi, b.frame.vz), 0.f);
  r.pdf = cosThetaI * OC_ONE_OVER_PI;
  r.value = b.Kd * OC_ONE_OVER_PI * cosThetaI;
  return r;
}
__device__ __forceinline__ EvalRes Specular_eval(const ObjBSDF &b, vec3f wo, vec3f wi)
{
  EvalRes r; r.value = make_vec3f(0.f); r.pdf = 0.f;
  const vec3f N = b.frame.vz;
  const float cosThetaO = dot(wo, N), cosThetaI = dot(wi, N);
  if (cosThetaO <= 0.0f || cosThetaI <= 0.0f) return r;
  const vec3f wh = normalize(wo + wi);
  const float cosThetaH = dot(wh, N);
  r.pdf = powerCosineSampleHemispherePDF(cosThetaH, b.ns) * (1.f / (4.0f * dot(wo, wh)));
  r.value = b.ks * powf(cosThetaH, b.ns);
  return r;
}
// MULTIBSDF_EVAL_* (bsdfs/MultiBSDF.ih:70-90); `skip` = lobe already sampled (-1: none)
__device__ __forceinline__ void OBJ_evalChildren(const ObjBSDF &b, vec3f wo, vec3f wi, int skip, EvalRes &e)
{
  if (b.impL > 0.f && skip != 0) {
    const EvalRes c = Lambert_eval(b, wi);
    e.value = e.value + c.value;
    e.pdf += c.pdf * b.impL;
  }
  if (b.impS > 0.f && skip != 2) {
    const EvalRes c = Specular_eval(b, wo, wi);
    e.value = e.value + c.value;
    e.pdf += c.pdf * b.impS;
  }
}
__device__ __forceinline__ EvalRes OBJ_eval(const ObjBSDF &b, vec3f wo, vec3f wi)
{
  EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
  if (b.thinGlass || b.pureTransmission) return e; // purely specular
  OBJ_evalChildren(b, wo, wi, -1, e);
  e.pdf *= 1.f / b.importanceSum;
  return e;
}
// MULTIBSDF_SAMPLE_* (bsdfs/MultiBSDF.ih:93-128)
// RobustThinDielectric_sample (bsdfs/RobustThinDielectric.ih:17-62)
__device__ __forceinline__ SampleRes RobustThinDielectric_sample(const ObjBSDF &b, vec3f wo, float ss)
{
  SampleRes res;
  const vec3f N = b.frame.vz;
  const float eta = b.ns;
  const float cosThetaO = fmaxf(dot(wo, N), 0.f);
  float cosThetaT;
  const float Fo = fresnelDielectricEx(cosThetaO, cosThetaT, eta);
  const float Ft = fresnelDielectric(cosThetaT, 1.f / eta);
  const float length = 1.f / cosThetaT;
  const vec3f A = expf3(b.Kd * length);
  const vec3f one = make_vec3f(1.f);
  const vec3f R = make_vec3f(Fo) + (1.f - Fo) * (1.f - Ft) * (Ft * sqr3(A)) * rcp3(one - sqr3(Ft * A));
  const vec3f T = (1.f - Fo) * A * (1.f - Ft) * (one + sqr3(Ft * A) * rcp3(one - sqr3(Ft * A)));
  const float pR = luminance(R), pT = luminance(T);
  const float p = fmaxf(0.25f, fminf(pR * (1.f / (pR + pT)), 0.75f));
  if (ss <= p) {
    res.wi = (2.0f * cosThetaO) * N - wo;
    res.type = SC_SPECULAR_REFLECTION;
    res.weight = R * (1.f / p);
  } else {
    res.wi = neg(wo);
    res.type = SC_SPECULAR_TRANSMISSION;
    res.weight = T * (1.f / (1.f - p));
  }
  res.pdf = INFINITY;
  return res;
}
__device__ __forceinline__ SampleRes OBJ_sample(const ObjBSDF &b, vec3f wo, vec2f s, float ss)
{
  if (b.thinGlass) return RobustThinDielectric_sample(b, wo, ss);
  if (b.pureTransmission) { // Transmission_sample (bsdfs/Transmission.ih:17-27)
    SampleRes r; r.wi = neg(wo); r.pdf = INFINITY; r.type = SC_SPECULAR_TRANSMISSION; r.weight = b.T;
    return r;
  }
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  if (b.importanceSum == 0.f) return res;
  const float x = ss * b.importanceSum;
  int choice = 0;
  float prefix = 0.f;
  float choiceImp = 0.f;
  if (x < (prefix += b.impL)) {
    choice = 0; choiceImp = b.impL;
    const vec3f localDir = cosineSampleHemisphere(s);
    res.wi = b.frame * localDir;
    res.pdf = localDir.z * OC_ONE_OVER_PI;
    res.type = SC_DIFFUSE_REFLECTION;
    res.weight = b.Kd;
  } else if (x < (prefix += b.impT)) {
    choice = 1; choiceImp = b.impT;
    res.wi = neg(wo);
    res.pdf = INFINITY;
    res.type = SC_SPECULAR_TRANSMISSION;
    res.weight = b.T;
  } else if (x < (prefix += b.impS)) {
    choice = 2; choiceImp = b.impS;
    const vec3f whl = powerCosineSampleHemisphere(b.ns, s);
    const vec3f wh = b.frame * whl;
    const float cosThetaOH = dot(wo, wh);
    res.wi = 2.f * cosThetaOH * wh - wo; // reflect(wo, wh, cosThetaOH)
    const float cosThetaI = dot(res.wi, b.frame.vz);
    if (cosThetaI <= 0.0f) {
      res.weight = make_vec3f(0.f); // below surface; pdf/type left as in the reference's zero result
    } else {
      res.pdf = powerCosineSampleHemispherePDF(whl.z, b.ns) * (1.f / (4.0f * cosThetaOH));
      res.type = SC_GLOSSY_REFLECTION;
      res.weight = b.specWeight * cosThetaOH;
    }
  }
  if ((res.weight.x == 0.f && res.weight.y == 0.f && res.weight.z == 0.f) || res.pdf == 0.0f) {
    res.weight = make_vec3f(0.f);
    return res;
  }
  if (res.type & SC_SPECULAR) {
    res.weight = res.weight * (b.importanceSum * (1.f / choiceImp));
    return res;
  }
  EvalRes e;
  e.value = res.weight * res.pdf;
  res.pdf *= choiceImp;
  e.pdf = res.pdf;
  OBJ_evalChildren(b, wo, res.wi, choice, e);
  res.pdf = e.pdf * (1.f / b.importanceSum);
  res.weight = e.value * (1.f / res.pdf);
  return res;
}
// ---- metal and plastic (extended-material kernels only) -------------------------------------------------------------
// Metal_getBSDF (materials/Metal.ispc:18-46, RGB form) and Plastic_getBSDF (materials/Plastic.ispc:25-67): fills `b` (frame,
// albedo, scattering type, kind) and the family state `x`; false: not one of these materials
// (out of line: the OBJ path of the extended-material kernels keeps its registers)
static __device__ __noinline__ void Ext_getBSDFImpl(ObjBSDF &b, ExtBSDF &x, const DifferentialGeometry &dg, const Medium &medium, vec3f rayDir);
__device__ __forceinline__ bool Ext_getBSDF(ObjBSDF &b, ExtBSDF &x, const DifferentialGeometry &dg, const Medium &medium, vec3f rayDir)
{
  const OcMaterial *m = dg.material;
  if (!m || m->type < OC_MATERIAL_METAL) return false;
  Ext_getBSDFImpl(b, x, dg, medium, rayDir);
  return true;
}
static __device__ __noinline__ void Ext_getBSDFImpl(ObjBSDF &b, ExtBSDF &x, const DifferentialGeometry &dg, const Medium &medium, vec3f rayDir)
{
  const OcMaterial *m = dg.material;
  b.valid = true;
  b.importanceSum = 1.f; b.impL = b.impT = b.impS = 0.f;
  b.Kd = b.T = b.ks = b.specWeight = make_vec3f(0.f); b.ns = 0.f;
  if (m->type == OC_MATERIAL_PRINCIPLED) {
    Principled_getBSDF(x.pr, m->principled, m->microfacetTables, dg, mediumIsOutside(medium, m->principled));
    b.kind = BSDF_KIND_PRINCIPLED;
    b.frame = x.pr.frame;
    b.frame.vz = x.pr.rootN; // (only the normal AOV reads it)
    b.albedo = x.pr.albedo;
    b.scatteringType = x.pr.scatteringType;
    return;
  }
  b.frame = frame(dg.Ns);
  if (m->type == OC_MATERIAL_CARPAINT) {
    CarPaint_getBSDF(x.car, b.frame.vz, b.albedo, b.scatteringType, *m->carPaint, m->microfacetTables, dg, rayDir);
    b.kind = BSDF_KIND_CARPAINT;
    return;
  }
  if (m->type == OC_MATERIAL_GLASS) { // Glass_getBSDF (materials/Glass.ispc:14-33): eta = the side the path is on over the other
    x.glassEta = mediumIsOutside(medium, m->principled) ? m->principled.outsideIor * (1.f / m->eta) : m->eta * (1.f / m->principled.outsideIor);
    x.glassNg = dg.Ng; x.glassNs = dg.Ns;
    b.kind = BSDF_KIND_GLASS;
    b.albedo = make_vec3f(1.0f);
    b.scatteringType = SC_SPECULAR;
    return;
  }
  if (m->type == OC_MATERIAL_VELVET) { // Velvet_getBSDF (materials/Velvet.ispc:22-60)
    VelvetBSDF &v = x.velvet;
    v.frame = b.frame;
    v.R0 = toV(m->Kd); v.b = m->Ns; v.R1 = toV(m->Ks); v.f = m->eta;
    v.imp0 = luminance(v.R0); v.imp1 = luminance(v.R1);
    v.importanceSum = 0.f; v.importanceSum += v.imp0; v.importanceSum += v.imp1;
    b.kind = BSDF_KIND_VELVET;
    b.albedo = make_vec3f(0.f) + v.R0 * 1.f + v.R1 * 1.f;
    b.scatteringType = SC_DIFFUSE_REFLECTION;
    return;
  }
  if (m->type == OC_MATERIAL_METALLICPAINT) { // MetallicPaint_getBSDF (materials/MetallicPaint.ispc:36-89)
    MetallicPaintSub &p = x.paint;
    vec3f color = toV(m->Kd) * make_vec3f(dg.color);
    if (m->mapKd.tex) { const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint); color = color * make_vec3f(t.x, t.y, t.z); }
    p.lambert.init(b.frame, color);
    p.imp0 = luminance(color); p.imp1 = 0.f;
    p.importanceSum = 0.f; p.importanceSum += p.imp0;
    p.albedo = make_vec3f(0.f) + color * 1.f;
    p.scatteringType = SC_DIFFUSE_REFLECTION;
    const float flakeAmount = m->attenuationScale;
    if (flakeAmount > 0.f) {
      ConductorBSDF &c = p.flakes;
      c.frame = b.frame; c.tables = m->microfacetTables; c.schlick = true;
      c.eta = toV(m->Ks); c.k = make_vec3f(flakeAmount);
      const vec3f Favg = fresnelSchlickAvg(c.eta, c.k);
      c.Eavg = mfAlbedoAvg(c.tables, m->Ns);
      c.fmsScale = sqr3(Favg) * c.Eavg / (make_vec3f(1.f) - Favg * (1.f - c.Eavg));
      c.mf = makeGGX(m->Ns, 0.f);
      c.roughness = m->Ns;
      p.imp1 = luminance(c.eta);
      p.importanceSum += p.imp1;
      p.albedo = p.albedo + (Favg * c.Eavg) * 1.f;
      p.scatteringType |= SC_GLOSSY_REFLECTION;
    }
    DielectricLayer_init(x.layer, p.albedo, p.scatteringType, m->eta, make_vec3f(1.0f), 1.0f, 1.0f);
    b.kind = BSDF_KIND_METALLICPAINT;
    b.albedo = x.layer.albedo; b.scatteringType = x.layer.scatteringType;
    return;
  }
  if (m->type == OC_MATERIAL_METAL || m->type == OC_MATERIAL_ALLOY) {
    // Metal_getBSDF (Metal.ispc:18-46, RGB form), Alloy_getBSDF (Alloy.ispc:16-51: Schlick Fresnel between color and edgeColor)
    const bool alloy = m->type == OC_MATERIAL_ALLOY;
    x.cond.frame = b.frame;
    x.cond.eta = toV(m->Kd); x.cond.k = toV(m->Ks);
    x.cond.schlick = alloy;
    vec3f Favg = make_vec3f(0.f); // Fresnel_dispatch_evalAvg of the RGB conductor is 0 (Fresnel.ih:307-320)
    if (alloy) {
      x.cond.eta = x.cond.eta * make_vec3f(dg.color);
      if (m->mapKd.tex) { const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint); x.cond.eta = x.cond.eta * make_vec3f(t.x, t.y, t.z); }
      if (m->mapKs.tex) { const vec4f t = texParamGet4f(m->mapKs, dg.st, dg.pixelFootprint); x.cond.k = x.cond.k * make_vec3f(t.x, t.y, t.z); }
      Favg = fresnelSchlickAvg(x.cond.eta, x.cond.k);
    }
    x.cond.tables = m->microfacetTables;
    if (m->Ns == 0.0f) {
      b.kind = BSDF_KIND_CONDUCTOR;
      b.albedo = Favg;
      b.scatteringType = SC_SPECULAR_REFLECTION;
    } else {
      float roughness = m->Ns;
      if (m->mapNs.tex) roughness = roughness * texParamGet4f(m->mapNs, dg.st, dg.pixelFootprint).x;
      b.kind = BSDF_KIND_MICROFACET_CONDUCTOR;
      x.cond.Eavg = mfAlbedoAvg(x.cond.tables, roughness);
      x.cond.fmsScale = sqr3(Favg) * x.cond.Eavg / (make_vec3f(1.f) - Favg * (1.f - x.cond.Eavg));
      x.cond.mf = makeGGX(roughness, 0.f);
      x.cond.roughness = roughness;
      b.albedo = Favg * x.cond.Eavg;
      b.scatteringType = SC_GLOSSY_REFLECTION;
    }
    return;
  }
  x.lambert.init(b.frame, toV(m->Kd));
  if (m->Ns == 0.0f) {
    b.kind = BSDF_KIND_PLASTIC_SMOOTH;
    DielectricLayer_init(x.layer, x.lambert.albedo, x.lambert.scatteringType, m->eta, make_vec3f(1.0f), 1.0f, 1.0f);
  } else {
    b.kind = BSDF_KIND_PLASTIC_ROUGH;
    MicrofacetDielectricLayer_init(x.layer, m->microfacetTables, x.lambert.albedo, x.lambert.scatteringType, m->eta, make_vec3f(1.0f), 1.0f,
        m->Ns, 0.f, 1.0f);
  }
  b.albedo = x.layer.albedo;
  b.scatteringType = x.layer.scatteringType;
}
static __device__ __noinline__ EvalRes Ext_eval(const ObjBSDF &b, const ExtBSDF &x, vec3f wo, vec3f wi)
{
  EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
  switch (b.kind) {
  case BSDF_KIND_MICROFACET_CONDUCTOR: return MicrofacetConductor_eval(x.cond, wo, wi);
  case BSDF_KIND_PLASTIC_SMOOTH: return DielectricLayer_eval(x.layer, b.frame.vz, x.lambert, wo, wi);
  case BSDF_KIND_PLASTIC_ROUGH: return MicrofacetDielectricLayer_eval(x.layer, b.frame, x.lambert, wo, wi);
  case BSDF_KIND_PRINCIPLED: return Principled_eval(x.pr, wo, wi);
  case BSDF_KIND_CARPAINT: return CarPaint_eval(x.car, wo, wi);
  case BSDF_KIND_VELVET: return x.velvet.eval(wo, wi);
  case BSDF_KIND_METALLICPAINT: return DielectricLayer_eval(x.layer, b.frame.vz, x.paint, wo, wi);
  default: return e; // smooth conductor: purely specular
  }
}
// RobustDielectric_sample (bsdfs/RobustDielectric.ih:27-176): reflection / refraction with the normals flipped to the incoming
// ray's side, grazing directions nudged off the surface, directions that end on the wrong side of the geometry mirrored back
__device__ __forceinline__ SampleRes RobustDielectric_sample(const ExtBSDF &x, vec3f wo, float randomF)
{
  SampleRes res;
  res.pdf = INFINITY;
  const vec3f incomingDir = neg(wo);
  vec3f shadingNormal = x.glassNs, geometricNormal = x.glassNg;
  const float eta = 1.f / x.glassEta;
  const float glass_epsilon = 1e-6f;
  const bool canReflect = (eta < 1.f - glass_epsilon || eta > 1.f + glass_epsilon);
  if (dot(incomingDir, shadingNormal) > 0.f) shadingNormal = neg(shadingNormal);
  if (dot(incomingDir, geometricNormal) > 0.f) geometricNormal = neg(geometricNormal);
  // ReflectionDirection (:79-88)
  const vec3f par = dot(incomingDir, shadingNormal) * shadingNormal;
  const vec3f dirR = (incomingDir - par) - par;
  // RefractionDirection (:34-77); the head-on test of the reference (1 - 1e-6 < |cos| && |cos| < 1 - 1e-6) is never true
  vec3f dirT;
  float FresnelR;
  bool mustReflect = false;
  {
    const float cos1 = dot(incomingDir, shadingNormal);
    const float cos1_abs = cos1 < 0.f ? -cos1 : cos1;
    const vec3f dirPar = cos1 * shadingNormal;
    const vec3f dirPerp = incomingDir - dirPar;
    const vec3f bb = normalize(dirPerp);
    const float sin1 = dot(dirPerp, bb);
    const float sin2 = sin1 / eta;
    if (sin2 > 1.0f) { dirT = make_vec3f(0.f); FresnelR = 1.f; mustReflect = true; }
    else {
      const float cos2 = sqrtf(clampf(1.f - sin2 * sin2, 0.f, 1.f));
      const vec3f N2 = normalize(dirPar);
      dirT = N2 * cos2 + bb * sin2;
      const float c1 = cos1_abs, c2 = cos2;
      const float r1 = (eta * c1 - c2) / (eta * c1 + c2);
      const float r2 = (c1 - eta * c2) / (c1 + eta * c2);
      FresnelR = 0.5f * (r1 * r1 + r2 * r2);
    }
    FresnelR = clampf(FresnelR, 0.f, 1.f);
  }
  res.weight = make_vec3f(1.f);
  if ((randomF < FresnelR && canReflect) || mustReflect) {
    res.wi = dirR;
    res.type = SC_SPECULAR_REFLECTION;
    const float geometricCosine = dot(res.wi, geometricNormal);
    if (geometricCosine >= 0.f) {
      if (geometricCosine < glass_epsilon) res.wi = normalize(res.wi + glass_epsilon * geometricNormal);
    } else
      res.wi = res.wi - 2.f * geometricNormal * geometricCosine;
  } else {
    res.wi = dirT;
    res.type = SC_SPECULAR_TRANSMISSION;
    const float geometricCosine = dot(res.wi, geometricNormal);
    if (geometricCosine <= 0.f) {
      if (geometricCosine > -glass_epsilon) res.wi = normalize(res.wi - glass_epsilon * geometricNormal);
    } else
      res.type = SC_SPECULAR_REFLECTION;
  }
  return res;
}
static __device__ __noinline__ SampleRes Ext_sample(const ObjBSDF &b, const ExtBSDF &x, vec3f wo, vec2f s, float ss)
{
  switch (b.kind) {
  case BSDF_KIND_GLASS: return RobustDielectric_sample(x, wo, ss);
  case BSDF_KIND_CONDUCTOR: return Conductor_sample(x.cond, wo);
  case BSDF_KIND_MICROFACET_CONDUCTOR: return MicrofacetConductor_sample(x.cond, wo, s);
  case BSDF_KIND_PLASTIC_SMOOTH: return DielectricLayer_sample(x.layer, b.frame.vz, x.lambert, wo, s, ss);
  case BSDF_KIND_PRINCIPLED: return Principled_sample(x.pr, wo, s, ss);
  case BSDF_KIND_CARPAINT: return CarPaint_sample(x.car, wo, s, ss);
  case BSDF_KIND_VELVET: return x.velvet.sample(wo, s, ss);
  case BSDF_KIND_METALLICPAINT: return DielectricLayer_sample(x.layer, b.frame.vz, x.paint, wo, s, ss);
  default: return MicrofacetDielectricLayer_sample(x.layer, b.frame, x.lambert, wo, s, ss);
  }
}

// materials/OBJ.ispc:93-111, materials/ThinGlass.ispc:52-64
// Principled_selectNextMedium (Principled.ispc:582-608); the other materials built here have no interface (Material.ispc)
__device__ __forceinline__ void materialSelectNextMedium(const DifferentialGeometry &dg, Medium &medium)
{
  const OcMaterial *m = dg.material;
  if (m && m->type == OC_MATERIAL_GLASS) { // Glass_selectNextMedium (Glass.ispc:53-63)
    if (mediumIsOutside(medium, m->principled)) { medium.ior = m->eta; medium.attenuation = toV(m->attenuation); }
    else { medium.ior = m->principled.outsideIor; medium.attenuation = toV(m->principled.outsideAttenuation); }
    return;
  }
  if (!m || m->type != OC_MATERIAL_PRINCIPLED) return;
  const OcPrincipled &q = m->principled;
  if (q.thin || q.transmission <= PR_EPS) return;
  if (mediumIsOutside(medium, q)) {
    float ior = q.ior;
    vec3f tc = toV(q.transmissionColor);
    float depth = q.transmissionDepth;
    if (const OcPrincipledMaps *mp = q.maps) { // ior, transmissionColor and transmissionDepth maps (:594-601)
      if (mp->map[OC_PRMAP_IOR].tex) ior *= texParamGet4f(mp->map[OC_PRMAP_IOR], dg.st, dg.pixelFootprint).x;
      if (mp->map[OC_PRMAP_TRANSMISSIONCOLOR].tex) {
        const vec4f t = texParamGet4f(mp->map[OC_PRMAP_TRANSMISSIONCOLOR], dg.st, dg.pixelFootprint);
        tc = tc * make_vec3f(t.x, t.y, t.z);
      }
      if (mp->map[OC_PRMAP_TRANSMISSIONDEPTH].tex) depth *= texParamGet4f(mp->map[OC_PRMAP_TRANSMISSIONDEPTH], dg.st, dg.pixelFootprint).x;
    }
    if (ior < 1.f) ior = 1.f / ior;
    tc = clamp01(tc);
    depth = fmaxf(depth, PR_EPS);
    medium.ior = ior;
    medium.attenuation = logf3(tc) / depth;
  } else {
    medium.ior = q.outsideIor;
    medium.attenuation = toV(q.outsideAttenuation);
  }
}

template <bool TEX = false>
__device__ __forceinline__ vec3f materialTransparency(const DifferentialGeometry &dg, vec3f rayDir, const Medium &medium = mediumVacuum())
{
  const OcMaterial *m = dg.material;
  if (TEX && m && m->type == OC_MATERIAL_PRINCIPLED)
    return Principled_getTransparency(m->principled, m->microfacetTables, dg, rayDir, mediumIsOutside(medium, m->principled));
  if (TEX && m && m->type == OC_MATERIAL_GLASS) { // Glass_getTransparency (Glass.ispc:35-51)
    const float eta = mediumIsOutside(medium, m->principled) ? m->principled.outsideIor * (1.f / m->eta) : m->eta * (1.f / m->principled.outsideIor);
    const float cosThetaO = fmaxf(-dot(rayDir, dg.Ns), 0.0f);
    return make_vec3f(1.0f - fresnelDielectric(cosThetaO, eta));
  }
  if (m && m->type == OC_MATERIAL_THINGLASS) {
    const float cosThetaO = fmaxf(-dot(rayDir, dg.Ns), 0.0f);
    return RobustThinDielectric_getTransparency(cosThetaO, m->eta, ThinGlass_getAttenuation<TEX>(m, dg));
  }
  if (TEX && m && m->type == OC_MATERIAL_LUMINOUS) return make_vec3f(m->transparency); // Luminous_getTransparency
  if (!m || m->type != OC_MATERIAL_OBJ) return make_vec3f(0.f);
  float d = m->d;
  if (TEX && m->mapD.tex) d = d * texParamGet4f(m->mapD, dg.st, dg.pixelFootprint).x;
  d = d * dg.color.w;
  if (TEX && m->mapKd.tex && m->mapKd.tex->hasAlpha) d *= texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint).w;
  return toV(m->Tf) * d + make_vec3f(1.f - d);
}

// Material_dispatch_getEmission (render/MaterialDispatch.ispc; materials/Luminous.ispc:62-70): radiance * map_color
__device__ __forceinline__ vec3f materialEmission(const DifferentialGeometry &dg)
{
  const OcMaterial *m = dg.material;
  if (!m || !m->isEmissive) return make_vec3f(0.f);
  vec3f e = toV(m->emission);
  if (m->type == OC_MATERIAL_PRINCIPLED) { // Principled_getEmission (Principled.ispc:1055-1062)
    if (m->principled.maps && m->principled.maps->map[OC_PRMAP_EMISSION].tex) {
      const vec4f t = texParamGet4f(m->principled.maps->map[OC_PRMAP_EMISSION], dg.st, dg.pixelFootprint);
      e = e * make_vec3f(t.x, t.y, t.z);
    }
    return e;
  }
  if (m->mapKd.tex) {
    const vec4f t = texParamGet4f(m->mapKd, dg.st, dg.pixelFootprint);
    e = e * make_vec3f(t.x, t.y, t.z);
  }
  return e;
}

// ---- lights -------------------------------------------------------------------------------
struct LightSample { vec3f weight, dir; float dist, pdf; };
struct LightEval { vec3f radiance; float pdf; };

struct SphericalQuad
{
  vec3f e0, e1, z;
  float S, k, n0z, n0z2, n2z, x0, x1, y0, y1, z0;
};
// a / s, bit for bit, for vectors with many ZERO components (the edges of an axis-aligned light): an IEEE division whose numerator
// is 0 leaves the division's fast path (exponent 0) for a ~50-instruction subroutine -- on C1 that was 11 % of the shade step's
// instructions, all from the two edge normalisations below (ncu source page, profiles/r2_c1_trace_shade_ncu.txt). For a normal
// (finite, non-zero, non-denormal) divisor 0 / s == 0 * s exactly, sign included.
__device__ __forceinline__ float divZeroAware(float a, float s, bool sNormal) { return (a == 0.f && sNormal) ? a * s : a / s; }
__device__ __forceinline__ vec3f divZeroAware(vec3f a, float s)
{
  const bool sNormal = ((__float_as_uint(s) >> 23) & 0xffu) - 1u < 254u;
  return make_vec3f(divZeroAware(a.x, s, sNormal), divZeroAware(a.y, s, sNormal), divZeroAware(a.z, s, sNormal));
}
__device__ __forceinline__ SphericalQuad SphericalQuad_create(vec3f q0, vec3f e0, vec3f e1, vec3f n, vec3f P)
{
  SphericalQuad q;
  const float e0l = length(e0), e1l = length(e1);
  q.e0 = divZeroAware(e0, e0l); q.e1 = divZeroAware(e1, e1l); q.z = n;
  const vec3f d = q0 - P;
  q.z0 = dot(d, q.z);
  q.z = q.z0 > 0.f ? neg(q.z) : q.z;
  q.z0 = -fabsf(q.z0);
  q.x0 = dot(d, q.e0); q.y0 = dot(d, q.e1);
  q.x1 = q.x0 + e0l; q.y1 = q.y0 + e1l;
  const vec3f v00 = make_vec3f(q.x0, q.y0, q.z0), v01 = make_vec3f(q.x0, q.y1, q.z0),
              v10 = make_vec3f(q.x1, q.y0, q.z0), v11 = make_vec3f(q.x1, q.y1, q.z0);
  const vec3f n0 = normalize(cross(v00, v10));
  q.n0z = n0.z; q.n0z2 = n0.z * n0.z;
  const vec3f n1 = normalize(cross(v10, v11)), n2 = normalize(cross(v11, v01)), n3 = normalize(cross(v01, v00));
  q.n2z = n2.z;
  const float g0 = acosf(clampf(-dot(n0, n1), -1.f, 1.f)), g1 = acosf(clampf(-dot(n1, n2), -1.f, 1.f)),
              g2 = acosf(clampf(-dot(n2, n3), -1.f, 1.f)), g3 = acosf(clampf(-dot(n3, n0), -1.f, 1.f));
  q.k = OC_TWO_PI - g2 - g3;
  q.S = fmaxf(0.f, g0 + g1 - q.k);
  return q;
}
__device__ __forceinline__ float signf(float x) { return (float)(0.f < x) - (float)(x < 0.f); }
__device__ __forceinline__ vec3f sampleSphericalQuad(const SphericalQuad &q, vec2f s)
{
  const float au = s.x * q.S + q.k;
  float sinau, cosau;
  sincosf(au, &sinau, &cosau);
  const float fu = (cosau * q.n0z - q.n2z) * (1.f / sinau);
  const float cu = clampf(rsqrtf(fu * fu + sqr(q.n0z)) * signf(fu), -1.f, 1.f);
  const float x0 = -(cu * q.z0) * rsqrtf(1.f - cu * cu);
  const float xu = clampf(x0, q.x0, q.x1);
  const float ds = sqrtf(sqr(xu) + sqr(q.z0));
  const float ds2 = sqr(ds);
  const float h0 = q.y0 * rsqrtf(ds2 + sqr(q.y0));
  const float h1 = q.y1 * rsqrtf(ds2 + sqr(q.y1));
  const float hv = s.y * (h1 - h0) + h0;
  const float hv2 = sqr(hv);
  const float yv = (hv2 < 1.f - 1e-4f) ? (hv * ds) * rsqrtf(1.f - hv2) : q.y1;
  return xu * q.e0 + yv * q.e1 + q.z0 * q.z;
}
__device__ __forceinline__ float solve_quadratic(float A, float B, float C)
{
  if (A == 0.f) return -C / B;
  float rad = B * B - 4.f * A * C;
  if (rad < 0.f) return 0.f;
  rad = sqrtf(rad);
  float root;
  if (B >= 0.f) {
    root = (-B - rad) / (2.f * A);
    if (root < 0.f || root > 1.f) root = 2.f * C / (-B - rad);
  } else {
    root = 2.f * C / (-B + rad);
    if (root < 0.f || root > 1.f) root = (-B + rad) / (2.f * A);
  }
  return root;
}
__device__ __forceinline__ vec3f computeCosineWeightedRNG(vec3f q0, vec3f e0, vec3f e1, vec3f P, vec3f Ng, vec2f s)
{
  vec3f wp[4];
  float prob[4];
  wp[0] = q0; wp[1] = q0 + e0; wp[2] = q0 + e1; wp[3] = wp[1] + e1;
#pragma unroll
  for (int i = 0; i < 4; i++)
    prob[i] = fmaxf(0.f, dot(normalize(wp[i] - P), Ng));
  if ((prob[0] + prob[2]) == 0.f)
    return make_vec3f(s.x, s.y, 1.f);
  const float totProb = prob[0] + prob[2] + prob[1] + prob[3];
  const float W00 = prob[0], W01 = prob[2], W10 = prob[1], W11 = prob[3];
  float u;
  {
    const float A = W10 + W11 - W01 - W00, B = 2.f * (W01 + W00);
    if (A + B == 0.f) u = s.x;
    else u = solve_quadratic(A, B, -(W00 + W01 + W10 + W11) * s.x);
  }
  float v;
  {
    const float A = (1.f - u) * W01 + u * W11 - (1.f - u) * W00 - u * W10;
    const float B = 2.f * (1.f - u) * W00 + 2.f * u * W10;
    if (A + B == 0.f) v = s.y;
    else v = solve_quadratic(A, B, -((1.f - u) * W01 + u * W11 + (1.f - u) * W00 + u * W10) * s.y);
  }
  const float pdf = 4.f * ((1.f - u) * (1.f - v) * prob[0] + u * (1.f - v) * prob[1] + u * v * prob[3] + (1.f - u) * v * prob[2])
      / totProb;
  return make_vec3f(u, v, pdf);
}

__device__ __forceinline__ LightSample QuadLight_sampleArea(const OcLight &l, vec3f P, vec2f s)
{
  LightSample r;
  const vec3f p = toV(l.position) + toV(l.edge1) * s.x + toV(l.edge2) * s.y;
  const vec3f dir = p - P;
  const float dist = length(dir);
  r.dir = dir / dist;
  r.dist = dist;
  const float cosd = dot(toV(l.nnormal), r.dir);
  r.pdf = l.ppdf * sqr(dist) / fabsf(cosd);
  const float weight = cosd > 0.f ? 1.f / r.pdf : 0.f;
  r.weight = toV(l.radiance) * weight;
  return r;
}
__device__ __forceinline__ LightSample QuadLight_sample(const OcLight &l, const DifferentialGeometry &dg, vec2f s)
{
  if (s.x == 0.0f && s.y == 0.0f)
    return QuadLight_sampleArea(l, dg.P, make_vec2f(0.5f, 0.5f));
  LightSample r;
  const vec3f pos = toV(l.position), e1 = toV(l.edge1), e2 = toV(l.edge2), nn = toV(l.nnormal);
  const SphericalQuad quad = SphericalQuad_create(pos, e1, e2, neg(nn), dg.P);
  const vec3f cosW = computeCosineWeightedRNG(pos, e1, e2, dg.P, dg.Ng, s);
  const vec3f dir = sampleSphericalQuad(quad, make_vec2f(cosW.x, cosW.y));
  const float dist = length(dir);
  r.dir = dir / dist;
  r.dist = dist;
  const float cosd = dot(nn, r.dir);
  r.pdf = quad.S == 0.f ? 0.f : cosW.z / quad.S;
  r.weight = (cosd > 0.f) && (r.pdf != 0.f) ? toV(l.radiance) * (1.f / r.pdf) : make_vec3f(0.f);
  return r;
}
__device__ __forceinline__ LightEval QuadLight_eval(const OcLight &l, vec3f P, vec3f dir, float minDist, float maxDist)
{
  LightEval r; r.radiance = make_vec3f(0.f); r.pdf = 0.f;
  const vec3f nn = toV(l.nnormal), pos = toV(l.position), e1 = toV(l.edge1), e2 = toV(l.edge2);
  const float cosd = dot(nn, dir);
  if (cosd <= 0.f) return r;
  const vec3f c = pos - P;
  const vec3f rr = cross(c, dir);
  const float rcosd = 1.f / cosd;
  const float u = dot(rr, e1) * rcosd;
  const float v = -dot(rr, e2) * rcosd;
  if (fminf(u, v) < 0.f || fmaxf(u, v) * l.ppdf > 1.0f) return r;
  const float dist = dot(nn, c) * rcosd;
  if (dist <= minDist || dist > maxDist) return r;
  const SphericalQuad quad = SphericalQuad_create(pos, e1, e2, neg(nn), P);
  if (quad.S > 0.f) {
    r.radiance = toV(l.radiance);
    r.pdf = 1.f / quad.S;
  }
  return r;
}

// PointLight ("sphere" light): lights/PointLight.ispc:38-92 (sample), :131-160 (eval); no intensity distribution
__device__ __forceinline__ LightSample PointLight_sample(const OcLight &l, const DifferentialGeometry &dg, vec2f s)
{
  LightSample res;
  const vec3f dir = toV(l.position) - dg.P;
  const float dist2 = dot(dir, dir);
  const float invdist = 1.f / sqrtf(dist2);
  res.dir = dir * invdist;
  res.dist = dist2 * invdist;
  res.pdf = INFINITY;
  const float sinTheta = l.radius * invdist;
  if ((l.radius > 0.f) & (sinTheta > 0.005f)) {
    if (sinTheta < 1.f) {
      const float cosTheta = sqrtf(1.f - sinTheta * sinTheta);
      const vec3f localDir = uniformSampleCone(cosTheta, s);
      res.dir = frame(res.dir) * localDir;
      res.pdf = 1.f / (OC_TWO_PI * (1.0f - cosTheta));
      const float c = localDir.z;
      const float disc = l.radius * l.radius - (1.f - c * c) * dist2;
      res.dist = c * res.dist - sqrtf(disc > 0.f ? disc : 0.f);
      res.weight = toV(l.radiance) / res.pdf;
    } else
      res.weight = make_vec3f(0.f);
  } else
    res.weight = toV(l.intensity) * (invdist * invdist);
  return res;
}
__device__ __forceinline__ LightEval PointLight_eval(const OcLight &l, vec3f P, vec3f dir, float minDist, float maxDist)
{
  LightEval res; res.radiance = make_vec3f(0.f); res.pdf = 0.f;
  const vec3f A = toV(l.position) - P;
  const float centerDist2 = dot(A, A);
  const float sinTheta2 = (l.radius * l.radius) * (1.f / centerDist2);
  if (l.radius > 0.f && sqrtf(sinTheta2) > 0.005f) {
    // intersectSphere (common/Intersect.ih:29-78), entry hit only
    const float rd2 = 1.0f / dot(dir, dir);
    const float projCO = dot(A, dir) * rd2;
    const vec3f perp = A - projCO * dir;
    const float l2 = dot(perp, perp), r2 = l.radius * l.radius;
    if (!(l2 > r2)) {
      const float t = projCO - sqrtf((r2 - l2) * rd2);
      if (t > minDist && t < maxDist) {
        const float cosTheta = sqrtf(1.f - sinTheta2);
        res.pdf = 1.f / (OC_TWO_PI * (1.0f - cosTheta));
        res.radiance = toV(l.radiance);
      }
    }
  }
  return res;
}

// SpotLight: lights/SpotLight.ispc:26-86 (AngularAttenuation, Sample), :124-164 (Eval); no intensity distribution
__device__ __forceinline__ float SpotLight_attenuation(const OcLight &l, float cosAngle)
{
  return fmaxf(0.f, fminf((cosAngle - l.cosAngle) * l.cosAngleScale, 1.f));
}
__device__ __forceinline__ LightSample SpotLight_sample(const OcLight &l, const DifferentialGeometry &dg, vec2f s)
{
  LightSample res;
  const vec3f direction = toV(l.fz);
  res.dir = toV(l.position) - dg.P;
  if (l.radius > 0.0f) {
    linear3f l2w; l2w.vx = toV(l.fx); l2w.vy = toV(l.fy); l2w.vz = direction;
    const float r = sqrtf(lerpf(s.x, l.innerRadius * l.innerRadius, l.radius * l.radius)); // uniformSampleRing
    float sinPhi, cosPhi;
    sincosf(OC_TWO_PI * s.y, &sinPhi, &cosPhi);
    res.dir = l2w * make_vec3f(r * cosPhi, r * sinPhi, 0.f) + res.dir;
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
    res.weight = toV(l.radiance) * weight;
  } else {
    res.pdf = INFINITY;
    weight *= fabsf(cosAngle);
    res.weight = toV(l.intensity) * ((invdist * invdist) * weight);
  }
  return res;
}
__device__ __forceinline__ LightEval SpotLight_eval(const OcLight &l, vec3f P, vec3f dir, float minDist, float maxDist)
{
  LightEval res; res.radiance = make_vec3f(0.f); res.pdf = 0.f;
  if (l.radius > 0.f) {
    const vec3f direction = toV(l.fz);
    const float cosAngle = -dot(direction, dir);
    if (cosAngle > l.cosAngle) {
      const vec3f vp = P - toV(l.position);
      const float t = dot(vp, direction) * (1.f / cosAngle);
      if ((t > minDist) & (t <= maxDist)) {
        const vec3f vd = vp + t * dir;
        const float d2 = dot(vd, vd);
        if (d2 < l.radius * l.radius && d2 > l.innerRadius * l.innerRadius) {
          res.radiance = toV(l.radiance) * SpotLight_attenuation(l, cosAngle);
          res.pdf = l.ppdf * ((t * t) / fabsf(cosAngle));
        }
      }
    }
  }
  return res;
}

struct Sample1D { int idx; float frac, pdf, prob; };
__device__ __forceinline__ Sample1D Distribution1D_sample(int size, const float *cdf, int start, float s);
struct SampleAreaRes { vec3f pos, normal; vec2f st; float epsilon; };
// Mesh_sampleArea (geometry/Mesh.ispc:270-383): uniform point on a triangle / quad (Embree's split of the quad)
__device__ __forceinline__ SampleAreaRes Mesh_sampleArea(const OcGeometry *g, uint32_t primID, const affine3f &xfm, vec2f s)
{
  SampleAreaRes res;
  uint32_t ix, iy, iz, iw;
  bool quad = false;
  if (g->type == OC_GEOM_TRIANGLES) {
    ix = get1ui(g->index, primID, 0); iy = get1ui(g->index, primID, 1); iz = get1ui(g->index, primID, 2); iw = iz;
  } else {
    ix = get1ui(g->index, primID, 0); iy = get1ui(g->index, primID, 1); iz = get1ui(g->index, primID, 2); iw = get1ui(g->index, primID, 3);
    quad = iz != iw;
  }
  const vec3f a = get3f(g->vertex, ix), b = get3f(g->vertex, iy), d = get3f(g->vertex, iw);
  const vec3f c = quad ? get3f(g->vertex, iz) : d;
  res.epsilon = reduce_max(vmax(vmax(vabs(a), vabs(b)), vmax(vabs(c), vabs(d)))) * OC_ULP_EPSILON;
  res.epsilon *= fmaxf(fabsf(xfm.l.vx.x), fmaxf(fabsf(xfm.l.vy.y), fabsf(xfm.l.vz.z)));
  const vec3f e1 = xfmVector(xfm, a - d), e2 = xfmVector(xfm, b - d);
  const vec3f m1 = cross(e1, e2);
  res.normal = m1;
  vec3f v0 = a, v1 = b, v2 = d;
  vec2f sp = s;
  bool invertUV = false;
  if (quad) {
    const vec3f e3 = xfmVector(xfm, c - d);
    const vec3f m2 = cross(e2, e3);
    const float a1 = length(m1), a2 = length(m2);
    const float p1 = a1 * (1.f / (a1 + a2));
    if (s.x < p1) sp.x *= 1.f / p1;
    else {
      sp.x = (s.x - p1) * (1.f / (1.f - p1));
      invertUV = true;
      v0 = c; v1 = d; v2 = b;
      res.normal = m2;
    }
  }
  const float su = sqrtf(sp.x);
  const vec2f uv = make_vec2f(sp.y * su, (1.0f - sp.y) * su); // uniformSampleTriangleUV
  const vec3f localPos = v0 + uv.x * (v1 - v0) + uv.y * (v2 - v0);
  res.normal = normalize(res.normal);
  res.pos = xfmPoint(xfm, localPos);
  res.st = invertUV ? make_vec2f(1.0f - uv.x, 1.0f - uv.y) : uv;
  if (g->texcoord.addr) {
    const float u = res.st.x, v = res.st.y;
    const vec4f uv4 = quad ? make_vec4f((1 - v) * (1 - u), (1 - v) * u, v * u, v * (1 - u)) : make_vec4f(1.f - u - v, u, v, 0.f);
    const uint32_t fo = primID * (g->type == OC_GEOM_TRIANGLES ? 3u : 4u);
    const bool fv = g->flags & OC_GEOMFLAG_TEXCOORD_FACEVARYING;
    const vec2f ta = get2f(g->texcoord, fv ? fo : ix), tb = get2f(g->texcoord, fv ? fo + 1 : iy), tc = get2f(g->texcoord, fv ? fo + 2 : iz);
    if (g->type == OC_GEOM_TRIANGLES) res.st = uv4.x * ta + uv4.y * tb + uv4.z * tc;
    else {
      const vec2f td = get2f(g->texcoord, fv ? fo + 3 : iw);
      res.st = uv4.x * ta + uv4.y * tb + uv4.z * tc + uv4.w * td;
    }
  }
  return res;
}
// Spheres_sampleArea (geometry/Spheres.ispc:44-74), sphere type OSP_SPHERE
__device__ __forceinline__ SampleAreaRes Spheres_sampleArea(const OcGeometry *g, uint32_t primID, const affine3f &xfm, const affine3f &rcp_xfm, vec2f s)
{
  const vec4f vertex = get4f(g->sphere, primID);
  const vec3f localNormal = uniformSampleSphere(vertex.w, s);
  SampleAreaRes res;
  res.pos = xfmPoint(xfm, make_vec3f(vertex) + localNormal);
  res.normal = normalize(xfmVector(transposed(rcp_xfm.l), localNormal));
  res.st = g->texcoord.addr ? get2f(g->texcoord, primID) : make_vec2f(0.f, 0.f);
  res.epsilon = vertex.w * OC_ULP_EPSILON;
  res.epsilon *= fmaxf(fabsf(xfm.l.vx.x), fmaxf(fabsf(xfm.l.vy.y), fabsf(xfm.l.vz.z)));
  return res;
}
struct LightSample;
// GeometryLight_sample (render/pathtracer/GeometryLight.ispc:62-128): a primitive by area, a point on it, its emission
static __device__ __noinline__ LightSample GeometryLight_sample(const OcWorld *world, const OcRenderer &ren, const OcLight &l,
    const DifferentialGeometry &dg, vec2f s);

// CylinderLight (lights/CylinderLight.ispc:30-231): the rectangle that bounds the cylinder as seen from P is sampled by solid
// angle (cosine-weighted, like the quad light), the sample counts if its direction meets the cylinder's body
// (intersectCylinder, common/Intersect.ih:80-131: entry.hit only for body hits). Like the hdri light it is known to the
// kernels of the extended instantiation only.
__device__ __forceinline__ bool cylinderBodyEntry(vec3f rayOrg, vec3f rayDir, vec3f v0, vec3f v1, float radius, float &t)
{
  const vec3f cZ = v1 - v0, q = rayOrg - v0;
  const float z2 = dot(cZ, cZ), d = dot(cZ, rayDir), c = dot(cZ, q);
  const float A = z2 - sqr(d);
  const float B = z2 * dot(q, rayDir) - c * d;
  const float C = z2 * dot(q, q) - sqr(c) - sqr(radius) * z2;
  float radical = B * B - A * C;
  if (radical < 0.f) return false;
  radical = sqrtf(radical);
  const float tin = (-B - radical) / A;
  const float yin = c + tin * d;
  t = tin;
  return yin > 0.f && yin < z2;
}
__device__ __forceinline__ void cylinderComputeRectBounds(vec3f position0, vec3f position1, float radius, vec3f P, vec3f &q0, vec3f &e0,
    vec3f &e1, float &S)
{
  vec3f O = position0;
  vec3f Z = position1 - position0;
  const float H = length(Z);
  Z = Z / H;
  vec3f Dp = P - O;
  float h0 = -dot(Dp, Z);
  if (h0 < -H) {
    O = position1;
    Z = -1.f * Z;
    Dp = P - O;
    h0 = -dot(Dp, Z);
  }
  const vec3f D = -h0 * Z - Dp;
  const float d = length(D);
  const float srad = fminf(radius / d, 1.f);
  const float q = fmaxf(1.f - sqr(srad), 4.e-6f);
  const float b = d * q;
  const float c = radius * sqrtf(q);
  const float b2c2 = sqr(b) + sqr(c);
  const vec3f X = normalize(cross(D, Z));
  const float h1 = (1.f + srad) * (h0 + H);
  const float hO = h0 < 0.f ? (1.f + srad) * h0 : h0;
  q0 = P + D * q - c * X + hO * Z;
  e0 = 2.f * c * X;
  e1 = (h1 - hO) * Z;
  S = 2.f * atanf(c * h1 / (b * sqrtf(b2c2 + sqr(h1)))) - 2.f * atanf(c * hO / (b * sqrtf(b2c2 + sqr(hO)))); // cylinderAxialSolidAngle
}
static __device__ __noinline__ LightSample CylinderLight_sample(const OcLight &l, const DifferentialGeometry &dg, vec2f sp)
{
  LightSample res; res.weight = make_vec3f(0.f); res.pdf = 0.f; res.dir = make_vec3f(0.f); res.dist = 0.f;
  const vec3f p0 = toV(l.position), p1 = toV(l.edge1);
  float S = 0.f, pdf = 0.f;
  vec3f dir, q0, e0, e1;
  cylinderComputeRectBounds(p0, p1, l.radius, dg.P, q0, e0, e1, S);
  if (sp.x == 0.0f && sp.y == 0.0f) { // scivis: towards the centre
    dir = 0.5f * (p0 + p1) - dg.P;
    if (S > 0.f) pdf = 1.f;
  } else {
    const vec3f n = normalize(cross(e0, e1));
    const SphericalQuad quad = SphericalQuad_create(q0, e0, e1, n, dg.P);
    const vec3f cosW = computeCosineWeightedRNG(q0, e0, e1, dg.P, dg.Ng, sp);
    dir = sampleSphericalQuad(quad, make_vec2f(cosW.x, cosW.y));
    pdf = cosW.z;
  }
  if (S <= 0.f || pdf == 0.f) return res;
  res.dir = normalize(dir);
  float t;
  if (cylinderBodyEntry(dg.P, res.dir, p0, p1, l.radius, t)) {
    res.pdf = pdf / S;
    res.weight = toV(l.radiance) * (1.f / res.pdf);
    res.dist = t;
  }
  return res;
}
static __device__ __noinline__ LightEval CylinderLight_eval(const OcLight &l, vec3f P, vec3f dir, float minDist, float maxDist)
{
  LightEval res; res.radiance = make_vec3f(0.f); res.pdf = 0.f;
  const vec3f p0 = toV(l.position), p1 = toV(l.edge1);
  float t;
  if (cylinderBodyEntry(P, dir, p0, p1, l.radius, t) && minDist < t && t <= maxDist) {
    vec3f q0, e0, e1;
    float S;
    cylinderComputeRectBounds(p0, p1, l.radius, P, q0, e0, e1, S);
    if (S > 0.f) { res.radiance = toV(l.radiance); res.pdf = 1.f / S; }
  }
  return res;
}

// HDRILight_sample / _eval (lights/HDRILight.ispc:14-46, 82-110) with Distribution2D_sample / _pdf (math/Distribution2D.ispc:9-43);
// only the kernels with texture support know this light (a frame with one runs them)
static __device__ __noinline__ LightSample HDRILight_sample(const OcLight &l, vec2f s)
{
  LightSample res;
  const OcTexture2D &map = *l.hdriMap;
  const vec2f rcpSize = make_vec2f(1.f / map.sizeX, 1.f / map.sizeY);
  const Sample1D sy = Distribution1D_sample(map.sizeY, l.hdriCdfY, 0, s.y);
  const Sample1D sx = Distribution1D_sample(map.sizeX, l.hdriCdfX, sy.idx * map.sizeX, s.x);
  vec2f uv = make_vec2f((sx.idx + sx.frac) * rcpSize.x, (sy.idx + sy.frac) * rcpSize.y);
  const float pdf2d = sx.pdf * sy.pdf;
  const vec2f halfTexel = make_vec2f((1.f / (float)map.sizeX) * 0.5f, (1.f / (float)map.sizeY) * 0.5f);
  uv = make_vec2f(uv.x - halfTexel.x, uv.y - halfTexel.y);
  const float phi = OC_TWO_PI * uv.x, theta = OC_PI * uv.y;
  float sinTheta, cosTheta;
  sincosf(theta, &sinTheta, &cosTheta);
  const vec3f localDir = cartesian(phi, sinTheta, cosTheta);
  linear3f l2w; l2w.vx = toV(l.fx); l2w.vy = toV(l.fy); l2w.vz = toV(l.fz);
  res.dir = l2w * localDir;
  res.pdf = pdf2d * 0.050660591821168885722f * (1.f / sinTheta);
  res.dist = INFINITY;
  uv.y = clampf(uv.y, halfTexel.y, 1.0f - halfTexel.y);
  const vec4f t = Texture2D_get(l.hdriMap, uv, FLT_MIN);
  res.weight = make_vec3f(t.x, t.y, t.z) * toV(l.radiance) / res.pdf;
  return res;
}
static __device__ __noinline__ LightEval HDRILight_eval(const OcLight &l, vec3f dir, float maxDist)
{
  LightEval res; res.radiance = make_vec3f(0.f); res.pdf = 0.f;
  if (INFINITY > maxDist) return res;
  const OcTexture2D &map = *l.hdriMap;
  linear3f w2l; w2l.vx = toV(l.wx); w2l.vy = toV(l.wy); w2l.vz = toV(l.wz);
  const vec3f localDir = w2l * dir;
  const float u = atan2f(localDir.y, localDir.x) * OC_ONE_OVER_TWO_PI;
  const float v = acosf(clampf(localDir.z, -1.f, 1.f)) * OC_ONE_OVER_PI;
  const vec2f halfTexel = make_vec2f((1.f / (float)map.sizeX) * 0.5f, (1.f / (float)map.sizeY) * 0.5f);
  const vec2f uvc = make_vec2f(u, clampf(v, halfTexel.y, 1.0f - halfTexel.y));
  const vec4f t = Texture2D_get(l.hdriMap, uvc, FLT_MIN);
  res.radiance = make_vec3f(t.x, t.y, t.z) * toV(l.radiance);
  vec2f uv = make_vec2f(uvc.x + halfTexel.x, uvc.y + halfTexel.y);
  uv.x = uv.x - floorf(uv.x);
  // Distribution2D_pdf
  const int idy = min(max((int)(uv.y * map.sizeY), 0), map.sizeY - 1);
  const float befy = idy == 0 ? 0.0f : __ldg(l.hdriCdfY + idy - 1);
  const float pdf_y = (__ldg(l.hdriCdfY + idy) - befy) * map.sizeY;
  const int idx = min(max((int)(uv.x * map.sizeX), 0), map.sizeX - 1);
  const int x0 = idy * map.sizeX;
  const float befx = idx == 0 ? 0.0f : __ldg(l.hdriCdfX + x0 + idx - 1);
  const float pdf_x = (__ldg(l.hdriCdfX + x0 + idx) - befx) * map.sizeX;
  res.pdf = pdf_x * pdf_y;
  res.pdf *= 0.050660591821168885722f * (1.f / sqrtf(1.f - sqr(localDir.z)));
  return res;
}

__device__ __forceinline__ LightSample Light_sample(const OcLight &l, const DifferentialGeometry &dg, vec2f s)
{
  LightSample r;
  if (l.type == OC_LIGHT_QUAD)
    return QuadLight_sample(l, dg, s);
  if (l.type == OC_LIGHT_SPOT)
    return SpotLight_sample(l, dg, s);
  if (l.type == OC_LIGHT_SPHERE)
    return PointLight_sample(l, dg, s);
  if (l.type == OC_LIGHT_AMBIENT) {
    if (reduce_max(vabs(dg.Ns)) > 0) {
      const vec3f localDir = cosineSampleHemisphere(s);
      r.dir = frame(dg.Ns) * localDir;
      r.pdf = localDir.z * OC_ONE_OVER_PI;
    } else {
      r.dir = uniformSampleSphere(1.f, s);
      r.pdf = 1.f / OC_FOUR_PI;
    }
    r.dist = INFINITY;
    r.weight = toV(l.radiance) * (1.f / r.pdf);
    return r;
  }
  // distant
  r.dir = toV(l.fz);
  r.dist = INFINITY;
  r.pdf = l.pdf;
  if (l.cosAngle < 0.99999988f) {
    linear3f f; f.vx = toV(l.fx); f.vy = toV(l.fy); f.vz = toV(l.fz);
    r.dir = f * uniformSampleCone(l.cosAngle, s);
  }
  r.weight = toV(l.radiance);
  return r;
}
__device__ __forceinline__ LightEval Light_eval(const OcLight &l, vec3f P, vec3f Ns, vec3f dir, float minDist, float maxDist)
{
  LightEval r; r.radiance = make_vec3f(0.f); r.pdf = 0.f;
  if (l.type == OC_LIGHT_QUAD)
    return QuadLight_eval(l, P, dir, minDist, maxDist);
  if (l.type == OC_LIGHT_SPHERE)
    return PointLight_eval(l, P, dir, minDist, maxDist);
  if (l.type == OC_LIGHT_SPOT)
    return SpotLight_eval(l, P, dir, minDist, maxDist);
  if (l.type == OC_LIGHT_AMBIENT) {
    r.radiance = INFINITY <= maxDist ? toV(l.radiance) : make_vec3f(0.f);
    if (reduce_max(vabs(Ns)) > 0) r.pdf = fmaxf(dot(Ns, dir), 0.f) * OC_ONE_OVER_PI;
    else r.pdf = 1.f / OC_FOUR_PI;
    return r;
  }
  if (INFINITY <= maxDist && l.cosAngle < 0.99999988f && dot(toV(l.fz), dir) > l.cosAngle) {
    r.radiance = toV(l.radiance) * l.pdf;
    r.pdf = l.pdf;
  }
  return r;
}

// Distribution1D_sample (math/Distribution1D.ih:26-57)
__device__ __forceinline__ Sample1D Distribution1D_sample(int size, const float *cdf, int start, float s)
{
  int first = start, len = size;
  while (len > 0) {
    const int half = len >> 1, mid = first + half;
    if (s < __ldg(cdf + mid)) len = half;
    else { first = mid + 1; len -= half + 1; }
  }
  Sample1D r;
  r.idx = first - start;
  const float bef = first == start ? 0.0f : __ldg(cdf + first - 1);
  const float dpdf = __ldg(cdf + first) - bef;
  r.prob = dpdf;
  r.pdf = dpdf * size;
  r.frac = (s - bef) * (1.f / dpdf);
  return r;
}

// pixel filter (pf/PixelFilter.ih:20-28, pf/LUTPixelFilter.ispc:9-15, Distribution2D.ispc:9-25)
static __device__ __noinline__ LightSample GeometryLight_sample(const OcWorld *world, const OcRenderer &ren, const OcLight &l,
    const DifferentialGeometry &dg, vec2f s)
{
  LightSample res;
  res.weight = make_vec3f(0.0f); res.dir = make_vec3f(0.f); res.dist = INFINITY; res.pdf = 0.f;
  const OcModel *model = l.geoModel;
  const OcGeometry *geo = model->geom;
  const Sample1D sample = Distribution1D_sample((int)l.geoNumPrimitives, l.geoDistribution, 0, s.x);
  const uint32_t primID = (uint32_t)__ldg(l.geoPrimIDs + sample.idx);
  const OcMaterial *mat = getMaterial(world, ren, model, primID);
  if (!mat) return res;
  const vec2f ns = make_vec2f(sample.frac, s.y); // reproject s.x
  const OcInstance *inst = l.geoInstance;
  const affine3f xfm = toA(inst->xfm), rxfm = toA(inst->rcp_xfm);
  const SampleAreaRes as = geo->type == OC_GEOM_SPHERES ? Spheres_sampleArea(geo, primID, xfm, rxfm, ns) : Mesh_sampleArea(geo, primID, xfm, ns);
  const vec3f dir = as.pos - dg.P;
  const float dist = length(dir);
  res.dir = dir / dist;
  const float rcosd = 1.f / fabsf(dot(as.normal, res.dir));
  res.pdf = l.geoPdf * sqr(dist) * rcosd;
  res.dist = dist - fmaxf(as.epsilon, calcEpsilon(as.pos, dist)) * rcosd;
  DifferentialGeometry texDg;
  texDg.primID = 0xffffffffu; texDg.st = as.st; texDg.pixelFootprint = FLT_MIN; texDg.P = as.pos; texDg.Ns = as.normal; texDg.Ng = as.normal;
  texDg.material = mat;
  const vec3f radiance = materialEmission(texDg);
  res.weight = radiance * (1.f / res.pdf);
  return res;
}

__device__ __forceinline__ vec2f pixelFilterSample(const OcRenderer &r, vec2f s)
{
  if (r.pfType == OC_PF_POINT) return make_vec2f(0.f, 0.f);
  if (r.pfType == OC_PF_BOX) return (s + make_vec2f(-0.5f, -0.5f)) * r.pfWidth;
  const Sample1D sy = Distribution1D_sample((int)r.pfLutH, r.pfCdfY, 0, s.y);
  const int x0 = sy.idx * (int)r.pfLutW;
  const Sample1D sx = Distribution1D_sample((int)r.pfLutW, r.pfCdfX, x0, s.x);
  const vec2f uv = make_vec2f((sx.idx + sx.frac) * (1.f / (float)r.pfLutW), (sy.idx + sy.frac) * (1.f / (float)r.pfLutH));
  return (uv + make_vec2f(-0.5f, -0.5f)) * r.pfWidth;
}

// PerspectiveCamera_initRay (camera/PerspectiveCamera.ispc:16-64), OrthographicCamera_initRay
// (camera/OrthographicCamera.ispc:12-50); no stereo / motion blur
// side-by-side / top-bottom stereo (PerspectiveCamera.ispc:28-40): the half of the image the PIXEL lies in picks the eye, the
// sample position is stretched back over the whole image plane; returns the eye's offset sign (0: no per-pixel stereo)
__device__ __forceinline__ float cameraStereoSplit(const OcCamera &cam, vec2f &screen, vec2f pixelCenter)
{
  if (cam.stereoMode < 3u) return 0.f;
  const bool sbs = cam.stereoMode == 3u;
  float &split = sbs ? screen.x : screen.y;
  const float pc = sbs ? pixelCenter.x : pixelCenter.y;
  split *= 2.f;
  if (pc < 0.5f) return -1.f;
  split -= 1.f;
  return 1.f;
}
// PanoramicCamera_initRay (camera/PanoramicCamera.ispc:14-92): latitude / longitude directions around the camera frame, the eye
// offset perpendicular to each ray's direction; returns sin(theta) for the ray cone
// (by value: a Ray handed over by reference would move the callers' ray into local memory)
struct PanoRay { vec3f org, dir; float sinTheta; };
__device__ __forceinline__ PanoRay panoramicInitRay(const OcCamera &cam, vec2f screen, vec2f pixelCenter)
{
  float offset = 0.f;
  if (cam.stereoMode == 1u) offset = -cam.ipdOffset.x;
  else if (cam.stereoMode == 2u) offset = cam.ipdOffset.x;
  else offset = cameraStereoSplit(cam, screen, pixelCenter) * cam.ipdOffset.x;
  screen = make_vec2f(lerpf(screen.x, cam.subImage[0], cam.subImage[2]), lerpf(screen.y, cam.subImage[1], cam.subImage[3]));
  const float phi = OC_TWO_PI * screen.x, theta = OC_PI * screen.y;
  float sinTheta, cosTheta;
  sincosf(theta, &sinTheta, &cosTheta);
  const vec3f localDir = cartesian(phi, sinTheta, cosTheta);
  linear3f f; f.vx = toV(cam.du_size); f.vy = toV(cam.dv_up); f.vz = toV(cam.dir_00);
  const vec3f dir = f * make_vec3f(-localDir.y, -localDir.z, localDir.x);
  PanoRay r;
  r.org = toV(cam.org) + offset * cross(dir, f.vy); r.dir = dir; r.sinTheta = sinTheta;
  return r;
}
__device__ __forceinline__ void cameraInitRay(const OcCamera &cam, Ray &ray, vec2f screen, vec2f lens, vec2f pixelCenter)
{
  if (cam.type == OC_CAMERA_PANORAMIC) {
    const PanoRay r = panoramicInitRay(cam, screen, pixelCenter);
    setRay(ray, r.org, r.dir, cam.nearClip, INFINITY);
    return;
  }
  vec3f org = toV(cam.org);
  if (cam.type != OC_CAMERA_ORTHOGRAPHIC) {
    const float eye = cameraStereoSplit(cam, screen, pixelCenter);
    if (eye < 0.f) org = org - toV(cam.ipdOffset);
    else if (eye > 0.f) org = org + toV(cam.ipdOffset);
  }
  screen = make_vec2f(lerpf(screen.x, cam.subImage[0], cam.subImage[2]), lerpf(screen.y, cam.subImage[1], cam.subImage[3]));
  if (cam.type == OC_CAMERA_ORTHOGRAPHIC) {
    setRay(ray, toV(cam.org) + screen.x * toV(cam.du_size) + screen.y * toV(cam.dv_up), toV(cam.dir_00), cam.nearClip, INFINITY);
    return;
  }
  vec3f dir = toV(cam.dir_00) + screen.x * toV(cam.du_size) + screen.y * toV(cam.dv_up);
  if (cam.scaledAperture > 0.f) {
    const float r = sqrtf(lens.x) * cam.scaledAperture;
    float sp, cp;
    sincosf(OC_TWO_PI * lens.y, &sp, &cp);
    const vec3f lp = ((r * cp) * toV(cam.du_size)) + (((r * sp) * cam.aspect) * toV(cam.dv_up));
    org = org + lp;
    dir = dir - lp;
  }
  setRay(ray, org, normalize(dir), cam.nearClip, INFINITY);
}
// the ray cone a camera ray starts with (PerspectiveCamera.ispc:43-63, OrthographicCamera.ispc:50-53), before the
// renderer scales it to pixel units; `screen` is the sample position before the sub-image mapping
__device__ __forceinline__ RayCone cameraInitRayCone(const OcCamera &cam, vec2f screen, vec2f pixelCenter)
{
  RayCone rc;
  const float subH = fabsf(cam.subImage[3] - cam.subImage[1]);
  if (cam.type == OC_CAMERA_PANORAMIC) { // PanoramicCamera.ispc:84-91
    rc.width = 0.f;
    rc.dwdt = OC_PI * panoramicInitRay(cam, screen, pixelCenter).sinTheta * subH;
    if (cam.stereoMode == 4u) rc.dwdt *= 2.f;
    return rc;
  }
  if (cam.type == OC_CAMERA_ORTHOGRAPHIC) {
    rc.width = cam.imgPlaneSizeY * subH; rc.dwdt = 0.f;
    return rc;
  }
  cameraStereoSplit(cam, screen, pixelCenter);
  screen = make_vec2f(lerpf(screen.x, cam.subImage[0], cam.subImage[2]), lerpf(screen.y, cam.subImage[1], cam.subImage[3]));
  const vec3f dir = toV(cam.dir_00) + screen.x * toV(cam.du_size) + screen.y * toV(cam.dv_up);
  const float coneShortening = 1.0f / sqrtf(dot(dir, dir));
  rc.width = 0.f;
  rc.dwdt = cam.imgPlaneSizeY * subH * coneShortening;
  if (cam.stereoMode == 4u) rc.dwdt *= 2.f;
  return rc;
}

} // namespace oc
