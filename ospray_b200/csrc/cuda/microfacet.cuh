// Microfacet building blocks of the reference's material library: GGX distribution with visible-normal sampling,
// conductor Fresnel, the albedo tables of the energy compensation, smooth and rough conductor BSDFs.
// reference: render/bsdfs/GGXDistribution.ih:16-178, MicrofacetDistribution.ih:10-22, Fresnel.ih:53-81,
// MicrofacetAlbedo.ih:12-168 (what a table entry integrates), MicrofacetAlbedoTables.ispc:10-160 (the tables),
// MicrofacetAlbedoTables.ih:10-150 + math/interpolation.ih:11-86 (lookups), Conductor.ih:17-45, MicrofacetConductor.ih:26-138.
#pragma once

namespace oc {

// table block layout (floats), built once per device by albedoTablesKernel: sizes of MicrofacetAlbedoTablesShared.h:6-12
#define OC_MFA_SIZE 32  // albedo_dir [32 x 32] (cosThetaO fastest, roughness), albedo_avg [32]
#define OC_MFD_SIZE 16  // dielectric tables [16 x 16 x 16] (cosThetaO, eta, roughness) and [16 x 16] averages
#define OC_MFT_ALBEDO_DIR 0
#define OC_MFT_ALBEDO_AVG (OC_MFT_ALBEDO_DIR + OC_MFA_SIZE * OC_MFA_SIZE)
#define OC_MFT_DIEL_DIR (OC_MFT_ALBEDO_AVG + OC_MFA_SIZE)
#define OC_MFT_DIEL_AVG (OC_MFT_DIEL_DIR + OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE)
#define OC_MFT_DIELRCP_DIR (OC_MFT_DIEL_AVG + OC_MFD_SIZE * OC_MFD_SIZE)
#define OC_MFT_DIELRCP_AVG (OC_MFT_DIELRCP_DIR + OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE)
#define OC_MFT_DIELREFL_DIR (OC_MFT_DIELRCP_AVG + OC_MFD_SIZE * OC_MFD_SIZE)
#define OC_MFT_DIELREFL_AVG (OC_MFT_DIELREFL_DIR + OC_MFD_SIZE * OC_MFD_SIZE * OC_MFD_SIZE)
#define OC_MFS_SIZE 16  // sheen albedo table [16 x 16] (cosThetaO fastest, roughness)
#define OC_MFT_SHEEN_DIR (OC_MFT_DIELREFL_AVG + OC_MFD_SIZE * OC_MFD_SIZE)
#define OC_MFT_TOTAL (OC_MFT_SHEEN_DIR + OC_MFS_SIZE * OC_MFS_SIZE)

__device__ __forceinline__ float rcpSafe(float f) { return 1.f / (fabsf(f) < 1e-8f ? (f < 0.f ? -1e-8f : 1e-8f) : f); }

// ---- GGX --------------------------------------------------------------------------------------------------------------
struct GGX { vec2f alpha; };
__device__ __forceinline__ GGX makeGGX(float roughness, float anisotropy)
{
  const float aspect = sqrtf(1.f - 0.9f * anisotropy);
  GGX g;
  g.alpha = make_vec2f(fmaxf(sqr(roughness) / aspect, 0.001f), fmaxf(sqr(roughness) * aspect, 0.001f));
  return g;
}
__device__ __forceinline__ float ggxD(const GGX &g, vec3f wh)
{
  const float cosTheta2 = sqr(wh.z);
  const float e = (sqr(wh.x / g.alpha.x) + sqr(wh.y / g.alpha.y)) / cosTheta2;
  return 1.f / (OC_PI * g.alpha.x * g.alpha.y * sqr(cosTheta2 * (1.f + e)));
}
__device__ __forceinline__ float ggxLambda(const GGX &g, vec3f wo)
{
  const float invA2 = (sqr(wo.x * g.alpha.x) + sqr(wo.y * g.alpha.y)) / sqr(wo.z);
  return 0.5f * (-1.f + sqrtf(1.f + invA2));
}
__device__ __forceinline__ float ggxG1(const GGX &g, vec3f wo, float cosThetaOH)
{
  if (wo.z * cosThetaOH <= 0.f) return 0.f;
  return 1.f / (1.f + ggxLambda(g, wo));
}
__device__ __forceinline__ float ggxG2(const GGX &g, vec3f wo, vec3f wi, float cosThetaOH, float cosThetaIH)
{
  if (wo.z * cosThetaOH <= 0.f || wi.z * cosThetaIH <= 0.f) return 0.f;
  return 1.f / (1.f + ggxLambda(g, wo) + ggxLambda(g, wi));
}
__device__ __forceinline__ float ggxEvalVisible(const GGX &g, vec3f wh, vec3f wo, float cosThetaOH, float &pdf)
{
  const float D = ggxD(g, wh);
  pdf = ggxG1(g, wo, cosThetaOH) * fabsf(cosThetaOH) * D / fabsf(wo.z);
  return D;
}
// visible-normal sampling [Heitz 2017] (GGXDistribution.ih:140-178); wo in the upper hemisphere
__device__ __forceinline__ vec3f ggxSampleVisible(const GGX &g, vec3f wo, float &pdf, vec2f s)
{
  const vec3f V = normalize(make_vec3f(g.alpha.x * wo.x, g.alpha.y * wo.y, wo.z));
  const vec3f T1 = (V.z < 0.9999f) ? normalize(cross(V, make_vec3f(0.f, 0.f, 1.f))) : make_vec3f(1.f, 0.f, 0.f);
  const vec3f T2 = cross(T1, V);
  const float a = 1.f / (1.f + V.z);
  const float r = sqrtf(s.x);
  const float phi = (s.y < a) ? s.y / a * OC_PI : OC_PI + (s.y - a) / (1.f - a) * OC_PI;
  float sinPhi, cosPhi;
  sincosf(phi, &sinPhi, &cosPhi);
  const float P1 = r * cosPhi;
  const float P2 = r * sinPhi * ((s.y < a) ? 1.f : V.z);
  vec3f wh = P1 * T1 + P2 * T2 + sqrtf(fmaxf(0.f, 1.f - P1 * P1 - P2 * P2)) * V;
  wh = normalize(make_vec3f(g.alpha.x * wh.x, g.alpha.y * wh.y, fmaxf(0.f, wh.z)));
  const float cosThetaOH = dot(wo, wh);
  pdf = ggxG1(g, wo, cosThetaOH) * fabsf(cosThetaOH) * ggxD(g, wh) / fabsf(wo.z);
  return wh;
}

// ---- Fresnel ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fresnelDielectricAvg(float eta) // Fresnel.ih:53-59
{
  if (eta < 1.f) eta = 1.f / eta;
  return (eta - 1.f) / (4.08567f + 1.00071f * eta);
}
__device__ __forceinline__ float fresnelConductor1(float cosI, float eta, float k) // Fresnel.ih:73-81
{
  const float tmp = sqr(eta) + sqr(k);
  const float Rpar = (tmp * sqr(cosI) - eta * (2.0f * cosI) + 1.f) * (1.f / (tmp * sqr(cosI) + eta * (2.0f * cosI) + 1.f));
  const float Rper = (tmp - 2.0f * eta * cosI + sqr(cosI)) * (1.f / (tmp + 2.0f * eta * cosI + sqr(cosI)));
  return 0.5f * (Rpar + Rper);
}
__device__ __forceinline__ vec3f fresnelConductorRGB(float cosI, vec3f eta, vec3f k) // FresnelConductorRGBUniform_eval
{
  return make_vec3f(fresnelConductor1(cosI, eta.x, k.x), fresnelConductor1(cosI, eta.y, k.y), fresnelConductor1(cosI, eta.z, k.z));
}
__device__ __forceinline__ vec3f reflectDir(vec3f I, vec3f N, float cosI) { return (2.0f * cosI) * N - I; }
__device__ __forceinline__ vec3f refractDir(vec3f I, vec3f N, float cosI, float cosT, float eta) { return eta * (cosI * N - I) - cosT * N; }

// ---- table lookups (math/interpolation.ih) -------------------------------------------------------------------------
__device__ __forceinline__ float mfLerp(float f, float a, float b) { return (1.f - f) * a + f * b; }
__device__ __forceinline__ float interp1D(float x, const float *f, int size)
{
  const float xc = clampf(x, 0.f, (float)(size - 1));
  const float s = xc - floorf(xc);
  const int x0 = min((int)xc, size - 1), x1 = min(x0 + 1, size - 1);
  return mfLerp(s, __ldg(f + x0), __ldg(f + x1));
}
__device__ __forceinline__ float interp2D(float px, float py, const float *f, int nx, int ny_)
{
  const float xc = clampf(px, 0.f, (float)(nx - 1)), yc = clampf(py, 0.f, (float)(ny_ - 1));
  const float sx = xc - floorf(xc), sy = yc - floorf(yc);
  const int x0 = min((int)xc, nx - 1), x1 = min(x0 + 1, nx - 1), y0 = min((int)yc, ny_ - 1), y1 = min(y0 + 1, ny_ - 1);
  const float f0 = mfLerp(sx, __ldg(f + x0 + y0 * nx), __ldg(f + x1 + y0 * nx));
  const float f1 = mfLerp(sx, __ldg(f + x0 + y1 * nx), __ldg(f + x1 + y1 * nx));
  return mfLerp(sy, f0, f1);
}
__device__ __forceinline__ float interp3D(float px, float py, float pz, const float *f, int n)
{
  const float xc = clampf(px, 0.f, (float)(n - 1)), yc = clampf(py, 0.f, (float)(n - 1)), zc = clampf(pz, 0.f, (float)(n - 1));
  const float sx = xc - floorf(xc), sy = yc - floorf(yc), sz = zc - floorf(zc);
  const int x0 = min((int)xc, n - 1), x1 = min(x0 + 1, n - 1), y0 = min((int)yc, n - 1), y1 = min(y0 + 1, n - 1);
  const int z0 = min((int)zc, n - 1), z1 = min(z0 + 1, n - 1);
  const int ny = n, nz = n * n;
  const float f00 = mfLerp(sx, __ldg(f + x0 + y0 * ny + z0 * nz), __ldg(f + x1 + y0 * ny + z0 * nz));
  const float f01 = mfLerp(sx, __ldg(f + x0 + y1 * ny + z0 * nz), __ldg(f + x1 + y1 * ny + z0 * nz));
  const float f10 = mfLerp(sx, __ldg(f + x0 + y0 * ny + z1 * nz), __ldg(f + x1 + y0 * ny + z1 * nz));
  const float f11 = mfLerp(sx, __ldg(f + x0 + y1 * ny + z1 * nz), __ldg(f + x1 + y1 * ny + z1 * nz));
  return mfLerp(sz, mfLerp(sy, f00, f01), mfLerp(sy, f10, f11));
}
__device__ __forceinline__ float mfAlbedo(const float *t, float cosThetaO, float roughness)
{
  return interp2D(cosThetaO * (OC_MFA_SIZE - 1), roughness * (OC_MFA_SIZE - 1), t + OC_MFT_ALBEDO_DIR, OC_MFA_SIZE, OC_MFA_SIZE);
}
__device__ __forceinline__ float mfAlbedoAvg(const float *t, float roughness)
{
  return interp1D(roughness * (OC_MFA_SIZE - 1), t + OC_MFT_ALBEDO_AVG, OC_MFA_SIZE);
}
// MicrofacetDielectricReflectionAlbedoTable_eval / _evalAvg: eta in [1/3, 1]
__device__ __forceinline__ float mfDielReflAlbedo(const float *t, float cosThetaO, float eta, float roughness)
{
  const float minEta = 1.f / 3.f, maxEta = 1.f;
  const float etaParam = (eta - minEta) / (maxEta - minEta);
  return interp3D(cosThetaO * (OC_MFD_SIZE - 1), etaParam * (OC_MFD_SIZE - 1), roughness * (OC_MFD_SIZE - 1), t + OC_MFT_DIELREFL_DIR, OC_MFD_SIZE);
}
__device__ __forceinline__ float mfDielReflAlbedoAvg(const float *t, float eta, float roughness)
{
  const float minEta = 1.f / 3.f, maxEta = 1.f;
  const float etaParam = (eta - minEta) / (maxEta - minEta);
  return interp2D(etaParam * (OC_MFD_SIZE - 1), roughness * (OC_MFD_SIZE - 1), t + OC_MFT_DIELREFL_AVG, OC_MFD_SIZE, OC_MFD_SIZE);
}

// ---- what a table entry integrates (MicrofacetAlbedo.ih) -----------------------------------------------------------
__device__ __forceinline__ float mfAlbedoSample(float cosThetaO, const GGX &g, vec2f s)
{
  cosThetaO = fmaxf(cosThetaO, 1e-6f);
  const vec3f wo = make_vec3f(cos2sin(cosThetaO), 0.f, cosThetaO);
  float whPdf;
  const vec3f wh = ggxSampleVisible(g, wo, whPdf, s);
  const float cosThetaOH = dot(wo, wh);
  const vec3f wi = reflectDir(wo, wh, cosThetaOH);
  if (wi.z <= 0.f) return 0.f;
  const float G = ggxG2(g, wo, wi, cosThetaOH, dot(wi, wh));
  return G * rcpSafe(ggxG1(g, wo, cosThetaOH));
}
// reflectionOnly: MicrofacetDielectricReflectionAlbedo_sample, else MicrofacetDielectricAlbedo_sample
__device__ __forceinline__ float mfDielAlbedoSample(float cosThetaO, float eta, const GGX &g, vec2f s, bool reflectionOnly)
{
  cosThetaO = fmaxf(cosThetaO, 1e-6f);
  const vec3f wo = make_vec3f(cos2sin(cosThetaO), 0.f, cosThetaO);
  float whPdf;
  const vec3f wh = ggxSampleVisible(g, wo, whPdf, s);
  const float cosThetaOH = dot(wo, wh);
  float cosThetaTH;
  const float F = reflectionOnly ? fresnelDielectric(cosThetaOH, eta) : fresnelDielectricEx(cosThetaOH, cosThetaTH, eta);
  float weight = 0.f;
  vec3f wi = reflectDir(wo, wh, cosThetaOH);
  if (wi.z > 0.f) {
    const float G = ggxG2(g, wo, wi, cosThetaOH, dot(wi, wh));
    weight += F * (G * rcpSafe(ggxG1(g, wo, cosThetaOH)));
  }
  if (reflectionOnly) return weight;
  wi = refractDir(wo, wh, cosThetaOH, cosThetaTH, eta);
  if (wi.z < 0.f) {
    const float G = ggxG2(g, wo, wi, cosThetaOH, dot(wi, wh));
    weight += (1.f - F) * (G * rcpSafe(ggxG1(g, wo, cosThetaOH)));
  }
  return weight;
}

// ---- sheen distribution (SheenDistribution.ih:10-88; [Conty and Kulla 2017]) ---------------------------------------------
__device__ __forceinline__ float sheenD(float r, float cosTheta)
{
  const float sinTheta = cos2sin(cosTheta);
  const float invr = 1.f / r;
  return (2.f + invr) * powf(sinTheta, invr) * (float)(0.5 / 3.14159265358979323846);
}
__device__ __forceinline__ float sheenL(float r, float x)
{
  const float w0 = sqr(1.f - r), w1 = 1.f - w0;
  const float a = w0 * 25.3245f + w1 * 21.5473f;
  const float b = w0 * 3.32435f + w1 * 3.82987f;
  const float c = w0 * 0.16801f + w1 * 0.19823f;
  const float d = w0 * -1.27393f + w1 * -1.97760f;
  const float e = w0 * -4.85967f + w1 * -4.32054f;
  return a / (1.f + b * powf(x, c)) + d * x + e;
}
__device__ __forceinline__ float sheenLambda(float r, float cosTheta)
{
  if (cosTheta < 0.5f) return expf(sheenL(r, cosTheta));
  return expf(2.f * sheenL(r, 0.5f) - sheenL(r, 1.f - cosTheta));
}
__device__ __forceinline__ float sheenLambdaI(float r, float cosTheta)
{
  const float x = 1.f + 2.f * sqr(sqr(sqr(1.f - cosTheta)));
  return powf(sheenLambda(r, cosTheta), x);
}
__device__ __forceinline__ float sheenG2(float r, float cosThetaO, float cosThetaI, float cosThetaOH, float cosThetaIH)
{
  if (cosThetaO * cosThetaOH <= 0.f || cosThetaI * cosThetaIH <= 0.f) return 0.f;
  return 1.f / (1.f + sheenLambda(r, cosThetaO) + sheenLambdaI(r, cosThetaI));
}
__device__ __forceinline__ vec3f uniformSampleHemisphere(vec2f s) // math/sampling.ih:33-39
{
  const float phi = (float)(2.0 * 3.14159265358979323846) * s.x;
  const float cosTheta = s.y, sinTheta = cos2sin(s.y);
  float sp, cp;
  sincosf(phi, &sp, &cp);
  return make_vec3f(cp * sinTheta, sp * sinTheta, cosTheta);
}
#define OC_ONE_OVER_TWO_PI ((float)(0.5 / 3.14159265358979323846))
// MicrofacetSheenAlbedo_sample (MicrofacetAlbedo.ih:116-141)
__device__ __forceinline__ float mfSheenAlbedoSample(float cosThetaO, float r, vec2f s)
{
  cosThetaO = fmaxf(cosThetaO, 1e-6f);
  const vec3f wo = make_vec3f(cos2sin(cosThetaO), 0.f, cosThetaO);
  const vec3f wi = uniformSampleHemisphere(s);
  const vec3f wh = normalize(wo + wi);
  const float D = sheenD(r, wh.z);
  const float G = sheenG2(r, cosThetaO, wi.z, dot(wo, wh), dot(wi, wh));
  return D * G * (1.f / (4.f * cosThetaO * OC_ONE_OVER_TWO_PI));
}
// MicrofacetSheenAlbedoTable_eval (MicrofacetAlbedoTables.ih:145-153)
__device__ __forceinline__ float mfSheenAlbedo(const float *t, float cosThetaO, float roughness)
{
  return interp2D(cosThetaO * (OC_MFS_SIZE - 1), roughness * (OC_MFS_SIZE - 1), t + OC_MFT_SHEEN_DIR, OC_MFS_SIZE, OC_MFS_SIZE);
}

// ---- conductors -------------------------------------------------------------------------------------------------------
struct ConductorBSDF
{
  linear3f frame;
  vec3f eta, k;       // complex refraction index per channel (FresnelConductorRGBUniform, FresnelConductorArtistic);
                      // schlick: reflectivity at normal incidence (r) and at grazing angles (g) (FresnelSchlick)
  bool schlick;
  GGX mf;
  float roughness, Eavg;
  vec3f fmsScale;
  const float *tables;
};
// Fresnel_dispatch_eval of the conductor's Fresnel term (Fresnel.ih:165-171 Schlick, :73-81 conductor)
__device__ __forceinline__ vec3f conductorFresnel(const ConductorBSDF &b, float cosI)
{
  if (b.schlick) {
    const float c = 1.f - cosI;
    const float f = sqr(sqr(c)) * c;
    return (1.f - f) * b.eta + f * b.k;
  }
  return fresnelConductorRGB(cosI, b.eta, b.k);
}
// FresnelSchlick_evalAvg (Fresnel.ih:175-185)
__device__ __forceinline__ vec3f fresnelSchlickAvg(vec3f r, vec3f g)
{
  const float p = 1.f / 5.f;
  const float p2 = p * p;
  return (2.f * g * p2 + r + 3.f * p * r) * (1.f / (1.f + 3.f * p + 2.f * p2));
}
// Conductor_sample (Conductor.ih:17-29)
__device__ __forceinline__ SampleRes Conductor_sample(const ConductorBSDF &b, vec3f wo)
{
  SampleRes r;
  const vec3f N = b.frame.vz;
  const float c = dot(wo, N);
  r.wi = reflectDir(wo, N, c);
  r.pdf = INFINITY;
  r.type = SC_SPECULAR_REFLECTION;
  r.weight = conductorFresnel(b, c);
  return r;
}
__device__ __forceinline__ vec3f toLocal(const linear3f &f, vec3f v) { return make_vec3f(dot(f.vx, v), dot(f.vy, v), dot(f.vz, v)); }
// MicrofacetConductor_eval (MicrofacetConductor.ih:26-67)
__device__ __forceinline__ EvalRes MicrofacetConductor_eval(const ConductorBSDF &b, vec3f wo, vec3f wi)
{
  EvalRes res; res.value = make_vec3f(0.f); res.pdf = 0.f;
  const vec3f N = b.frame.vz;
  const float cosThetaO = dot(wo, N), cosThetaI = dot(wi, N);
  if (cosThetaO <= 0.f || cosThetaI <= 0.f) return res;
  const vec3f wh = normalize(wi + wo);
  const float cosThetaOH = dot(wo, wh), cosThetaIH = dot(wi, wh);
  const vec3f wo0 = toLocal(b.frame, wo), wi0 = toLocal(b.frame, wi), wh0 = toLocal(b.frame, wh);
  const vec3f F = conductorFresnel(b, cosThetaOH);
  float whPdf;
  const float D = ggxEvalVisible(b.mf, wh0, wo0, cosThetaOH, whPdf);
  const float G = ggxG2(b.mf, wo0, wi0, cosThetaOH, cosThetaIH);
  const float Eo = mfAlbedo(b.tables, cosThetaO, b.roughness), Ei = mfAlbedo(b.tables, cosThetaI, b.roughness);
  const vec3f fms = b.fmsScale * ((1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - b.Eavg))) * cosThetaI);
  res.pdf = whPdf * (1.f / (4.f * fabsf(cosThetaOH)));
  res.value = F * (D * G * (1.f / (4.f * cosThetaO))) + fms;
  return res;
}
// MicrofacetConductor_sample (MicrofacetConductor.ih:69-116)
__device__ __forceinline__ SampleRes MicrofacetConductor_sample(const ConductorBSDF &b, vec3f wo, vec2f s)
{
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  const vec3f N = b.frame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const vec3f wo0 = toLocal(b.frame, wo);
  float whPdf;
  const vec3f wh = b.frame * ggxSampleVisible(b.mf, wo0, whPdf, s);
  res.wi = reflectDir(wo, wh, dot(wo, wh));
  const float cosThetaI = dot(res.wi, N);
  if (cosThetaI <= 0.f) { res.wi = make_vec3f(0.f); return res; }
  const float cosThetaOH = dot(wo, wh), cosThetaIH = dot(res.wi, wh);
  const vec3f wi0 = toLocal(b.frame, res.wi);
  const vec3f F = conductorFresnel(b, cosThetaOH);
  const float G = ggxG2(b.mf, wo0, wi0, cosThetaOH, cosThetaIH);
  const float Eo = mfAlbedo(b.tables, cosThetaO, b.roughness), Ei = mfAlbedo(b.tables, cosThetaI, b.roughness);
  const vec3f fms = b.fmsScale * ((1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - b.Eavg))) * cosThetaI);
  res.type = SC_GLOSSY_REFLECTION;
  res.pdf = whPdf * (1.f / (4.f * fabsf(cosThetaOH)));
  res.weight = F * (G * rcpSafe(ggxG1(b.mf, wo0, cosThetaOH))) + (fms * (1.f / res.pdf));
  return res;
}

// ---- dielectric coatings over a substrate ---------------------------------------------------------------------------
// DielectricLayer.ih / MicrofacetDielectricLayer.ih: a Fresnel (smooth) or GGX (rough) coating whose reflection is added to
// the substrate's BSDF seen through it (refraction ignored, absorption along the refracted path, energy conservation after
// Kulla-Conty / Kelemen). The substrate is any type with eval(wo, wi), sample(wo, s, ss), albedo and scatteringType.
struct LambertSub
{
  linear3f frame;
  vec3f R;
  vec3f albedo;
  uint32_t scatteringType;
  __device__ __forceinline__ void init(const linear3f &f, vec3f r) { frame = f; R = r; albedo = r; scatteringType = SC_DIFFUSE_REFLECTION; }
  __device__ __forceinline__ EvalRes eval(vec3f, vec3f wi) const // Lambert.ih:10-19
  {
    EvalRes r;
    const float c = fmaxf(dot(wi, frame.vz), 0.f);
    r.pdf = c * OC_ONE_OVER_PI;
    r.value = R * OC_ONE_OVER_PI * c;
    return r;
  }
  __device__ __forceinline__ SampleRes sample(vec3f, vec2f s, float) const // Lambert.ih:21-31
  {
    const vec3f l = cosineSampleHemisphere(s);
    SampleRes r;
    r.wi = frame * l; r.pdf = l.z * OC_ONE_OVER_PI; r.type = SC_DIFFUSE_REFLECTION; r.weight = R;
    return r;
  }
};

struct LayerParams
{
  float weight, eta; // eta <= 1 (as stored by the constructors)
  vec3f transmittance;
  float thickness;
  float Favg;        // smooth: fresnelDielectricAvg(eta) * weight
  GGX mf;            // rough
  float roughness, Ravg, Eavg, fmsScale;
  const float *tables;
  vec3f albedo;      // of the layered BSDF
  uint32_t scatteringType;
};
__device__ __forceinline__ vec3f powv(vec3f a, float e) { return make_vec3f(powf(a.x, e), powf(a.y, e), powf(a.z, e)); }
__device__ __forceinline__ float refractCos(float cosI, float eta) { return sqrtf(fmaxf(sqrCosT(cosI, eta), 0.0f)); } // Optics.ih:51-55
__device__ __forceinline__ vec3f lerpv(float f, vec3f a, vec3f b) { return (1.f - f) * a + f * b; }

// DielectricLayer_Constructor (DielectricLayer.ih:27-50)
__device__ __forceinline__ void DielectricLayer_init(LayerParams &L, vec3f subAlbedo, uint32_t subType, float eta, vec3f transmittance,
    float thickness, float weight)
{
  L.Favg = fresnelDielectricAvg(eta) * weight;
  L.albedo = subAlbedo * (make_vec3f(1.0f - weight) + powv(transmittance, thickness) * (1.f - L.Favg));
  L.scatteringType = SC_SPECULAR_REFLECTION | subType;
  L.eta = (eta <= 1.f) ? eta : 1.f / eta;
  L.transmittance = transmittance; L.thickness = thickness; L.weight = weight;
}
// MicrofacetDielectricLayer_Constructor (MicrofacetDielectricLayer.ih:39-77). The reference reads self->eta for Ravg BEFORE
// it assigns it (line :55 vs :68): it sees whatever the shading context's memory held. The intended value is used here.
__device__ __forceinline__ void MicrofacetDielectricLayer_init(LayerParams &L, const float *tables, vec3f subAlbedo, uint32_t subType, float eta,
    vec3f transmittance, float thickness, float roughness, float anisotropy, float weight)
{
  L.tables = tables;
  L.eta = (eta <= 1.f) ? eta : 1.f / eta;
  L.Eavg = mfAlbedoAvg(tables, roughness);
  const float Favg = fresnelDielectricAvg(eta);
  L.fmsScale = Favg * (1.f - L.Eavg) * (1.f / (1.f - Favg * L.Eavg));
  L.Ravg = (mfDielReflAlbedoAvg(tables, L.eta, roughness) + L.fmsScale * (1.f - L.Eavg)) * weight;
  L.albedo = subAlbedo * (make_vec3f(1.0f - weight) + powv(transmittance, thickness) * (1.f - Favg));
  L.scatteringType = SC_GLOSSY_REFLECTION | subType;
  L.transmittance = transmittance; L.thickness = thickness; L.weight = weight;
  L.mf = makeGGX(roughness, anisotropy);
  L.roughness = roughness;
}

// DIELECTRICLAYER_EVAL / _SAMPLE (DielectricLayer.ih:84-198); `albedo`, `scatteringType` = the substrate's
template <class Sub>
__device__ __forceinline__ EvalRes DielectricLayer_eval(const LayerParams &L, vec3f N, const Sub &sub, vec3f wo, vec3f wi)
{
  EvalRes zero; zero.value = make_vec3f(0.f); zero.pdf = 0.f;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return zero;
  EvalRes substrate = sub.eval(wo, wi);
  float cosThetaO1;
  const float F = fresnelDielectricEx(cosThetaO, cosThetaO1, L.eta);
  const float Fo = F * L.weight;
  const float cosThetaI = dot(wi, N);
  float cosThetaI1;
  const float Fi = fresnelDielectricEx(fabsf(cosThetaI), cosThetaI1, L.eta) * L.weight;
  float length = 1.f / cosThetaO1 + 1.f / cosThetaI1;
  if (cosThetaI <= 0.f) length *= 0.5f;
  substrate.value = lerpv(L.weight, substrate.value, substrate.value * powv(L.transmittance, L.thickness * length));
  float T;
  if (sub.scatteringType & ~SC_DIFFUSE) T = fminf(1.f - Fo, 1.f - Fi);
  else T = (1.f - Fo) * (1.f - Fi) * (1.f / (1.f - L.Favg));
  substrate.value = substrate.value * T;
  const float coatPickProb = Fo / ((1.f - F) * reduce_max(sub.albedo) + Fo);
  substrate.pdf *= (1.f - coatPickProb);
  return substrate;
}
template <class Sub>
__device__ __forceinline__ SampleRes DielectricLayer_sample(const LayerParams &L, vec3f N, const Sub &sub, vec3f wo, vec2f s, float ss)
{
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  float cosThetaO1;
  const float F = fresnelDielectricEx(cosThetaO, cosThetaO1, L.eta);
  const float Fo = F * L.weight;
  const float coatPickProb = Fo / ((1.f - F) * reduce_max(sub.albedo) + Fo);
  if (ss < coatPickProb) {
    res.type = SC_SPECULAR_REFLECTION;
    res.wi = reflectDir(wo, N, cosThetaO);
    res.pdf = INFINITY;
    res.weight = make_vec3f(Fo / coatPickProb);
  } else {
    const float ss1 = (ss - coatPickProb) * (1.f / (1.f - coatPickProb));
    res = sub.sample(wo, s, ss1);
    if (reduce_max(res.weight) <= 0.f) return res;
    const float cosThetaI = dot(res.wi, N);
    float cosThetaI1;
    const float Fi = fresnelDielectricEx(fabsf(cosThetaI), cosThetaI1, L.eta) * L.weight;
    float length = 1.f / cosThetaO1 + 1.f / cosThetaI1;
    if (cosThetaI <= 0.f) length *= 0.5f;
    res.weight = lerpv(L.weight, res.weight, res.weight * powv(L.transmittance, L.thickness * length));
    float T;
    if (sub.scatteringType & ~SC_DIFFUSE) T = fminf(1.f - Fo, 1.f - Fi);
    else T = (1.f - Fo) * (1.f - Fi) * (1.f / (1.f - L.Favg));
    res.weight = res.weight * (T * (1.f / (1.f - coatPickProb)));
    res.pdf *= (1.f - coatPickProb);
  }
  return res;
}

// MICROFACETDIELECTRICLAYER_EVAL / _SAMPLE (MicrofacetDielectricLayer.ih:150-387)
template <class Sub>
__device__ __forceinline__ EvalRes MicrofacetDielectricLayer_eval(const LayerParams &L, const linear3f &frame, const Sub &sub, vec3f wo, vec3f wi)
{
  EvalRes zero; zero.value = make_vec3f(0.f); zero.pdf = 0.f;
  const vec3f N = frame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return zero;
  const float cosThetaI = dot(wi, N);
  EvalRes substrate = sub.eval(wo, wi);
  float length = 1.f / refractCos(cosThetaO, L.eta) + 1.f / refractCos(fabsf(cosThetaI), L.eta);
  if (cosThetaI <= 0.f) length *= 0.5f;
  substrate.value = lerpv(L.weight, substrate.value, substrate.value * powv(L.transmittance, L.thickness * length));
  const float Eo = mfAlbedo(L.tables, cosThetaO, L.roughness), Ei = mfAlbedo(L.tables, fabsf(cosThetaI), L.roughness);
  const float Ro = (mfDielReflAlbedo(L.tables, cosThetaO, L.eta, L.roughness) + L.fmsScale * (1.f - Eo)) * L.weight;
  const float Ri = (mfDielReflAlbedo(L.tables, fabsf(cosThetaI), L.eta, L.roughness) + L.fmsScale * (1.f - Ei)) * L.weight;
  float T;
  if (sub.scatteringType & ~SC_DIFFUSE) T = fminf(1.f - Ro, 1.f - Ri);
  else T = (1.f - Ro) * (1.f - Ri) * (1.f / (1.f - L.Ravg));
  substrate.value = substrate.value * T;
  const float coatPickProb = L.weight * Ro / ((1.f - Ro) * reduce_max(sub.albedo) + L.weight * Ro);
  const float substratePickProb = 1.f - coatPickProb;
  if (cosThetaI > 0.f) {
    const vec3f wh = normalize(wo + wi);
    const float cosThetaOH = dot(wo, wh);
    const float F = fresnelDielectric(cosThetaOH, L.eta) * L.weight;
    const float cosThetaIH = dot(wi, wh);
    const vec3f wo0 = toLocal(frame, wo), wi0 = toLocal(frame, wi), wh0 = toLocal(frame, wh);
    float whPdf;
    const float D = ggxEvalVisible(L.mf, wh0, wo0, cosThetaOH, whPdf);
    const float G = ggxG2(L.mf, wo0, wi0, cosThetaOH, cosThetaIH);
    const float fms = L.fmsScale * ((1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - L.Eavg))) * cosThetaI);
    const float coatPdf = whPdf * (1.f / (4.f * cosThetaOH));
    const vec3f coatValue = make_vec3f(F * D * G * (1.f / (4.f * cosThetaO)) + fms);
    substrate.pdf = coatPickProb * coatPdf + substratePickProb * substrate.pdf;
    substrate.value = coatValue + substrate.value;
  } else
    substrate.pdf *= substratePickProb;
  return substrate;
}
template <class Sub>
__device__ __forceinline__ SampleRes MicrofacetDielectricLayer_sample(const LayerParams &L, const linear3f &frame, const Sub &sub, vec3f wo, vec2f s, float ss)
{
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  const vec3f N = frame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const vec3f wo0 = toLocal(frame, wo);
  const float Eo = mfAlbedo(L.tables, cosThetaO, L.roughness);
  // (the sample macro weights only the compensation term with the layer weight: MicrofacetDielectricLayer.ih:262-266)
  const float Ro = mfDielReflAlbedo(L.tables, cosThetaO, L.eta, L.roughness) + L.fmsScale * (1.f - Eo) * L.weight;
  vec3f wh;
  EvalRes substrate;
  const float coatPickProb = L.weight * Ro / ((1.f - Ro) * reduce_max(sub.albedo) + L.weight * Ro);
  const float substratePickProb = 1.f - coatPickProb;
  if (ss < coatPickProb) {
    float whPdf;
    wh = frame * ggxSampleVisible(L.mf, wo0, whPdf, s);
    res.type = SC_GLOSSY_REFLECTION;
    res.wi = reflectDir(wo, wh, dot(wo, wh));
    substrate = sub.eval(wo, res.wi);
  } else {
    const float ss1 = (ss - coatPickProb) * (1.f / substratePickProb);
    res = sub.sample(wo, s, ss1);
    if (reduce_max(res.weight) <= 0.f) return res;
    substrate.pdf = res.pdf;
    substrate.value = res.weight * ((res.type & SC_SPECULAR) ? 1.f : res.pdf);
    wh = normalize(wo + res.wi);
  }
  const float cosThetaI = dot(res.wi, N);
  const float cosThetaOH = dot(wo, wh);
  float length = 1.f / refractCos(cosThetaO, L.eta) + 1.f / refractCos(fabsf(cosThetaI), L.eta);
  if (cosThetaI <= 0.f) length *= 0.5f;
  substrate.value = lerpv(L.weight, substrate.value, substrate.value * powv(L.transmittance, L.thickness * length));
  const float Ei = mfAlbedo(L.tables, fabsf(cosThetaI), L.roughness);
  const float Ri = (mfDielReflAlbedo(L.tables, fabsf(cosThetaI), L.eta, L.roughness) + L.fmsScale * (1.f - Ei)) * L.weight;
  float T;
  if (sub.scatteringType & ~SC_DIFFUSE) T = fminf(1.f - Ro, 1.f - Ri);
  else T = (1.f - Ro) * (1.f - Ri) * (1.f / (1.f - L.Ravg));
  substrate.value = substrate.value * T;
  if (res.type & SC_SPECULAR) res.weight = substrate.value * (1.f / substratePickProb);
  else if (cosThetaI <= 0.f) {
    res.pdf = substratePickProb * substrate.pdf;
    res.weight = substrate.value * (1.f / res.pdf);
  } else {
    const float F = fresnelDielectric(cosThetaOH, L.eta) * L.weight;
    const float cosThetaIH = dot(res.wi, wh);
    const vec3f wi0 = toLocal(frame, res.wi), wh0 = toLocal(frame, wh);
    float whPdf;
    const float D = ggxEvalVisible(L.mf, wh0, wo0, cosThetaOH, whPdf);
    const float G = ggxG2(L.mf, wo0, wi0, cosThetaOH, cosThetaIH);
    const float fms = L.fmsScale * ((1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - L.Eavg))) * cosThetaI);
    const float coatPdf = whPdf * (1.f / (4.f * cosThetaOH));
    const vec3f coatValue = make_vec3f(F * D * G * (1.f / (4.f * cosThetaO)) + fms);
    res.pdf = coatPickProb * coatPdf + substratePickProb * substrate.pdf;
    res.weight = (coatValue + substrate.value) * (1.f / res.pdf);
  }
  return res;
}

// ---- velvet (materials/Velvet.ispc:14-99): MultiBSDF of a Minneart lobe (bsdfs/Minneart.ih) and a Velvety lobe (Velvety.ih) ------
struct VelvetBSDF
{
  linear3f frame;
  vec3f R0, R1;   // reflectance, horizon scattering colour
  float b, f;     // back scattering exponent, horizon falloff
  float imp0, imp1, importanceSum;
  __device__ __forceinline__ EvalRes lobe(int id, vec3f wo, vec3f wi) const
  {
    EvalRes r;
    const float cosThetaI = clampf(dot(wi, frame.vz), 0.f, 1.f);
    r.pdf = cosThetaI * OC_ONE_OVER_PI;
    if (id == 0) {
      const float backScatter = powf(clampf(dot(wo, wi), 0.f, 1.f), b);
      r.value = (backScatter * cosThetaI * OC_ONE_OVER_PI) * R0;
    } else {
      const float cosThetaO = clampf(dot(wo, frame.vz), 0.f, 1.f);
      const float sinThetaO = sqrtf(1.0f - cosThetaO * cosThetaO);
      const float horizonScatter = powf(sinThetaO, f);
      r.value = (horizonScatter * cosThetaI * OC_ONE_OVER_PI) * R1;
    }
    return r;
  }
  __device__ __forceinline__ EvalRes eval(vec3f wo, vec3f wi) const
  {
    EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
    if (imp0 > 0.f) { const EvalRes c = lobe(0, wo, wi); e.value = e.value + c.value * 1.f; e.pdf += c.pdf * imp0; }
    if (imp1 > 0.f) { const EvalRes c = lobe(1, wo, wi); e.value = e.value + c.value * 1.f; e.pdf += c.pdf * imp1; }
    e.pdf *= 1.f / importanceSum;
    return e;
  }
  __device__ __forceinline__ SampleRes sample(vec3f wo, vec2f s, float ss) const
  {
    SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
    if (importanceSum == 0.f) return res;
    const float x = ss * importanceSum;
    int choice;
    float prefix = 0.f;
    if (x < (prefix += imp0)) choice = 0;
    else if (x < (prefix += imp1)) choice = 1;
    else return res;
    res.wi = frame * cosineSampleHemisphere(s);
    res.type = SC_DIFFUSE_REFLECTION;
    const EvalRes ev = lobe(choice, wo, res.wi);
    res.pdf = ev.pdf;
    res.weight = ev.value * (1.f / ev.pdf) * 1.f;
    if ((res.weight.x == 0.f && res.weight.y == 0.f && res.weight.z == 0.f) || res.pdf == 0.0f) { res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0; return res; }
    const float choiceImp = choice == 0 ? imp0 : imp1, otherImp = choice == 0 ? imp1 : imp0;
    EvalRes e;
    e.value = res.weight * res.pdf;
    res.pdf *= choiceImp;
    e.pdf = res.pdf;
    if (otherImp > 0.f) { const EvalRes c = lobe(1 - choice, wo, res.wi); e.value = e.value + c.value * 1.f; e.pdf += c.pdf * otherImp; }
    res.pdf = e.pdf * (1.f / importanceSum);
    res.weight = e.value * (1.f / res.pdf);
    return res;
  }
};

// ---- metallic paint (materials/MetallicPaint.ispc:22-161): a smooth dielectric coat over MultiBSDF(Lambert, flakes =
// MicrofacetConductor with Schlick Fresnel); this is the substrate the DielectricLayer templates take
struct MetallicPaintSub
{
  LambertSub lambert;
  ConductorBSDF flakes;
  float imp0, imp1, importanceSum; // luminance(color), luminance(flakeColor) (0: no flakes)
  vec3f albedo;
  uint32_t scatteringType;
  __device__ __forceinline__ void others(vec3f wo, vec3f wi, int skip, EvalRes &e) const
  {
    if (imp0 > 0.f && skip != 0) { const EvalRes c = lambert.eval(wo, wi); e.value = e.value + c.value * 1.f; e.pdf += c.pdf * imp0; }
    if (imp1 > 0.f && skip != 1) { const EvalRes c = MicrofacetConductor_eval(flakes, wo, wi); e.value = e.value + c.value * 1.f; e.pdf += c.pdf * imp1; }
  }
  __device__ __forceinline__ EvalRes eval(vec3f wo, vec3f wi) const
  {
    EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
    others(wo, wi, -1, e);
    e.pdf *= 1.f / importanceSum;
    return e;
  }
  __device__ __forceinline__ SampleRes sample(vec3f wo, vec2f s, float ss) const
  {
    SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
    if (importanceSum == 0.f) return res;
    const float x = ss * importanceSum;
    int choice;
    float prefix = 0.f;
    if (x < (prefix += imp0)) { choice = 0; res = lambert.sample(wo, s, 0.f); }
    else if (x < (prefix += imp1)) { choice = 1; res = MicrofacetConductor_sample(flakes, wo, s); }
    else return res;
    res.weight = res.weight * 1.f;
    if ((res.weight.x == 0.f && res.weight.y == 0.f && res.weight.z == 0.f) || res.pdf == 0.0f) { res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0; return res; }
    EvalRes e;
    e.value = res.weight * res.pdf;
    res.pdf *= choice == 0 ? imp0 : imp1;
    e.pdf = res.pdf;
    others(wo, res.wi, choice, e);
    res.pdf = e.pdf * (1.f / importanceSum);
    res.weight = e.value * (1.f / res.pdf);
    return res;
  }
};

} // namespace oc
