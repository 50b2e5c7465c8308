// The principled material: the reference's layered uber-BSDF as one statically composed structure.
// reference: render/materials/Principled.ispc:34-428 (Principled_getBSDF), :430-573 (getTransparency), :575-601
// (selectNextMedium), :603-1040 (the eval / sample dispatch through the nesting), bsdfs/MultiBSDF.ih:58-128,
// OrenNayar.ih:18-45, Dielectric.ih:18-50, MicrofacetDielectric.ih:27-268, Fresnel.ih:227-284 (artistic conductor).
//
//   root   MultiBSDF(2)   [0] Transmission(1 - opacity)            [1] sheen( coat( base ) )
//   sheen  passthrough | MicrofacetSheenLayer over the coated base
//   coat   passthrough | DielectricLayer | MicrofacetDielectricLayer over base
//   base   MultiBSDF(3)   [0] metal: Conductor | MicrofacetConductor (artistic Fresnel)
//                         [1] glass: Transmission | Dielectric | MicrofacetDielectric | ThinDielectric | ThinMicrofacetDielectric
//                         [2] plastic: passthrough | DielectricLayer | MicrofacetDielectricLayer over diffuse
//   diffuse MultiBSDF(2)  [0] Lambert | OrenNayar                  [1] LambertTransmission (thin: backlight)
#pragma once

namespace oc {

#define PR_EPS 1e-5f

struct PrDiffuse
{
  linear3f frame;
  vec3f R, albedo;
  float roughness; // OrenNayar when > 0 (the material's roughness >= EPS), else Lambert
  bool orenNayar;
  uint32_t scatteringType;
  float weight, importance, importanceSum; // the enclosing 2-child MultiBSDF (child 0)
  __device__ __forceinline__ EvalRes lobeEval(vec3f wo, vec3f wi) const
  {
    EvalRes r;
    if (!orenNayar) {
      const float c = fmaxf(dot(wi, frame.vz), 0.f);
      r.pdf = c * OC_ONE_OVER_PI; r.value = R * OC_ONE_OVER_PI * c;
      return r;
    }
    const float cosThetaO = fmaxf(dot(wo, frame.vz), 0.f), cosThetaI = fmaxf(dot(wi, frame.vz), 0.f);
    const float s = dot(wo, wi) - cosThetaO * cosThetaI;
    const float sInvt = (s > 0.f) ? (s * rcpSafe(fmaxf(cosThetaO, cosThetaI))) : s;
    r.value = R * (cosThetaI * (1.f + roughness * sInvt) * (1.f / (OC_PI + (OC_PI / 2.f - 2.f / 3.f) * roughness)));
    r.pdf = cosThetaI * OC_ONE_OVER_PI;
    return r;
  }
  __device__ __forceinline__ SampleRes lobeSample(vec3f wo, vec2f s) const
  {
    const vec3f l = cosineSampleHemisphere(s);
    SampleRes r;
    r.wi = frame * l; r.type = SC_DIFFUSE_REFLECTION;
    if (!orenNayar) { r.pdf = l.z * OC_ONE_OVER_PI; r.weight = R; return r; }
    const EvalRes e = lobeEval(wo, r.wi);
    r.pdf = e.pdf; r.weight = e.value * (1.f / e.pdf);
    return r;
  }
  // second child: LambertTransmission (LambertTransmission.ih:10-32), thin materials with backlight
  float weightT, importanceT; // 0: absent
  // DiffuseBSDF_eval / _sample (Principled.ispc:634-661): the MultiBSDF macros over the two children
  __device__ __forceinline__ void evalChildren(vec3f wo, vec3f wi, int skip, EvalRes &e) const
  {
    if (importance > 0.f && skip != 0) {
      const EvalRes c = lobeEval(wo, wi);
      e.value = e.value + c.value * weight;
      e.pdf += c.pdf * importance;
    }
    if (importanceT > 0.f && skip != 1) {
      const float cosThetaI = fmaxf(-dot(wi, frame.vz), 0.f);
      e.value = e.value + (R * OC_ONE_OVER_PI * cosThetaI) * weightT;
      e.pdf += (cosThetaI * OC_ONE_OVER_PI) * importanceT;
    }
  }
  __device__ __forceinline__ EvalRes eval(vec3f wo, vec3f wi) const
  {
    EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
    evalChildren(wo, wi, -1, e);
    e.pdf *= 1.f / importanceSum;
    return e;
  }
  __device__ __forceinline__ SampleRes sample(vec3f wo, vec2f s, float ss) const
  {
    SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
    if (importanceSum == 0.f) return res;
    const float x = ss * importanceSum;
    int choice;
    float prefix = 0.f, choiceImp;
    if (x < (prefix += importance)) {
      choice = 0; choiceImp = importance;
      res = lobeSample(wo, s);
      res.weight = res.weight * weight;
    } else if (x < (prefix += importanceT)) {
      choice = 1; choiceImp = importanceT;
      const vec3f l = cosineSampleHemisphere(s);
      res.wi = neg(frame * l); res.pdf = l.z * OC_ONE_OVER_PI; res.type = SC_DIFFUSE_TRANSMISSION; res.weight = R;
      res.weight = res.weight * weightT;
    } else return res;
    if ((res.weight.x == 0.f && res.weight.y == 0.f && res.weight.z == 0.f) || res.pdf == 0.0f) { res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0; return res; }
    EvalRes e;
    e.value = res.weight * res.pdf;
    res.pdf *= choiceImp;
    e.pdf = res.pdf;
    evalChildren(wo, res.wi, choice, e);
    res.pdf = e.pdf * (1.f / importanceSum);
    res.weight = e.value * (1.f / res.pdf);
    return res;
  }
};

struct PrGlass
{
  linear3f frame;
  int kind; // 0 Transmission(1), 1 Dielectric, 2 MicrofacetDielectric, 3 ThinDielectric, 4 ThinMicrofacetDielectric
  float eta;
  vec3f attenuation;   // thin: log(transmissionColor) / transmissionDepth * thickness
  float thinEavg, thinFmsScale;
  GGX mf;
  float roughness, EavgEta, EavgRcpEta, fmsRatio;
  const float *tables;
  vec3f albedo;
  uint32_t scatteringType;
};
// MicrofacetDielectricAlbedoTable_eval / _evalAvg (MicrofacetAlbedoTables.ih:29-78)
__device__ __forceinline__ float mfDielAlbedo(const float *t, float cosThetaO, float eta, float roughness)
{
  const bool le = eta <= 1.f;
  const float minEta = le ? 1.f / 3.f : 1.f, maxEta = le ? 1.f : 3.f;
  const float etaParam = (eta - minEta) / (maxEta - minEta);
  return interp3D(cosThetaO * (OC_MFD_SIZE - 1), etaParam * (OC_MFD_SIZE - 1), roughness * (OC_MFD_SIZE - 1),
      t + (le ? OC_MFT_DIEL_DIR : OC_MFT_DIELRCP_DIR), OC_MFD_SIZE);
}
__device__ __forceinline__ float mfDielAlbedoAvg(const float *t, float eta, float roughness)
{
  const bool le = eta <= 1.f;
  const float minEta = le ? 1.f / 3.f : 1.f, maxEta = le ? 1.f : 3.f;
  const float etaParam = (eta - minEta) / (maxEta - minEta);
  return interp2D(etaParam * (OC_MFD_SIZE - 1), roughness * (OC_MFD_SIZE - 1), t + (le ? OC_MFT_DIEL_AVG : OC_MFT_DIELRCP_AVG), OC_MFD_SIZE, OC_MFD_SIZE);
}
__device__ __forceinline__ void PrGlass_initMicrofacet(PrGlass &g, const float *tables, float eta, float roughness, float anisotropy)
{
  g.kind = 2; g.eta = eta; g.tables = tables; g.roughness = roughness;
  g.albedo = make_vec3f(0.9f); g.scatteringType = SC_GLOSSY | SC_DIFFUSE;
  g.mf = makeGGX(roughness, anisotropy);
  g.EavgEta = mfDielAlbedoAvg(tables, eta, roughness);
  g.EavgRcpEta = mfDielAlbedoAvg(tables, 1.f / eta, roughness);
  const float FavgEta = fresnelDielectricAvg(eta), FavgRcpEta = fresnelDielectricAvg(1.f / eta);
  const float a = (1.f - FavgEta) * (1.f / (1.f - g.EavgRcpEta));
  const float b = (1.f - FavgRcpEta) * (1.f / (1.f - g.EavgEta));
  const float x = b * (1.f / (a + b));
  g.fmsRatio = 1.f - x * (1.f - FavgEta);
}
__device__ __forceinline__ EvalRes MicrofacetDielectric_evalSingle(const PrGlass &g, vec3f wo, vec3f wi)
{
  EvalRes res; res.value = make_vec3f(0.f); res.pdf = 0.f;
  const vec3f N = g.frame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const float cosThetaI = dot(wi, N);
  const bool isReflection = cosThetaI > 0.f;
  vec3f wh = isReflection ? wo + wi : g.eta * wo + wi;
  wh = normalize(wh);
  if (dot(wh, N) < 0.f) wh = neg(wh);
  const float cosThetaOH = dot(wo, wh), cosThetaIH = dot(wi, wh);
  const float F = fresnelDielectric(cosThetaOH, g.eta);
  float value;
  if (isReflection) {
    res.pdf = F * (1.f / (4.f * fabsf(cosThetaOH)));
    value = F * (1.f / (4.f * cosThetaO));
  } else {
    res.pdf = (1.f - F) * (1.f / sqr(g.eta)) * fabsf(cosThetaIH) * (1.f / sqr(cosThetaOH + (1.f / g.eta) * cosThetaIH));
    value = res.pdf * fabsf(cosThetaOH) * (1.f / cosThetaO);
  }
  const vec3f wo0 = toLocal(g.frame, wo), wi0 = toLocal(g.frame, wi), wh0 = toLocal(g.frame, wh);
  float whPdf;
  const float D = ggxEvalVisible(g.mf, wh0, wo0, cosThetaOH, whPdf);
  const float G = ggxG2(g.mf, wo0, wi0, cosThetaOH, cosThetaIH);
  res.pdf *= whPdf;
  res.value = make_vec3f(value * D * G);
  return res;
}
__device__ __forceinline__ EvalRes MicrofacetDielectric_fms(const PrGlass &g, float Eo, float cosThetaI, bool reflection)
{
  EvalRes fms;
  if (reflection) {
    const float Ei = mfDielAlbedo(g.tables, cosThetaI, g.eta, g.roughness);
    fms.pdf = g.fmsRatio * (cosThetaI * OC_ONE_OVER_PI);
    fms.value = make_vec3f(g.fmsRatio * (1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - g.EavgEta))) * cosThetaI);
  } else {
    const float Ei = mfDielAlbedo(g.tables, fabsf(cosThetaI), 1.f / g.eta, g.roughness);
    fms.pdf = (1.f - g.fmsRatio) * (fabsf(cosThetaI) * OC_ONE_OVER_PI);
    fms.value = make_vec3f((1.f - g.fmsRatio) * (1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - g.EavgRcpEta))) * fabsf(cosThetaI));
  }
  return fms;
}
// ThinMicrofacetDielectric_Constructor / _eval / _sample (ThinMicrofacetDielectric.ih:25-178)
__device__ __forceinline__ void PrGlass_initThinMicrofacet(PrGlass &g, const float *tables, float eta, float roughness, float anisotropy, vec3f attenuation)
{
  g.kind = 4; g.eta = eta; g.tables = tables; g.roughness = roughness; g.attenuation = attenuation;
  g.albedo = expf3(attenuation); g.scatteringType = SC_GLOSSY_REFLECTION | SC_SPECULAR_TRANSMISSION;
  g.mf = makeGGX(roughness, anisotropy);
  g.thinEavg = mfAlbedoAvg(tables, roughness);
  const float Favg = fresnelDielectricAvg(eta);
  g.thinFmsScale = Favg * (1.f - g.thinEavg) * (1.f / (1.f - Favg * g.thinEavg));
}
__device__ __forceinline__ EvalRes ThinMicrofacetDielectric_eval(const PrGlass &g, vec3f wo, vec3f wi)
{
  EvalRes res; res.value = make_vec3f(0.f); res.pdf = 0.f;
  const vec3f N = g.frame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const float cosThetaI = dot(wi, N);
  if (!(cosThetaI > 0.f)) return res;
  const vec3f wh = normalize(wo + wi);
  const float cosThetaOH = dot(wo, wh), cosThetaIH = dot(wi, wh);
  const vec3f wo0 = toLocal(g.frame, wo), wi0 = toLocal(g.frame, wi), wh0 = toLocal(g.frame, wh);
  const float F = fresnelDielectric(cosThetaOH, g.eta);
  float whPdf;
  const float D = ggxEvalVisible(g.mf, wh0, wo0, cosThetaOH, whPdf);
  const float G = ggxG2(g.mf, wo0, wi0, cosThetaOH, cosThetaIH);
  const float Eo = mfAlbedo(g.tables, cosThetaO, g.roughness), Ei = mfAlbedo(g.tables, cosThetaI, g.roughness);
  const float fms = g.thinFmsScale * ((1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - g.thinEavg))) * cosThetaI);
  const float R = mfDielReflAlbedo(g.tables, cosThetaO, g.eta, g.roughness) + g.thinFmsScale * (1.f - Eo);
  res.pdf = whPdf * (1.f / (4.f * cosThetaOH)) * R;
  res.value = make_vec3f(F * D * G * (1.f / (4.f * cosThetaO)) + fms);
  return res;
}
__device__ __forceinline__ SampleRes ThinMicrofacetDielectric_sample(const PrGlass &g, vec3f wo, vec2f s, float ss)
{
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  const vec3f N = g.frame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const float Eo = mfAlbedo(g.tables, cosThetaO, g.roughness);
  const float R = mfDielReflAlbedo(g.tables, cosThetaO, g.eta, g.roughness) + g.thinFmsScale * (1.f - Eo);
  if (ss < R) {
    const vec3f wo0 = toLocal(g.frame, wo);
    float whPdf;
    const vec3f wh = g.frame * ggxSampleVisible(g.mf, wo0, whPdf, s);
    res.wi = reflectDir(wo, wh, dot(wo, wh));
    const float cosThetaI = dot(res.wi, N);
    if (cosThetaI <= 0.f) { res.wi = make_vec3f(0.f); return res; }
    const float cosThetaOH = dot(wo, wh), cosThetaIH = dot(res.wi, wh);
    const vec3f wi0 = toLocal(g.frame, res.wi);
    const float F = fresnelDielectric(cosThetaOH, g.eta);
    const float G = ggxG2(g.mf, wo0, wi0, cosThetaOH, cosThetaIH);
    const float Ei = mfAlbedo(g.tables, cosThetaI, g.roughness);
    const float fms = g.thinFmsScale * ((1.f - Eo) * (1.f - Ei) * (1.f / (OC_PI * (1.f - g.thinEavg))) * cosThetaI);
    res.type = SC_GLOSSY_REFLECTION;
    res.pdf = whPdf * (1.f / (4.f * fabsf(cosThetaOH))) * R;
    res.weight = make_vec3f(F * G * rcpSafe(ggxG1(g.mf, wo0, cosThetaOH)) * (1.f / R)) + make_vec3f(fms * (1.f / res.pdf));
  } else {
    const float length = 1.f / refractCos(cosThetaO, g.eta);
    res.type = SC_SPECULAR_TRANSMISSION;
    res.wi = neg(wo);
    res.pdf = INFINITY;
    res.weight = expf3(g.attenuation * length);
  }
  return res;
}
__device__ __forceinline__ EvalRes PrGlass_eval(const PrGlass &g, vec3f wo, vec3f wi)
{
  EvalRes res; res.value = make_vec3f(0.f); res.pdf = 0.f;
  if (g.kind == 4) return ThinMicrofacetDielectric_eval(g, wo, wi);
  if (g.kind != 2) return res; // Transmission, Dielectric, ThinDielectric: delta lobes
  const float cosThetaO = dot(wo, g.frame.vz);
  if (cosThetaO <= 0.f) return res;
  const float cosThetaI = dot(wi, g.frame.vz);
  const float Eo = mfDielAlbedo(g.tables, cosThetaO, g.eta, g.roughness);
  const EvalRes fms = MicrofacetDielectric_fms(g, Eo, cosThetaI, cosThetaI > 0.f);
  const EvalRes single = MicrofacetDielectric_evalSingle(g, wo, wi);
  res.pdf = Eo * single.pdf + (1.f - Eo) * fms.pdf;
  res.value = single.value + fms.value;
  return res;
}
__device__ __forceinline__ SampleRes PrGlass_sample(const PrGlass &g, vec3f wo, vec2f s, float ss)
{
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  const vec3f N = g.frame.vz;
  if (g.kind == 0) { res.wi = neg(wo); res.pdf = INFINITY; res.type = SC_SPECULAR_TRANSMISSION; res.weight = make_vec3f(1.f); return res; }
  if (g.kind == 4) return ThinMicrofacetDielectric_sample(g, wo, s, ss);
  if (g.kind == 3) { // ThinDielectric_sample (ThinDielectric.ih:25-53)
    const float cosThetaO = fmaxf(dot(wo, N), 0.f);
    float cosThetaT;
    const float F = fresnelDielectricEx(cosThetaO, cosThetaT, g.eta);
    res.pdf = INFINITY;
    if (ss <= F) { res.wi = reflectDir(wo, N, cosThetaO); res.type = SC_SPECULAR_REFLECTION; res.weight = make_vec3f(1.f); }
    else { res.wi = neg(wo); res.type = SC_SPECULAR_TRANSMISSION; res.weight = expf3(g.attenuation * (1.f / cosThetaT)); }
    return res;
  }
  if (g.kind == 1) { // Dielectric_sample (Dielectric.ih:24-50)
    const float cosThetaO = fmaxf(dot(wo, N), 0.f);
    float cosThetaT;
    const float F = fresnelDielectricEx(cosThetaO, cosThetaT, g.eta);
    res.pdf = INFINITY;
    res.weight = make_vec3f(1.f);
    if (ss <= F) { res.wi = reflectDir(wo, N, cosThetaO); res.type = SC_SPECULAR_REFLECTION; }
    else { res.wi = refractDir(wo, N, cosThetaO, cosThetaT, g.eta); res.type = SC_SPECULAR_TRANSMISSION; }
    return res;
  }
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const float Eo = mfDielAlbedo(g.tables, cosThetaO, g.eta, g.roughness);
  const float singlePickProb = Eo;
  bool doReflection;
  if (ss <= singlePickProb) {
    const vec3f wo0 = toLocal(g.frame, wo);
    float whPdf;
    const vec3f wh = g.frame * ggxSampleVisible(g.mf, wo0, whPdf, s);
    const float cosThetaOH = dot(wo, wh);
    float cosThetaTH;
    const float F = fresnelDielectricEx(cosThetaOH, cosThetaTH, g.eta);
    const float ss1 = ss * (1.f / singlePickProb);
    doReflection = ss1 <= F;
    if (doReflection) { res.type = SC_GLOSSY_REFLECTION; res.wi = reflectDir(wo, wh, cosThetaOH); }
    else { res.type = SC_GLOSSY_TRANSMISSION; res.wi = refractDir(wo, wh, cosThetaOH, cosThetaTH, g.eta); }
  } else {
    const float ss1 = (ss - singlePickProb) * (1.f / (1.f - singlePickProb));
    doReflection = ss1 <= g.fmsRatio;
    res.wi = g.frame * cosineSampleHemisphere(s);
    if (doReflection) res.type = SC_DIFFUSE_REFLECTION;
    else { res.type = SC_DIFFUSE_TRANSMISSION; res.wi = neg(res.wi); }
  }
  const float cosThetaI = dot(res.wi, N);
  if ((doReflection && cosThetaI <= 0.f) || (!doReflection && cosThetaI >= 0.f)) {
    res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
    return res;
  }
  const EvalRes fms = MicrofacetDielectric_fms(g, Eo, cosThetaI, doReflection);
  const EvalRes single = MicrofacetDielectric_evalSingle(g, wo, res.wi);
  res.pdf = singlePickProb * single.pdf + (1.f - singlePickProb) * fms.pdf;
  res.weight = (single.value + fms.value) * (1.f / res.pdf);
  return res;
}

struct PrPlastic
{
  linear3f frame;
  int layerKind; // 0 passthrough (specular == 0), 1 DielectricLayer, 2 MicrofacetDielectricLayer
  LayerParams layer;
  PrDiffuse diffuse;
  vec3f albedo;
  uint32_t scatteringType;
  __device__ __forceinline__ EvalRes eval(vec3f wo, vec3f wi) const
  {
    if (layerKind == 0 || layer.weight < PR_EPS) return diffuse.eval(wo, wi);
    if (layerKind == 2) return MicrofacetDielectricLayer_eval(layer, frame, diffuse, wo, wi);
    return DielectricLayer_eval(layer, frame.vz, diffuse, wo, wi);
  }
  __device__ __forceinline__ SampleRes sample(vec3f wo, vec2f s, float ss) const
  {
    if (layerKind == 0 || layer.weight < PR_EPS) return diffuse.sample(wo, s, ss);
    if (layerKind == 2) return MicrofacetDielectricLayer_sample(layer, frame, diffuse, wo, s, ss);
    return DielectricLayer_sample(layer, frame.vz, diffuse, wo, s, ss);
  }
};

// the 3-child base MultiBSDF; albedo / scatteringType are read by the coat layered over it
struct PrBase
{
  linear3f frame;
  float w[3], imp[3], importanceSum;
  int metalKind; // 0 none, 1 Conductor, 2 MicrofacetConductor
  ConductorBSDF metal;
  bool hasGlass, hasPlastic;
  PrGlass glass;
  PrPlastic plastic;
  vec3f albedo;
  uint32_t scatteringType;

  __device__ __forceinline__ EvalRes childEval(int id, vec3f wo, vec3f wi) const
  {
    if (id == 0) { if (metalKind == 2) return MicrofacetConductor_eval(metal, wo, wi); EvalRes z; z.value = make_vec3f(0.f); z.pdf = 0.f; return z; }
    if (id == 1) return PrGlass_eval(glass, wo, wi);
    return plastic.eval(wo, wi);
  }
  __device__ __forceinline__ void evalChildren(vec3f wo, vec3f wi, int skip, EvalRes &e) const
  {
#pragma unroll
    for (int id = 0; id < 3; id++)
      if (imp[id] > 0.f && skip != id) {
        EvalRes c = childEval(id, wo, wi);
        c.value = c.value * w[id];
        e.value = e.value + c.value;
        e.pdf += c.pdf * imp[id];
      }
  }
  __device__ EvalRes eval(vec3f wo, vec3f wi) const
  {
    EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
    evalChildren(wo, wi, -1, e);
    e.pdf *= 1.f / importanceSum;
    return e;
  }
  __device__ SampleRes sample(vec3f wo, vec2f s, float ss) const
  {
    SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
    if (importanceSum == 0.f) return res;
    const float x = ss * importanceSum;
    int choice = -1;
    float prefix = 0.f;
#pragma unroll
    for (int id = 0; id < 3; id++)
      if (choice < 0 && x < (prefix += imp[id])) {
        choice = id;
        const float ssc = (x + imp[id] - prefix) * (1.f / imp[id]);
        if (id == 0) res = metalKind == 2 ? MicrofacetConductor_sample(metal, wo, s) : Conductor_sample(metal, wo);
        else if (id == 1) res = PrGlass_sample(glass, wo, s, ssc);
        else res = plastic.sample(wo, s, ssc);
        res.weight = res.weight * w[id];
      }
    if (choice < 0 || (res.weight.x == 0.f && res.weight.y == 0.f && res.weight.z == 0.f) || res.pdf == 0.0f) {
      res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
      return res;
    }
    if (res.type & SC_SPECULAR) { res.weight = res.weight * (importanceSum * (1.f / imp[choice])); return res; }
    EvalRes e;
    e.value = res.weight * res.pdf;
    res.pdf *= imp[choice];
    e.pdf = res.pdf;
    evalChildren(wo, res.wi, choice, e);
    res.pdf = e.pdf * (1.f / importanceSum);
    res.weight = e.value * (1.f / res.pdf);
    return res;
  }
};

struct PrBSDF
{
  linear3f frame;
  float opacity;     // weight and importance of the coated base in the root MultiBSDF
  float transparency; // importance of the cut-out Transmission child (its weight is 1), 0 = absent
  bool hasBase;
  int coatKind; // 0 passthrough, 1 DielectricLayer, 2 MicrofacetDielectricLayer
  LayerParams coat;
  linear3f coatFrame;  // of the coat and the sheen layer (the coat's normal map)
  vec3f rootN;         // normal of the root MultiBSDF's frame = the first child added (MultiBSDF.ih:35-36): coat frame, else base frame
  float sheenWeight;   // 0: passthrough (MicrofacetSheenLayer.ih:14-47)
  vec3f sheenR;
  float sheenAlpha, sheenRoughness;
  const float *tables;
  PrBase base;
  vec3f albedo;
  uint32_t scatteringType;
  float importanceSum;
};

__device__ __forceinline__ vec3f clamp01(vec3f v) { return make_vec3f(clampf(v.x, 0.f, 1.f), clampf(v.y, 0.f, 1.f), clampf(v.z, 0.f, 1.f)); }
__device__ __forceinline__ vec3f rotateAxis(vec3f v, vec3f axis, float theta) // rkcommon rotate (Rodrigues), restated
{
  float s, c;
  sincosf(theta, &s, &c);
  return v * c + cross(axis, v) * s + axis * dot(axis, v) * (1.f - c);
}

// factor * get1f / get3f(map, dg, 1) for every factor that has a map (Principled.ispc, throughout); baseAlpha = the alpha of the
// base colour map's sample when that texture has an alpha channel (it multiplies the opacity), cutout = valid(opacityMap) ||
// hasAlpha(baseColorMap)
static __device__ __noinline__ void prApplyMaps(OcPrincipled &v, const OcPrincipledMaps &mp, const DifferentialGeometry &dg, float &baseAlpha, bool &cutout)
{
  auto has = [&](int i) { return mp.map[i].tex != nullptr; };
  auto get = [&](int i) { return texParamGet4f(mp.map[i], dg.st, dg.pixelFootprint); };
  auto mul3 = [&](OcVec3f &f, int i) { if (has(i)) { const vec4f t = get(i); f.x *= t.x; f.y *= t.y; f.z *= t.z; } };
  auto mul1 = [&](float &f, int i) { if (has(i)) f *= get(i).x; };
  if (has(OC_PRMAP_BASECOLOR)) {
    const vec4f t = get(OC_PRMAP_BASECOLOR);
    v.baseColor.x *= t.x; v.baseColor.y *= t.y; v.baseColor.z *= t.z;
    if (mp.map[OC_PRMAP_BASECOLOR].tex->hasAlpha) { baseAlpha = t.w; cutout = true; }
  }
  if (has(OC_PRMAP_OPACITY)) cutout = true;
  mul3(v.edgeColor, OC_PRMAP_EDGECOLOR); mul1(v.metallic, OC_PRMAP_METALLIC); mul1(v.diffuse, OC_PRMAP_DIFFUSE);
  mul1(v.specular, OC_PRMAP_SPECULAR); mul1(v.ior, OC_PRMAP_IOR); mul1(v.transmission, OC_PRMAP_TRANSMISSION);
  mul3(v.transmissionColor, OC_PRMAP_TRANSMISSIONCOLOR); mul1(v.transmissionDepth, OC_PRMAP_TRANSMISSIONDEPTH);
  mul1(v.roughness, OC_PRMAP_ROUGHNESS); mul1(v.anisotropy, OC_PRMAP_ANISOTROPY); mul1(v.rotation, OC_PRMAP_ROTATION);
  mul1(v.coat, OC_PRMAP_COAT); mul1(v.coatIor, OC_PRMAP_COATIOR); mul3(v.coatColor, OC_PRMAP_COATCOLOR);
  mul1(v.coatThickness, OC_PRMAP_COATTHICKNESS); mul1(v.coatRoughness, OC_PRMAP_COATROUGHNESS); mul1(v.sheen, OC_PRMAP_SHEEN);
  mul3(v.sheenColor, OC_PRMAP_SHEENCOLOR); mul1(v.sheenTint, OC_PRMAP_SHEENTINT); mul1(v.sheenRoughness, OC_PRMAP_SHEENROUGHNESS);
  mul1(v.opacity, OC_PRMAP_OPACITY); mul1(v.backlight, OC_PRMAP_BACKLIGHT); mul1(v.thickness, OC_PRMAP_THICKNESS);
}
// makeShadingFrame(dg, normalMap, normalRot, normalScale) (render/Material.ih:59-81)
__device__ __forceinline__ linear3f prNormalMapFrame(const DifferentialGeometry &dg, const OcTextureParam &map, const float *rot, float scale)
{
  linear3f f; f.vx = normalize(dg.dPds); f.vy = normalize(dg.dPdt); f.vz = dg.Ns;
  const vec4f t = texParamGet4f(map, dg.st, dg.pixelFootprint);
  vec3f localN = (make_vec3f(t.x, t.y, t.z) * 2.f - make_vec3f(1.f)) * make_vec3f(scale, scale, 1.f);
  const vec2f rotN = make_vec2f(rot[0] * localN.x + rot[2] * localN.y, rot[1] * localN.x + rot[3] * localN.y);
  localN.x = rotN.x; localN.y = rotN.y;
  const vec3f N = normalize(f * localN);
  vec3f V = cross(N, f.vx);
  if (dot(V, V) > 0.f) {
    V = normalize(V);
    linear3f r; r.vx = cross(V, N); r.vy = V; r.vz = N; return r;
  }
  return frame(N);
}

// Principled_getBSDF (Principled.ispc:84-428) for solid materials without sheen and without texture maps; mediumIsOutside =
// eq(currentMedium, outsideMedium)
__device__ __forceinline__ void Principled_getBSDF(PrBSDF &b, const OcPrincipled &m, const float *tables, const DifferentialGeometry &dg, bool mediumIsOutside)
{
  OcPrincipled v = m; // factor * map
  float baseAlpha = 1.f;
  bool cutout = false;
  const OcPrincipledMaps *mp = m.maps;
  if (mp) prApplyMaps(v, *mp, dg, baseAlpha, cutout);
  b.frame = (mp && mp->map[OC_PRMAP_NORMAL].tex) ? prNormalMapFrame(dg, mp->map[OC_PRMAP_NORMAL], mp->normalRot, mp->normalScale) : makeShadingFrame(dg);
  const float rotation = clampf(v.rotation, 0.f, 1.f);
  const float theta = 2.f * OC_PI * rotation;
  if (theta > 0.f) {
    b.frame.vx = rotateAxis(b.frame.vx, b.frame.vz, theta);
    b.frame.vy = cross(b.frame.vz, b.frame.vx);
  }
  const float opacity = clampf(v.opacity, 0.f, 1.f) * baseAlpha;
  b.coatFrame = b.frame;
  b.rootN = b.frame.vz;
  b.albedo = make_vec3f(0.f); b.scatteringType = 0; b.importanceSum = 0.f;
  b.hasBase = false; b.coatKind = 0; b.opacity = 0.f; b.transparency = 0.f;
  if (m.opacity > PR_EPS) {
    linear3f baseFrame = b.frame;
    if (mp && mp->map[OC_PRMAP_BASENORMAL].tex) {
      baseFrame = prNormalMapFrame(dg, mp->map[OC_PRMAP_BASENORMAL], mp->baseNormalRot, mp->baseNormalScale);
      if (theta > 0.f) {
        baseFrame.vx = rotateAxis(baseFrame.vx, baseFrame.vz, theta);
        baseFrame.vy = cross(baseFrame.vz, baseFrame.vx);
      }
    }
    PrBase &B = b.base;
    B.frame = baseFrame;
    B.albedo = make_vec3f(0.f); B.scatteringType = 0; B.importanceSum = 0.f;
    B.w[0] = B.w[1] = B.w[2] = 0.f; B.imp[0] = B.imp[1] = B.imp[2] = 0.f;
    B.metalKind = 0; B.hasGlass = B.hasPlastic = false;
    const vec3f baseColor = clamp01(toV(v.baseColor) * make_vec3f(dg.color));
    const float specular = clampf(v.specular, 0.f, 1.f), metallic = clampf(v.metallic, 0.f, 1.f);
    const float roughness = clampf(v.roughness, 0.f, 1.f), anisotropy = clampf(v.anisotropy, 0.f, 1.f);
    const bool fromOutside = m.thin ? true : mediumIsOutside;
    const float dielectric = 1.f - metallic;
    if (dielectric > PR_EPS) {
      const float transmission = clampf(v.transmission, 0.f, 1.f);
      float ior = v.ior;
      if (ior < 1.f) ior = 1.f / ior;
      float eta = fromOutside ? m.outsideIor * (1.f / ior) : ior * (1.f / m.outsideIor);
      eta = clampf(eta, 1.f / 3.f, 3.f);
      const float plastic = dielectric * (1.f - transmission);
      if (plastic > PR_EPS) {
        PrPlastic &P = B.plastic;
        P.frame = baseFrame;
        const float diffuse = clampf(v.diffuse, 0.f, 1.f);
        const vec3f diffuseColor = baseColor * diffuse;
        PrDiffuse &D = P.diffuse;
        D.frame = baseFrame; D.R = diffuseColor; D.roughness = roughness; D.orenNayar = !(m.roughness < PR_EPS);
        const float backlight = m.thin ? clampf(v.backlight, 0.f, 2.f) : 0.f;
        const float diffuseTransmission = backlight * 0.5f;
        const float diffuseReflection = 1.f - diffuseTransmission;
        D.weight = D.importance = D.weightT = D.importanceT = 0.f; D.importanceSum = 0.f;
        D.albedo = make_vec3f(0.f); D.scatteringType = 0;
        if (diffuseReflection > PR_EPS) {
          D.weight = D.importance = diffuseReflection; D.importanceSum += diffuseReflection;
          D.albedo = D.albedo + diffuseColor * diffuseReflection; D.scatteringType |= SC_DIFFUSE_REFLECTION;
        }
        if (diffuseTransmission > PR_EPS) {
          D.weightT = D.importanceT = diffuseTransmission; D.importanceSum += diffuseTransmission;
          D.albedo = D.albedo + diffuseColor * diffuseTransmission; D.scatteringType |= SC_DIFFUSE_TRANSMISSION;
        }
        if (m.specular > PR_EPS) {
          if (m.roughness < PR_EPS) { P.layerKind = 1; DielectricLayer_init(P.layer, D.albedo, D.scatteringType, eta, make_vec3f(1.f), 1.f, specular); }
          else { P.layerKind = 2; MicrofacetDielectricLayer_init(P.layer, tables, D.albedo, D.scatteringType, eta, make_vec3f(1.f), 1.f, roughness, anisotropy, specular); }
          P.albedo = P.layer.albedo; P.scatteringType = P.layer.scatteringType;
        } else {
          P.layerKind = 0; P.layer.weight = 0.f;
          P.albedo = D.albedo; P.scatteringType = D.scatteringType;
        }
        B.hasPlastic = true;
        B.w[2] = plastic; B.imp[2] = plastic * fmaxf(diffuse, specular);
        B.albedo = B.albedo + P.albedo * plastic; B.scatteringType |= P.scatteringType; B.importanceSum += B.imp[2];
      }
      const float glass = dielectric * transmission * specular;
      if (glass > PR_EPS) {
        PrGlass &G = B.glass;
        G.frame = baseFrame;
        if (fabsf(eta - 1.f) > PR_EPS) {
          if (!m.thin) {
            if (m.roughness < PR_EPS) { G.kind = 1; G.eta = eta; G.albedo = make_vec3f(1.0f); G.scatteringType = SC_SPECULAR; }
            else PrGlass_initMicrofacet(G, tables, eta, roughness, anisotropy);
          } else {
            const vec3f transmissionColor = clamp01(toV(v.transmissionColor));
            const float transmissionDepth = fmaxf(v.transmissionDepth, PR_EPS);
            const float thickness = fmaxf(v.thickness, 0.f);
            const vec3f attenuation = logf3(transmissionColor) / transmissionDepth * thickness;
            if (m.roughness < PR_EPS) { G.kind = 3; G.eta = eta; G.attenuation = attenuation; G.albedo = expf3(attenuation); G.scatteringType = SC_SPECULAR; }
            else PrGlass_initThinMicrofacet(G, tables, eta, roughness, anisotropy, attenuation);
          }
        } else { G.kind = 0; G.eta = 1.f; G.albedo = make_vec3f(1.f); G.scatteringType = SC_SPECULAR_TRANSMISSION; }
        B.hasGlass = true;
        B.w[1] = glass; B.imp[1] = glass;
        B.albedo = B.albedo + G.albedo * glass; B.scatteringType |= G.scatteringType; B.importanceSum += glass;
      }
    }
    const float conductor = metallic * (m.specularMetallic ? specular : 1.0f);
    if (conductor > PR_EPS) {
      const vec3f edgeColor = clamp01(toV(v.edgeColor));
      // FresnelConductorArtistic_create (Fresnel.ih:264-284)
      const vec3f r = vmin(baseColor, make_vec3f(0.99f)), g = edgeColor;
      const vec3f one = make_vec3f(1.f);
      const vec3f sq = make_vec3f(sqrtf(r.x), sqrtf(r.y), sqrtf(r.z));
      const vec3f n_min = (one - r) * rcp3(one + r);
      const vec3f n_max = (one + sq) * rcp3(one - sq);
      const vec3f n = g * n_min + (one - g) * n_max;
      const vec3f k2 = (sqr3(n + one) * r - sqr3(n - one)) * rcp3(one - r);
      ConductorBSDF &M = B.metal;
      M.frame = baseFrame; M.tables = tables; M.schlick = false;
      M.eta = n;
      M.k = make_vec3f(sqrtf(fmaxf(k2.x, 0.f)), sqrtf(fmaxf(k2.y, 0.f)), sqrtf(fmaxf(k2.z, 0.f))); // sqrt_safe
      // FresnelConductorArtistic_evalAvg (:248-262)
      const vec3f r2 = r * r, r3 = r2 * r, g2 = g * g, g3 = g2 * g;
      const vec3f Favg = make_vec3f(0.087237f) + 0.0230685f * g - 0.0864902f * g2 + 0.0774594f * g3 + 0.782654f * r - 0.136432f * r2 + 0.278708f * r3
          + 0.19744f * g * r + 0.0360605f * g2 * r - 0.2586f * g * r2;
      vec3f metalAlbedo;
      if (m.roughness < PR_EPS) { B.metalKind = 1; metalAlbedo = Favg; }
      else {
        B.metalKind = 2;
        M.Eavg = mfAlbedoAvg(tables, roughness);
        M.fmsScale = sqr3(Favg) * M.Eavg / (one - Favg * (1.f - M.Eavg));
        M.mf = makeGGX(roughness, anisotropy);
        M.roughness = roughness;
        metalAlbedo = Favg * M.Eavg;
      }
      B.w[0] = conductor; B.imp[0] = conductor;
      B.albedo = B.albedo + metalAlbedo * conductor;
      B.scatteringType |= (B.metalKind == 1 ? SC_SPECULAR_REFLECTION : SC_GLOSSY_REFLECTION);
      B.importanceSum += conductor;
    }
    // clear coat
    vec3f topAlbedo = B.albedo;
    uint32_t topType = B.scatteringType;
    if (m.coat > PR_EPS) {
      const float coat = clampf(v.coat, 0.f, 1.f);
      float coatIor = v.coatIor;
      if (coatIor < 1.f) coatIor = 1.f / coatIor;
      float coatEta = fromOutside ? m.outsideIor * (1.f / coatIor) : coatIor * (1.f / m.outsideIor);
      coatEta = clampf(coatEta, 1.f / 3.f, 3.f);
      const vec3f coatColor = clamp01(toV(v.coatColor));
      const float coatThickness = fmaxf(v.coatThickness, 0.f);
      if (mp && mp->map[OC_PRMAP_COATNORMAL].tex) b.coatFrame = prNormalMapFrame(dg, mp->map[OC_PRMAP_COATNORMAL], mp->coatNormalRot, mp->coatNormalScale);
      if (m.coatRoughness < PR_EPS || fabsf(coatEta - 1.f) < PR_EPS) {
        b.coatKind = 1;
        DielectricLayer_init(b.coat, B.albedo, B.scatteringType, coatEta, coatColor, coatThickness, coat);
      } else {
        b.coatKind = 2;
        MicrofacetDielectricLayer_init(b.coat, tables, B.albedo, B.scatteringType, coatEta, coatColor, coatThickness, clampf(v.coatRoughness, 0.f, 1.f), 0.f, coat);
      }
      topAlbedo = b.coat.albedo; topType = b.coat.scatteringType;
    } else
      b.coat.weight = 0.f;
    // sheen (Principled.ispc:392-413)
    b.sheenWeight = 0.f; b.tables = tables;
    if (m.sheen > PR_EPS) {
      const float sheen = clampf(v.sheen, 0.f, 1.f);
      vec3f sheenColor = clamp01(toV(v.sheenColor));
      const float sheenTint = clampf(v.sheenTint, 0.f, 1.f);
      sheenColor = lerpv(sheenTint, sheenColor, baseColor);
      const float sheenRoughness = clampf(v.sheenRoughness, 0.f, 1.f);
      b.sheenWeight = sheen; b.sheenR = sheenColor; b.sheenRoughness = sheenRoughness;
      b.sheenAlpha = fmaxf(sqr(sheenRoughness), 0.001f);
      topAlbedo = lerpv(sheen * 0.1f, topAlbedo, sheenColor);
      topType = SC_DIFFUSE_REFLECTION | topType;
    }
    b.hasBase = true;
    b.rootN = (m.coat > PR_EPS) ? b.coatFrame.vz : baseFrame.vz;
    b.opacity = opacity;
    b.albedo = b.albedo + topAlbedo * opacity; b.scatteringType |= topType; b.importanceSum += opacity;
  }
  if (1.f - m.opacity > PR_EPS || cutout) {
    const float transparency = 1.f - opacity;
    b.transparency = transparency;
    b.albedo = b.albedo + make_vec3f(transparency) * 1.f; b.scatteringType |= SC_SPECULAR_TRANSMISSION; b.importanceSum += transparency;
  }
}

// ClearCoatBSDF_eval / _sample (Principled.ispc:900-960): the coat over the base MultiBSDF
__device__ __forceinline__ EvalRes PrCoat_eval(const PrBSDF &b, vec3f wo, vec3f wi)
{
  if (b.coatKind == 0 || b.coat.weight < PR_EPS) return b.base.eval(wo, wi);
  if (b.coatKind == 2) return MicrofacetDielectricLayer_eval(b.coat, b.coatFrame, b.base, wo, wi);
  return DielectricLayer_eval(b.coat, b.coatFrame.vz, b.base, wo, wi);
}
__device__ __forceinline__ SampleRes PrCoat_sample(const PrBSDF &b, vec3f wo, vec2f s, float ss)
{
  if (b.coatKind == 0 || b.coat.weight < PR_EPS) return b.base.sample(wo, s, ss);
  if (b.coatKind == 2) return MicrofacetDielectricLayer_sample(b.coat, b.coatFrame, b.base, wo, s, ss);
  return DielectricLayer_sample(b.coat, b.coatFrame.vz, b.base, wo, s, ss);
}
// SheenLayerBSDF_eval / _sample (Principled.ispc:990-1019) = MICROFACETSHEENLAYER_EVAL / _SAMPLE (MicrofacetSheenLayer.ih:82-222)
__device__ __forceinline__ EvalRes PrSheen_eval(const PrBSDF &b, vec3f wo, vec3f wi)
{
  if (b.sheenWeight < PR_EPS) return PrCoat_eval(b, wo, wi);
  EvalRes zero; zero.value = make_vec3f(0.f); zero.pdf = 0.f;
  const vec3f N = b.coatFrame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return zero;
  const float cosThetaI = dot(wi, N);
  EvalRes substrate = PrCoat_eval(b, wo, wi);
  const float Ro = mfSheenAlbedo(b.tables, cosThetaO, b.sheenRoughness) * b.sheenWeight;
  const float Ri = mfSheenAlbedo(b.tables, fabsf(cosThetaI), b.sheenRoughness) * b.sheenWeight;
  const float T = fminf(1.f - Ro, 1.f - Ri);
  substrate.value = substrate.value * T;
  const float coatPickProb = Ro, substratePickProb = 1.f - coatPickProb;
  if (cosThetaI > 0.f) {
    const vec3f wh = normalize(wo + wi);
    const float cosThetaH = dot(wh, N), cosThetaOH = dot(wo, wh), cosThetaIH = dot(wi, wh);
    const float D = sheenD(b.sheenAlpha, cosThetaH);
    const float G = sheenG2(b.sheenAlpha, cosThetaO, cosThetaI, cosThetaOH, cosThetaIH);
    const vec3f coatValue = b.sheenR * (D * G * (1.f / (4.f * cosThetaO)) * b.sheenWeight);
    substrate.pdf = coatPickProb * OC_ONE_OVER_TWO_PI + substratePickProb * substrate.pdf;
    substrate.value = coatValue + substrate.value;
  } else
    substrate.pdf *= substratePickProb;
  return substrate;
}
__device__ __forceinline__ SampleRes PrSheen_sample(const PrBSDF &b, vec3f wo, vec2f s, float ss)
{
  if (b.sheenWeight < PR_EPS) return PrCoat_sample(b, wo, s, ss);
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  const vec3f N = b.coatFrame.vz;
  const float cosThetaO = dot(wo, N);
  if (cosThetaO <= 0.f) return res;
  const float Ro = mfSheenAlbedo(b.tables, cosThetaO, b.sheenRoughness) * b.sheenWeight;
  EvalRes substrate;
  const float coatPickProb = Ro, substratePickProb = 1.f - coatPickProb;
  if (ss < coatPickProb) {
    res.type = SC_DIFFUSE_REFLECTION;
    res.wi = b.coatFrame * uniformSampleHemisphere(s);
    substrate = PrCoat_eval(b, wo, res.wi);
  } else {
    const float ss1 = (ss - coatPickProb) * (1.f / substratePickProb);
    res = PrCoat_sample(b, wo, s, ss1);
    if (reduce_max(res.weight) <= 0.f) return res;
    substrate.pdf = res.pdf;
    substrate.value = res.weight * ((res.type & SC_SPECULAR) ? 1.f : res.pdf);
  }
  const float cosThetaI = dot(res.wi, N);
  const float Ri = mfSheenAlbedo(b.tables, fabsf(cosThetaI), b.sheenRoughness) * b.sheenWeight;
  const float T = fminf(1.f - Ro, 1.f - Ri);
  substrate.value = substrate.value * T;
  if (res.type & SC_SPECULAR) res.weight = substrate.value * (1.f / substratePickProb);
  else if (cosThetaI <= 0.f) {
    res.pdf = substratePickProb * substrate.pdf;
    res.weight = substrate.value * (1.f / res.pdf);
  } else {
    const vec3f wh = normalize(wo + res.wi);
    const float cosThetaH = dot(wh, N), cosThetaOH = dot(wo, wh), cosThetaIH = dot(res.wi, wh);
    const float D = sheenD(b.sheenAlpha, cosThetaH);
    const float G = sheenG2(b.sheenAlpha, cosThetaO, cosThetaI, cosThetaOH, cosThetaIH);
    const vec3f coatValue = b.sheenR * (D * G * (1.f / (4.f * cosThetaO)) * b.sheenWeight);
    res.pdf = coatPickProb * OC_ONE_OVER_TWO_PI + substratePickProb * substrate.pdf;
    res.weight = (coatValue + substrate.value) * (1.f / res.pdf);
  }
  return res;
}
// Principled_BSDF_eval / _sample (Principled.ispc:1000-1040): the root MultiBSDF (child 0 = cut-out transmission)
static __device__ __noinline__ EvalRes Principled_eval(const PrBSDF &b, vec3f wo, vec3f wi)
{
  EvalRes e; e.value = make_vec3f(0.f); e.pdf = 0.f;
  if (b.hasBase && b.opacity > 0.f) {
    EvalRes c = PrSheen_eval(b, wo, wi);
    c.value = c.value * b.opacity;
    e.value = e.value + c.value;
    e.pdf += c.pdf * b.opacity;
  }
  e.pdf *= 1.f / b.importanceSum;
  return e;
}
static __device__ __noinline__ SampleRes Principled_sample(const PrBSDF &b, vec3f wo, vec2f s, float ss)
{
  SampleRes res; res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
  if (b.importanceSum == 0.f) return res;
  const float x = ss * b.importanceSum;
  int choice;
  float prefix = 0.f;
  if (x < (prefix += b.transparency)) {
    choice = 0;
    res.wi = neg(wo); res.pdf = INFINITY; res.type = SC_SPECULAR_TRANSMISSION; res.weight = make_vec3f(b.transparency);
    res.weight = res.weight * 1.f;
  } else if (x < (prefix += (b.hasBase ? b.opacity : 0.f))) {
    choice = 1;
    const float ssc = (x + b.opacity - prefix) * (1.f / b.opacity);
    res = PrSheen_sample(b, wo, s, ssc);
    res.weight = res.weight * b.opacity;
  } else return res;
  if ((res.weight.x == 0.f && res.weight.y == 0.f && res.weight.z == 0.f) || res.pdf == 0.0f) {
    res.weight = make_vec3f(0.f); res.wi = make_vec3f(0.f); res.pdf = 0.f; res.type = 0;
    return res;
  }
  const float choiceImp = choice == 0 ? b.transparency : b.opacity;
  if (res.type & SC_SPECULAR) { res.weight = res.weight * (b.importanceSum * (1.f / choiceImp)); return res; }
  EvalRes e;
  e.value = res.weight * res.pdf;
  res.pdf *= choiceImp;
  e.pdf = res.pdf;
  // (child 0 has no evaluation; child 1 is the one sampled)
  res.pdf = e.pdf * (1.f / b.importanceSum);
  res.weight = e.value * (1.f / res.pdf);
  return res;
}

// Principled_getTransparency (Principled.ispc:430-573), same scope
__device__ __forceinline__ vec3f Principled_getTransparency(const OcPrincipled &m, const float *tables, const DifferentialGeometry &dg, vec3f rayDir, bool mediumIsOutside)
{
  vec3f T = make_vec3f(0.f);
  OcPrincipled v = m; // factor * map
  float baseAlpha = 1.f;
  bool cutout = false;
  if (m.maps) prApplyMaps(v, *m.maps, dg, baseAlpha, cutout);
  const float opacity = clampf(v.opacity, 0.f, 1.f) * baseAlpha;
  const float transparency = 1.f - opacity;
  const float transmission = clampf(v.transmission, 0.f, 1.f);
  if (transparency <= PR_EPS && transmission <= PR_EPS) return T;
  if (m.opacity > PR_EPS) {
    const float specular = clampf(v.specular, 0.f, 1.f), metallic = clampf(v.metallic, 0.f, 1.f);
    const float glass = (1.f - metallic) * transmission * specular;
    if (glass > PR_EPS) {
      const bool fromOutside = m.thin ? true : mediumIsOutside;
      float ior = v.ior;
      if (ior < 1.f) ior = 1.f / ior;
      float eta = fromOutside ? m.outsideIor * (1.f / ior) : ior * (1.f / m.outsideIor);
      eta = clampf(eta, 1.f / 3.f, 3.f);
      const float cosThetaO = fmaxf(-dot(rayDir, dg.Ns), 0.f);
      if (fabsf(eta - 1.f) > PR_EPS) {
        if (!m.thin) {
          if (m.roughness < PR_EPS) T = make_vec3f(1.f - fresnelDielectric(cosThetaO, eta)); // Dielectric_getTransparency
        } else {
          const vec3f transmissionColor = clamp01(toV(v.transmissionColor));
          const float transmissionDepth = fmaxf(v.transmissionDepth, PR_EPS);
          const float thickness = fmaxf(v.thickness, 0.f);
          const vec3f attenuation = logf3(transmissionColor) / transmissionDepth * thickness;
          if (m.roughness < PR_EPS) { // ThinDielectric_getTransparency (ThinDielectric.ih:88-98)
            float cosThetaT;
            const float F = fresnelDielectricEx(cosThetaO, cosThetaT, eta);
            T = (1.f - F) * expf3(attenuation * (1.f / cosThetaT));
          } else { // ThinMicrofacetDielectric_getTransparency (ThinMicrofacetDielectric.ih:209-231)
            const float roughness = clampf(v.roughness, 0.f, 1.f);
            const float Eavg = mfAlbedoAvg(tables, roughness);
            const float Favg = fresnelDielectricAvg(eta);
            const float fmsScale = Favg * (1.f - Eavg) * (1.f / (1.f - Favg * Eavg));
            const float Eo = mfAlbedo(tables, cosThetaO, roughness);
            const float R = mfDielReflAlbedo(tables, cosThetaO, eta, roughness) + fmsScale * (1.f - Eo);
            T = (1.f - R) * expf3(attenuation * (1.f / refractCos(cosThetaO, eta)));
          }
        }
      } else
        T = make_vec3f(1.f);
      if (reduce_max(T) > 0.f) {
        if (m.coat > PR_EPS) {
          const float coat = clampf(v.coat, 0.f, 1.f);
          float coatIor = v.coatIor;
          if (coatIor < 1.f) coatIor = 1.f / coatIor;
          float coatEta = fromOutside ? m.outsideIor * (1.f / coatIor) : coatIor * (1.f / m.outsideIor);
          const float coatEtaClamped = clampf(coatEta, 1.f / 3.f, 3.f);
          if (fabsf(coatEtaClamped - 1.f) > PR_EPS) {
            const vec3f coatColor = clamp01(toV(v.coatColor));
            const float coatThickness = fmaxf(v.coatThickness, 0.f);
            if (m.coatRoughness < PR_EPS) { // DielectricLayer_getTransparency (DielectricLayer.ih:52-76), unclamped eta
              float e = coatEta;
              if (e < 1.f) e = 1.f / e;
              float cosThetaO1;
              const float Fo = fresnelDielectricEx(cosThetaO, cosThetaO1, e) * coat;
              vec3f value = lerpv(coat, make_vec3f(1.f), powv(coatColor, coatThickness * (1.f / cosThetaO1)));
              T = T * (value * (1.f - Fo));
            } else { // MicrofacetDielectricLayer_getTransparency (MicrofacetDielectricLayer.ih:113-146)
              float e = coatEta;
              if (e > 1.f) e = 1.f / e;
              const float cr = clampf(v.coatRoughness, 0.f, 1.f);
              vec3f value = lerpv(coat, make_vec3f(1.f), powv(coatColor, coatThickness * (1.f / refractCos(cosThetaO, e))));
              const float Eavg = mfAlbedoAvg(tables, cr);
              const float Favg = fresnelDielectricAvg(e);
              const float fmsScale = Favg * (1.f - Eavg) * (1.f / (1.f - Favg * Eavg));
              const float Eo = mfAlbedo(tables, cosThetaO, cr);
              const float Ro = (mfDielReflAlbedo(tables, cosThetaO, e, cr) + fmsScale * (1.f - Eo)) * coat;
              T = T * (value * (1.f - Ro));
            }
          }
        }
        if (m.sheen > PR_EPS) { // MicrofacetSheenLayer_getTransparency (MicrofacetSheenLayer.ih:66-78)
          const float sheen = clampf(v.sheen, 0.f, 1.f), sheenRoughness = clampf(v.sheenRoughness, 0.f, 1.f);
          T = T * make_vec3f(1.f - mfSheenAlbedo(tables, cosThetaO, sheenRoughness) * sheen);
        }
        T = T * glass;
      }
    }
  }
  return T * opacity + make_vec3f(transparency);
}

// ---- carPaint (materials/CarPaint.ispc:29-365): flakes (a conductor on a perturbed frame where the cellular noise of
// shaders/Flakes.ih says there is one) or a diffuse base, under a smooth / rough dielectric coat whose colour flip-flops with the
// viewing angle over flakes
__device__ __forceinline__ uint32_t cpNoiseHash(vec3f cell, uint32_t seed) // noiseHash(intbits(floor(v)), seed) (shaders/Noise.ih:10-24)
{
  uint32_t h = seed;
  h = murmur3_mix(h, __float_as_uint(floorf(cell.x)));
  h = murmur3_mix(h, __float_as_uint(floorf(cell.y)));
  h = murmur3_mix(h, __float_as_uint(floorf(cell.z)));
  h ^= 3 * 4;
  return murmur3_finalize(h);
}
__device__ __forceinline__ float cpLcgFloat(uint32_t &state) { state = state * 1664525u + 1013904223u; return to_float_unorm(state); }
// Flakes_eval (shaders/Flakes.ih:17-65): nearest jittered cell centre; a flake there with probability `density`, its normal a
// Beckmann sample (bsdfs/BeckmannDistribution.ih:18-41)
static __device__ __noinline__ vec3f Flakes_eval(float scale, float density, float spread, float jitter, vec3f P, int &mask)
{
  vec3f N = make_vec3f(0.f, 0.f, 1.f);
  mask = 0;
  if (density <= 0.f || spread <= 0.f) return N;
  const float alpha = fmaxf(sqr(spread), 0.001f);
  P = P * scale;
  const vec3f thisCell = make_vec3f(floorf(P.x) + 0.5f, floorf(P.y) + 0.5f, floorf(P.z) + 0.5f);
  float f1 = 1000.f;
  uint32_t cellRnd = 0;
  for (int i = -1; i <= 1; i++)
    for (int j = -1; j <= 1; j++)
      for (int k = -1; k <= 1; k++) {
        const vec3f testCell = thisCell + make_vec3f((float)i, (float)j, (float)k);
        uint32_t rnd = cpNoiseHash(testCell, 0x537e6612u);
        const float rx = cpLcgFloat(rnd), ry = cpLcgFloat(rnd), rz = cpLcgFloat(rnd);
        const vec3f pos = testCell + jitter * (make_vec3f(rx, ry, rz) - make_vec3f(0.5f));
        const vec3f offset = pos - P;
        const float dist = dot(offset, offset);
        if (dist < f1) { f1 = dist; cellRnd = rnd; }
      }
  if (cpLcgFloat(cellRnd) <= density) {
    const float sx = cpLcgFloat(cellRnd), sy = cpLcgFloat(cellRnd);
    const float alpha2 = sqr(alpha);
    const float tanTheta2 = -alpha2 * logf(1.f - sx);
    const float cosTheta = 1.f / sqrtf(1.f + tanTheta2);
    const float sinTheta = cos2sin(cosTheta);
    const float phi = 2.f * OC_PI * sy;
    float sinPhi, cosPhi;
    sincosf(phi, &sinPhi, &cosPhi);
    N = make_vec3f(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta);
    mask = 1;
  }
  return normalize(N);
}
// makeShadingFrame(dg, localN) (render/Material.ih:36-57)
__device__ __forceinline__ linear3f cpFrameFromLocalNormal(const DifferentialGeometry &dg, vec3f localN)
{
  const linear3f f = makeShadingFrame(dg);
  const vec3f N = normalize(f * localN);
  vec3f V = cross(N, f.vx);
  if (dot(V, V) > 0.f) {
    V = normalize(V);
    linear3f r; r.vx = cross(V, N); r.vy = V; r.vz = N; return r;
  }
  return frame(N);
}
struct CarPaintBSDF
{
  int flakeMask, condKind;   // condKind: 1 Conductor, 2 MicrofacetConductor
  ConductorBSDF cond;
  PrDiffuse diff;            // Lambert | OrenNayar (lobeEval / lobeSample)
  int coatKind;              // 0 passthrough, 1 DielectricLayer, 2 MicrofacetDielectricLayer
  LayerParams coat;
  linear3f coatFrame;
  vec3f albedo;              // of the substrate (what the coat macros read), then of the whole
  uint32_t scatteringType;
  __device__ __forceinline__ EvalRes eval(vec3f wo, vec3f wi) const // BaseBSDF_eval
  {
    if (flakeMask) { if (condKind == 2) return MicrofacetConductor_eval(cond, wo, wi); EvalRes z; z.value = make_vec3f(0.f); z.pdf = 0.f; return z; }
    return diff.lobeEval(wo, wi);
  }
  __device__ __forceinline__ SampleRes sample(vec3f wo, vec2f s, float) const // BaseBSDF_sample
  {
    if (flakeMask) return condKind == 2 ? MicrofacetConductor_sample(cond, wo, s) : Conductor_sample(cond, wo);
    return diff.lobeSample(wo, s);
  }
};
static __device__ __noinline__ void CarPaint_getBSDF(CarPaintBSDF &b, vec3f &rootN, vec3f &rootAlbedo, uint32_t &rootType, const OcCarPaint &m,
    const float *tables, const DifferentialGeometry &dg, vec3f rayDir)
{
  auto has = [&](int i) { return m.map[i].tex != nullptr; };
  auto get1 = [&](int i) { return has(i) ? texParamGet4f(m.map[i], dg.st, dg.pixelFootprint).x : 1.f; };
  auto get3 = [&](int i) { if (!has(i)) return make_vec3f(1.f); const vec4f t = texParamGet4f(m.map[i], dg.st, dg.pixelFootprint); return make_vec3f(t.x, t.y, t.z); };
  const linear3f frm = has(OC_CPMAP_NORMAL) ? prNormalMapFrame(dg, m.map[OC_CPMAP_NORMAL], m.normalRot, m.normalScale) : makeShadingFrame(dg);
  const float flakeDensity = clampf(m.flakeDensity * get1(OC_CPMAP_FLAKEDENSITY), 0.f, 1.f);
  int flakeMask = 0;
  linear3f flakeFrame = frm;
  if (flakeDensity > PR_EPS) {
    const float flakeScale = fmaxf(m.flakeScale * get1(OC_CPMAP_FLAKESCALE), 0.f);
    const float flakeSpread = fmaxf(m.flakeSpread * get1(OC_CPMAP_FLAKESPREAD), 0.f);
    const float flakeJitter = clampf(m.flakeJitter * get1(OC_CPMAP_FLAKEJITTER), 0.f, 1.f);
    const vec3f flakeN = Flakes_eval(flakeScale, flakeDensity, flakeSpread, flakeJitter, dg.P, flakeMask);
    if (flakeMask) {
      flakeFrame = cpFrameFromLocalNormal(dg, flakeN);
      ConductorBSDF &c = b.cond;
      c.frame = flakeFrame; c.tables = tables;
      vec3f Favg;
      if (m.useFlakeColor) {
        c.schlick = true;
        c.eta = clamp01(toV(m.flakeColor) * get3(OC_CPMAP_FLAKECOLOR)); c.k = make_vec3f(1.f);
        Favg = fresnelSchlickAvg(c.eta, c.k);
      } else { // aluminium; the RGB conductor has no average (Fresnel.ih:307-320)
        c.schlick = false;
        c.eta = make_vec3f(1.69700277f, 0.879832864f, 0.5301736f); c.k = make_vec3f(9.30200672f, 6.27604008f, 4.89433956f);
        Favg = make_vec3f(0.f);
      }
      if (m.flakeRoughness < PR_EPS) { b.condKind = 1; b.albedo = Favg; b.scatteringType = SC_SPECULAR_REFLECTION; }
      else {
        const float r = fmaxf(m.flakeRoughness * get1(OC_CPMAP_FLAKEROUGHNESS), 0.f);
        b.condKind = 2;
        c.Eavg = mfAlbedoAvg(tables, r);
        c.fmsScale = sqr3(Favg) * c.Eavg / (make_vec3f(1.f) - Favg * (1.f - c.Eavg));
        c.mf = makeGGX(r, 0.f);
        c.roughness = r;
        b.albedo = Favg * c.Eavg; b.scatteringType = SC_GLOSSY_REFLECTION;
      }
    }
  }
  rootN = flakeMask ? flakeFrame.vz : frm.vz;
  if (!flakeMask) {
    const vec3f baseColor = clamp01(toV(m.baseColor) * get3(OC_CPMAP_BASECOLOR) * make_vec3f(dg.color));
    PrDiffuse &D = b.diff;
    D.frame = frm; D.R = baseColor;
    D.orenNayar = !(m.roughness < PR_EPS);
    D.roughness = D.orenNayar ? fmaxf(m.roughness * get1(OC_CPMAP_ROUGHNESS), 0.f) : 0.f;
    b.albedo = baseColor; b.scatteringType = SC_DIFFUSE_REFLECTION;
  }
  b.flakeMask = flakeMask;
  b.coatKind = 0; b.coat.weight = 0.f; b.coatFrame = frm;
  rootAlbedo = b.albedo; rootType = b.scatteringType;
  if ((m.coat > PR_EPS) && (fabsf(m.coatIor - 1.f) > PR_EPS || has(OC_CPMAP_COATIOR))) {
    const float coat = fmaxf(m.coat * get1(OC_CPMAP_COAT), 0.f);
    float coatIor = m.coatIor * get1(OC_CPMAP_COATIOR);
    if (coatIor < 1.f) coatIor = 1.f / coatIor;
    coatIor = clampf(coatIor, 1.f, 3.f);
    const vec3f coatColor = clamp01(toV(m.coatColor) * get3(OC_CPMAP_COATCOLOR));
    vec3f coatFinalColor = coatColor;
    if (flakeMask) {
      const float flipflopFalloff = clampf(m.flipflopFalloff * get1(OC_CPMAP_FLIPFLOPFALLOFF), 0.f, 1.f);
      if (flipflopFalloff < 1.f - PR_EPS) {
        const vec3f flipflopColor = clamp01(toV(m.flipflopColor) * get3(OC_CPMAP_FLIPFLOPCOLOR));
        const float cosThetaO = fmaxf(-dot(rayDir, flakeFrame.vz), 0.f);
        const float weight = powf(1.f - cosThetaO, 1.f / (1.f - flipflopFalloff));
        coatFinalColor = lerpv(weight, coatColor, flipflopColor);
      }
    }
    const float coatThickness = fmaxf(m.coatThickness * get1(OC_CPMAP_COATTHICKNESS), 0.f);
    b.coatFrame = has(OC_CPMAP_COATNORMAL) ? prNormalMapFrame(dg, m.map[OC_CPMAP_COATNORMAL], m.coatNormalRot, m.coatNormalScale) : makeShadingFrame(dg);
    if (m.coatRoughness < PR_EPS) {
      b.coatKind = 1;
      DielectricLayer_init(b.coat, b.albedo, b.scatteringType, 1.f / coatIor, coatFinalColor, coatThickness, coat);
    } else {
      b.coatKind = 2;
      MicrofacetDielectricLayer_init(b.coat, tables, b.albedo, b.scatteringType, 1.f / coatIor, coatFinalColor, coatThickness,
          fmaxf(m.coatRoughness * get1(OC_CPMAP_COATROUGHNESS), 0.f), 0.f, coat);
    }
    rootAlbedo = b.coat.albedo; rootType = b.coat.scatteringType;
    rootN = b.coatFrame.vz;
  }
}
__device__ __forceinline__ EvalRes CarPaint_eval(const CarPaintBSDF &b, vec3f wo, vec3f wi)
{
  if (b.coatKind == 0 || b.coat.weight < PR_EPS) return b.eval(wo, wi);
  if (b.coatKind == 2) return MicrofacetDielectricLayer_eval(b.coat, b.coatFrame, b, wo, wi);
  return DielectricLayer_eval(b.coat, b.coatFrame.vz, b, wo, wi);
}
__device__ __forceinline__ SampleRes CarPaint_sample(const CarPaintBSDF &b, vec3f wo, vec2f s, float ss)
{
  if (b.coatKind == 0 || b.coat.weight < PR_EPS) return b.sample(wo, s, ss);
  if (b.coatKind == 2) return MicrofacetDielectricLayer_sample(b.coat, b.coatFrame, b, wo, s, ss);
  return DielectricLayer_sample(b.coat, b.coatFrame.vz, b, wo, s, ss);
}

} // namespace oc
