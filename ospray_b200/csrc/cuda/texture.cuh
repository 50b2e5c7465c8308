// 2D textures: texel formats, wrap modes, nearest / bilinear filtering, MIP level selection from the ray-cone footprint.
// reference: texture/Texture2D.ispc:13-158 (wrap_coord, nearest_coords, bilinear_coords, calcLambda, Texture2D_get),
// texture/Texture2DFormats.ispc:216-400 (getTexel_*), texture/TextureParam.ispc:10-22 (2D transform of dg.st),
// texture/Texture2D.ih:20-47 (clamp2edge, get1f / get4f of a bare Texture2D, footprint flt_min).
// sRGB decoding follows rkcommon's srgb_to_linear (un-vendored: x <= 0.04045 ? x / 12.92 : ((x + 0.055) / 1.055)^2.4).
#pragma once
#include <cuda_fp16.h>

namespace oc {

enum { // OSPTextureFormat (OSPEnums.h)
  TEXF_RGBA8 = 0, TEXF_SRGBA, TEXF_RGBA32F, TEXF_RGB8, TEXF_SRGB, TEXF_RGB32F, TEXF_R8, TEXF_R32F, TEXF_L8, TEXF_RA8, TEXF_LA8,
  TEXF_RGBA16, TEXF_RGB16, TEXF_RA16, TEXF_R16, TEXF_RA32F, TEXF_RGBA16F, TEXF_RGB16F, TEXF_RA16F, TEXF_R16F
};
enum { TEX_FILTER_LINEAR = 0, TEX_FILTER_NEAREST = 1 };
enum { TEX_WRAP_REPEAT = 0, TEX_WRAP_MIRRORED_REPEAT = 1, TEX_WRAP_CLAMP_TO_EDGE = 2 };

// srgb_to_linear of an 8-bit texel value: the 256 possible results as a table (tools/gen_srgb_table.py: the C library's powf, i.e.
// the values the reference's scalar form computes), instead of a ~150-instruction powf body per channel per texel
static __device__ const float texSrgbTable[256] = {
#include "srgb_table.inc"
};
__device__ __forceinline__ float texSrgb8ToLinear(uint32_t c) { return __ldg(texSrgbTable + c); }
__device__ __forceinline__ float texHalf(uint32_t bits) { return __half2float(__ushort_as_half((unsigned short)bits)); }

// Texture2DFormats.ispc getTexel_*
static __device__ __noinline__ vec4f texGetTexel(uint32_t format, const void *data, uint32_t idx)
{
  const float B2F = 1.f / 255.f, S2F = 1.f / 65535.f;
  const uint8_t *u8 = (const uint8_t *)data;
  const uint16_t *u16 = (const uint16_t *)data;
  const uint32_t *u32 = (const uint32_t *)data;
  const float *f32 = (const float *)data;
  switch (format) {
  case TEXF_RGBA8:
  case TEXF_SRGBA: {
    const uint32_t c = __ldg(u32 + idx);
    if (format == TEXF_RGBA8) return make_vec4f((float)(c & 0xff), (float)((c >> 8) & 0xff), (float)((c >> 16) & 0xff), (float)(c >> 24)) * B2F;
    return make_vec4f(texSrgb8ToLinear(c & 0xff), texSrgb8ToLinear((c >> 8) & 0xff), texSrgb8ToLinear((c >> 16) & 0xff), (float)(c >> 24) * B2F);
  }
  case TEXF_RGBA32F: { const float4 v = __ldg((const float4 *)data + idx); return make_vec4f(v.x, v.y, v.z, v.w); }
  case TEXF_RGB8:
  case TEXF_SRGB: {
    const uint32_t r = __ldg(u8 + 3 * idx), g = __ldg(u8 + 3 * idx + 1), b = __ldg(u8 + 3 * idx + 2);
    if (format == TEXF_RGB8) return make_vec4f(make_vec3f((float)r, (float)g, (float)b) * B2F, 1.f);
    return make_vec4f(texSrgb8ToLinear(r), texSrgb8ToLinear(g), texSrgb8ToLinear(b), 1.f);
  }
  case TEXF_RGB32F: return make_vec4f(__ldg(f32 + 3 * idx), __ldg(f32 + 3 * idx + 1), __ldg(f32 + 3 * idx + 2), 1.f);
  case TEXF_R8: return make_vec4f(__ldg(u8 + idx) * B2F, 0.f, 0.f, 1.f);
  case TEXF_R32F: return make_vec4f(__ldg(f32 + idx), 0.f, 0.f, 1.f);
  case TEXF_L8: { const float l = texSrgb8ToLinear(__ldg(u8 + idx)); return make_vec4f(l, l, l, 1.f); }
  case TEXF_RA8: { const uint32_t t = __ldg(u16 + idx); return make_vec4f((t & 0xff) * B2F, 0.f, 0.f, (t >> 8) * B2F); }
  case TEXF_LA8: {
    const uint32_t t = __ldg(u16 + idx);
    const float l = texSrgb8ToLinear(t & 0xff);
    return make_vec4f(l, l, l, (t >> 8) * B2F);
  }
  case TEXF_RGBA16: {
    const uint32_t c0 = __ldg(u32 + 2 * idx), c1 = __ldg(u32 + 2 * idx + 1);
    return make_vec4f((float)(c0 & 0xffff), (float)(c0 >> 16), (float)(c1 & 0xffff), (float)(c1 >> 16)) * S2F;
  }
  case TEXF_RGB16:
    return make_vec4f(make_vec3f((float)__ldg(u16 + 3 * idx), (float)__ldg(u16 + 3 * idx + 1), (float)__ldg(u16 + 3 * idx + 2)) * S2F, 1.f);
  case TEXF_RA16: { const uint32_t t = __ldg(u32 + idx); return make_vec4f((t & 0xffff) * S2F, 0.f, 0.f, (t >> 16) * S2F); }
  case TEXF_R16: return make_vec4f(__ldg(u16 + idx) * S2F, 0.f, 0.f, 1.f);
  case TEXF_RA32F: return make_vec4f(__ldg(f32 + 2 * idx), 0.f, 0.f, __ldg(f32 + 2 * idx + 1));
  case TEXF_RGBA16F: {
    const uint32_t c0 = __ldg(u32 + 2 * idx), c1 = __ldg(u32 + 2 * idx + 1);
    return make_vec4f(texHalf(c0 & 0xffff), texHalf(c0 >> 16), texHalf(c1 & 0xffff), texHalf(c1 >> 16));
  }
  case TEXF_RGB16F: return make_vec4f(texHalf(__ldg(u16 + 3 * idx)), texHalf(__ldg(u16 + 3 * idx + 1)), texHalf(__ldg(u16 + 3 * idx + 2)), 1.f);
  case TEXF_RA16F: { const uint32_t t = __ldg(u32 + idx); return make_vec4f(texHalf(t & 0xffff), 0.f, 0.f, texHalf(t >> 16)); }
  case TEXF_R16F: return make_vec4f(texHalf(__ldg(u16 + idx)), 0.f, 0.f, 1.f);
  default: return make_vec4f(0.f, 0.f, 0.f, 1.f);
  }
}

__device__ __forceinline__ float texFrac(float x) { return x - floorf(x); }
// Texture2D.ispc:14-40
__device__ __forceinline__ float texWrapCoord(uint32_t wrapMode, float halfTexel, float p)
{
  float tc = 0.f;
  switch (wrapMode) {
  case TEX_WRAP_REPEAT:
    tc = texFrac(p - halfTexel) + halfTexel;
    tc = tc < 0.f ? 0.f : tc; // max(tc, 0.f) in the ISPC sense: NaN -> 0
    if (!(tc >= 0.f)) tc = 0.f;
    break;
  case TEX_WRAP_MIRRORED_REPEAT:
    tc = 2.f * texFrac(p * 0.5f);
    if (tc >= 1.f) tc = 2.f - tc;
    tc = clampf(tc, halfTexel, 1.f - halfTexel);
    break;
  case TEX_WRAP_CLAMP_TO_EDGE: tc = clampf(p, halfTexel, 1.f - halfTexel); break;
  }
  return tc;
}
__device__ __forceinline__ int texLevelSize(int s, int level) { return max(s >> level, 1); }
// Texture2D_nearest_l / Texture2D_bilinear_l (Texture2D.ispc:42-131)
__device__ __forceinline__ vec4f texSampleLevel(const OcTexture2D &t, vec2f st, int level)
{
  const int sx = texLevelSize(t.sizeX, level), sy = texLevelSize(t.sizeY, level);
  const void *data = t.data[level];
  if (t.filter & TEX_FILTER_NEAREST) {
    const float tx = texWrapCoord(t.wrapS, 0.f, st.x), ty = texWrapCoord(t.wrapT, 0.f, st.y);
    const int ix = min((int)(tx * (float)sx), sx - 1), iy = min((int)(ty * (float)sy), sy - 1);
    return texGetTexel(t.format, data, (uint32_t)(iy * sx + ix));
  }
  const float fsx = (float)sx, fsy = (float)sy;
  const float hx = (1.f / fsx) * 0.5f, hy = (1.f / fsy) * 0.5f;
  float tx = texWrapCoord(t.wrapS, hx, st.x), ty = texWrapCoord(t.wrapT, hy, st.y);
  tx = (tx - hx) * fsx; ty = (ty - hy) * fsy;
  const float fx = texFrac(tx), fy = texFrac(ty);
  const int x0 = min((int)tx, sx - 1), y0 = min((int)ty, sy - 1);
  int x1 = x0 + 1, y1 = y0 + 1;
  if (x1 >= sx) x1 = 0; // border cases as REPEAT
  if (y1 >= sy) y1 = 0;
  const vec4f c00 = texGetTexel(t.format, data, (uint32_t)(y0 * sx + x0)), c01 = texGetTexel(t.format, data, (uint32_t)(y0 * sx + x1)),
              c10 = texGetTexel(t.format, data, (uint32_t)(y1 * sx + x0)), c11 = texGetTexel(t.format, data, (uint32_t)(y1 * sx + x1));
  // bilerp: lerp(frac.y, lerp(frac.x, c00, c01), lerp(frac.x, c10, c11)), lerp(f, a, b) = (1 - f) * a + f * b
  const vec4f a = (1.f - fx) * c00 + fx * c01, b = (1.f - fx) * c10 + fx * c11;
  return (1.f - fy) * a + fy * b;
}
// Texture2D_get (Texture2D.ispc:133-158)
static __device__ __noinline__ vec4f Texture2D_get(const OcTexture2D *tp, vec2f st, float pixelFootprint)
{
  const OcTexture2D &t = *tp;
  int mipLevel = 0;
  float frac = 0.f;
  if (t.maxLevel) {
    float t0 = pixelFootprint * (float)max(t.sizeX, t.sizeY);
    if (!(t.filter & TEX_FILTER_NEAREST)) t0 *= 0.5f;
    float lambda = logf(t0) / logf(2.f);
    lambda = lambda > 0.f ? lambda : 0.f; // max(0.f, lambda): NaN -> 0
    if (!(lambda >= 0.f)) lambda = 0.f;
    mipLevel = min((int)lambda, t.maxLevel - 1);
    frac = fminf(lambda - (float)mipLevel, 1.f);
  }
  const vec4f t0 = texSampleLevel(t, st, mipLevel);
  if (frac == 0.f) return t0;
  const vec4f t1 = texSampleLevel(t, st, mipLevel + 1);
  return (1.f - frac) * t0 + frac * t1;
}

// get4f(TextureParam, dg) (TextureParam.ispc:10-22): xfmPoint(xform2f, st) = p + st.x * l.vx + st.y * l.vy
__device__ __forceinline__ vec4f texParamGet4f(const OcTextureParam &tp, vec2f st, float pixelFootprint)
{
  if (tp.transformFlags & OC_TEX_TRANSFORM_2D) {
    const float *x = tp.xform2f;
    st = make_vec2f(x[4] + st.x * x[0] + st.y * x[2], x[5] + st.x * x[1] + st.y * x[3]);
    pixelFootprint *= tp.scale2f;
  }
  return Texture2D_get(tp.tex, st, pixelFootprint);
}

} // namespace oc
