// texture2d objects and the texture parameters of materials: commit-time host work of the cuda device.
// reference: texture/Texture2D.cpp:56-146 (commit, MIP chain layout), texture/MipMapGeneration.ispc:8-95 (the 3x3 / 2x2
// box reduction of one level into the next), texture/Texture2DFormats.ispc (texel get / set of the 20 formats),
// render/Material.cpp:73-165 (getTextureParam, getTextureTransform2f). The kernels sample through texture.cuh.
#include <cmath>
#include <cstring>
#include "CudaDevice.h"
#include "Objects.h"

namespace ospray {
namespace cuda {

namespace {

inline float srgbToLinear(float x) { return x <= 0.04045f ? x * (1.f / 12.92f) : std::pow((x + 0.055f) * (1.f / 1.055f), 2.4f); }
inline float linearToSrgb(float f)
{
  const float c = std::max(f, 0.f);
  return c <= 0.0031308f ? 12.92f * c : std::pow(c, 1.f / 2.4f) * 1.055f - 0.055f;
}
inline uint32_t cvt8(float f) { return (uint32_t)std::rint(255.f * std::min(std::max(f, 0.f), 1.f)); }
inline uint32_t cvt16(float f) { return (uint32_t)std::rint(65535.f * std::min(std::max(f, 0.f), 1.f)); }
inline float halfToFloat(uint16_t h) { _Float16 v; std::memcpy(&v, &h, 2); return (float)v; }
inline uint16_t floatToHalf(float f) { const _Float16 v = (_Float16)f; uint16_t h; std::memcpy(&h, &v, 2); return h; }

size_t texelBytes(uint32_t f)
{
  switch (f) {
  case OSP_TEXTURE_RGBA8: case OSP_TEXTURE_SRGBA: case OSP_TEXTURE_R32F: case OSP_TEXTURE_RA16: case OSP_TEXTURE_RA16F: return 4;
  case OSP_TEXTURE_RGBA32F: return 16;
  case OSP_TEXTURE_RGB8: case OSP_TEXTURE_SRGB: return 3;
  case OSP_TEXTURE_RGB32F: return 12;
  case OSP_TEXTURE_R8: case OSP_TEXTURE_L8: return 1;
  case OSP_TEXTURE_RA8: case OSP_TEXTURE_LA8: case OSP_TEXTURE_R16: case OSP_TEXTURE_R16F: return 2;
  case OSP_TEXTURE_RGBA16: case OSP_TEXTURE_RGBA16F: case OSP_TEXTURE_RA32F: return 8;
  case OSP_TEXTURE_RGB16: case OSP_TEXTURE_RGB16F: return 6;
  default: return 0;
  }
}

// Texture2DFormats.ispc getTexel_*
vec4f getTexel(uint32_t fmt, const void *data, uint32_t idx)
{
  const float B = 1.f / 255.f, S = 1.f / 65535.f;
  const uint8_t *u8 = (const uint8_t *)data;
  const uint16_t *u16 = (const uint16_t *)data;
  const uint32_t *u32 = (const uint32_t *)data;
  const float *f32 = (const float *)data;
  switch (fmt) {
  case OSP_TEXTURE_RGBA8:
  case OSP_TEXTURE_SRGBA: {
    const uint32_t c = u32[idx];
    const vec4f v = make_vec4f((float)(c & 0xff) * B, (float)((c >> 8) & 0xff) * B, (float)((c >> 16) & 0xff) * B, (float)(c >> 24) * B);
    return fmt == OSP_TEXTURE_RGBA8 ? v : make_vec4f(srgbToLinear(v.x), srgbToLinear(v.y), srgbToLinear(v.z), v.w);
  }
  case OSP_TEXTURE_RGBA32F: return make_vec4f(f32[4 * idx], f32[4 * idx + 1], f32[4 * idx + 2], f32[4 * idx + 3]);
  case OSP_TEXTURE_RGB8:
  case OSP_TEXTURE_SRGB: {
    const vec3f v = make_vec3f((float)u8[3 * idx] * B, (float)u8[3 * idx + 1] * B, (float)u8[3 * idx + 2] * B);
    return fmt == OSP_TEXTURE_RGB8 ? make_vec4f(v, 1.f) : make_vec4f(srgbToLinear(v.x), srgbToLinear(v.y), srgbToLinear(v.z), 1.f);
  }
  case OSP_TEXTURE_RGB32F: return make_vec4f(f32[3 * idx], f32[3 * idx + 1], f32[3 * idx + 2], 1.f);
  case OSP_TEXTURE_R8: return make_vec4f(u8[idx] * B, 0.f, 0.f, 1.f);
  case OSP_TEXTURE_R32F: return make_vec4f(f32[idx], 0.f, 0.f, 1.f);
  case OSP_TEXTURE_L8: { const float l = srgbToLinear(u8[idx] * B); return make_vec4f(l, l, l, 1.f); }
  case OSP_TEXTURE_RA8: { const uint32_t t = u16[idx]; return make_vec4f((t & 0xff) * B, 0.f, 0.f, (t >> 8) * B); }
  case OSP_TEXTURE_LA8: { const uint32_t t = u16[idx]; const float l = srgbToLinear((t & 0xff) * B); return make_vec4f(l, l, l, (t >> 8) * B); }
  case OSP_TEXTURE_RGBA16: {
    const uint32_t c0 = u32[2 * idx], c1 = u32[2 * idx + 1];
    return make_vec4f((float)(c0 & 0xffff) * S, (float)(c0 >> 16) * S, (float)(c1 & 0xffff) * S, (float)(c1 >> 16) * S);
  }
  case OSP_TEXTURE_RGB16: return make_vec4f((float)u16[3 * idx] * S, (float)u16[3 * idx + 1] * S, (float)u16[3 * idx + 2] * S, 1.f);
  case OSP_TEXTURE_RA16: { const uint32_t t = u32[idx]; return make_vec4f((t & 0xffff) * S, 0.f, 0.f, (t >> 16) * S); }
  case OSP_TEXTURE_R16: return make_vec4f(u16[idx] * S, 0.f, 0.f, 1.f);
  case OSP_TEXTURE_RA32F: return make_vec4f(f32[2 * idx], 0.f, 0.f, f32[2 * idx + 1]);
  case OSP_TEXTURE_RGBA16F: return make_vec4f(halfToFloat(u16[4 * idx]), halfToFloat(u16[4 * idx + 1]), halfToFloat(u16[4 * idx + 2]), halfToFloat(u16[4 * idx + 3]));
  case OSP_TEXTURE_RGB16F: return make_vec4f(halfToFloat(u16[3 * idx]), halfToFloat(u16[3 * idx + 1]), halfToFloat(u16[3 * idx + 2]), 1.f);
  case OSP_TEXTURE_RA16F: return make_vec4f(halfToFloat(u16[2 * idx]), 0.f, 0.f, halfToFloat(u16[2 * idx + 1]));
  case OSP_TEXTURE_R16F: return make_vec4f(halfToFloat(u16[idx]), 0.f, 0.f, 1.f);
  default: return make_vec4f(0.f, 0.f, 0.f, 1.f);
  }
}

// Texture2DFormats.ispc setTexel_*
void setTexel(uint32_t fmt, void *data, uint32_t idx, vec4f v)
{
  uint8_t *u8 = (uint8_t *)data;
  uint16_t *u16 = (uint16_t *)data;
  uint32_t *u32 = (uint32_t *)data;
  float *f32 = (float *)data;
  switch (fmt) {
  case OSP_TEXTURE_RGBA8: u32[idx] = cvt8(v.x) | (cvt8(v.y) << 8) | (cvt8(v.z) << 16) | (cvt8(v.w) << 24); break;
  case OSP_TEXTURE_SRGBA: // linear_to_srgba8: alpha stays linear
    u32[idx] = cvt8(linearToSrgb(v.x)) | (cvt8(linearToSrgb(v.y)) << 8) | (cvt8(linearToSrgb(v.z)) << 16) | (cvt8(std::max(v.w, 0.f)) << 24);
    break;
  case OSP_TEXTURE_RGBA32F: f32[4 * idx] = v.x; f32[4 * idx + 1] = v.y; f32[4 * idx + 2] = v.z; f32[4 * idx + 3] = v.w; break;
  case OSP_TEXTURE_RGB8: u8[3 * idx] = cvt8(v.x); u8[3 * idx + 1] = cvt8(v.y); u8[3 * idx + 2] = cvt8(v.z); break;
  case OSP_TEXTURE_SRGB: u8[3 * idx] = cvt8(linearToSrgb(v.x)); u8[3 * idx + 1] = cvt8(linearToSrgb(v.y)); u8[3 * idx + 2] = cvt8(linearToSrgb(v.z)); break;
  case OSP_TEXTURE_RGB32F: f32[3 * idx] = v.x; f32[3 * idx + 1] = v.y; f32[3 * idx + 2] = v.z; break;
  case OSP_TEXTURE_R8: u8[idx] = cvt8(v.x); break;
  case OSP_TEXTURE_R32F: f32[idx] = v.x; break;
  case OSP_TEXTURE_L8: u8[idx] = cvt8(linearToSrgb(v.x)); break;
  case OSP_TEXTURE_RA8: u16[idx] = (uint16_t)(cvt8(v.x) | (cvt8(v.w) << 8)); break;
  case OSP_TEXTURE_LA8: u16[idx] = (uint16_t)(cvt8(linearToSrgb(v.x)) | (cvt8(v.w) << 8)); break;
  case OSP_TEXTURE_RGBA16: u32[2 * idx] = cvt16(v.x) | (cvt16(v.y) << 16); u32[2 * idx + 1] = cvt16(v.z) | (cvt16(v.w) << 16); break;
  case OSP_TEXTURE_RGB16: u16[3 * idx] = cvt16(v.x); u16[3 * idx + 1] = cvt16(v.y); u16[3 * idx + 2] = cvt16(v.z); break;
  case OSP_TEXTURE_RA16: u32[idx] = cvt16(v.x) | (cvt16(v.w) << 16); break;
  case OSP_TEXTURE_R16: u16[idx] = cvt16(v.x); break;
  case OSP_TEXTURE_RA32F: f32[2 * idx] = v.x; f32[2 * idx + 1] = v.w; break;
  case OSP_TEXTURE_RGBA16F: u16[4 * idx] = floatToHalf(v.x); u16[4 * idx + 1] = floatToHalf(v.y); u16[4 * idx + 2] = floatToHalf(v.z); u16[4 * idx + 3] = floatToHalf(v.w); break;
  case OSP_TEXTURE_RGB16F: u16[3 * idx] = floatToHalf(v.x); u16[3 * idx + 1] = floatToHalf(v.y); u16[3 * idx + 2] = floatToHalf(v.z); break;
  case OSP_TEXTURE_RA16F: u16[2 * idx] = floatToHalf(v.x); u16[2 * idx + 1] = floatToHalf(v.w); break;
  case OSP_TEXTURE_R16F: u16[idx] = floatToHalf(v.x); break;
  default: break;
  }
}

// MipMap_generate (MipMapGeneration.ispc:8-95): every destination texel is the weighted sum of a 3x3 (odd source extent)
// or 2x2 (even) block of the finer level; a 1-D level (one extent already 1) reduces along its row
void mipGenerate(uint32_t fmt, const void *src, int sx, int sy, void *dst, int dx, int dy)
{
  if (sx > 1 && sy > 1) {
    for (int y = 0; y < dy; y++) {
      const int dsty = y * dx, srcy = 2 * y * sx;
      float rx = 1.f / (float)sx, ry = 1.f / (float)sy;
      float hx = (float)dx * rx, hy = (float)dy * ry;
      int ofs2x = 2, ofs2y = srcy + 2 * sx;
      if ((sx & 1) == 0) { hx = 0.5f; rx = 0.f; ofs2x = 0; }
      if ((sy & 1) == 0) { hy = 0.5f; ry = 0.f; ofs2y = srcy; }
      const float wy0 = hy - y * ry, wy2 = ry + y * ry;
      for (int x = 0; x < dx; x++) {
        const float tx = x * rx;
        const int s = 2 * x;
        auto row = [&](int base) {
          return (hx - tx) * getTexel(fmt, src, base + s) + hx * getTexel(fmt, src, base + s + 1) + (rx + tx) * getTexel(fmt, src, base + s + ofs2x);
        };
        const vec4f v0 = row(srcy), v1 = row(srcy + sx), v2 = row(ofs2y);
        setTexel(fmt, dst, dsty + x, wy0 * v0 + hy * v1 + wy2 * v2);
      }
    }
  } else if (sx > 1 || sy > 1) {
    const int sn = std::max(sx, sy), dn = std::max(dx, dy);
    float r = 1.f / (float)sn, h = (float)dn * r;
    int ofs2 = 2;
    if ((sn & 1) == 0) { h = 0.5f; r = 0.f; ofs2 = 0; }
    for (int x = 0; x < dn; x++) {
      const float tx = x * r;
      const int s = 2 * x;
      setTexel(fmt, dst, x, (h - tx) * getTexel(fmt, src, s) + h * getTexel(fmt, src, s + 1) + (r + tx) * getTexel(fmt, src, s + ofs2));
    }
  }
}

} // namespace

void Texture2D::commit()
{
  texData = getParamObject<Data>("data");
  if (!texData || texData->numItems.z > 1)
    throw std::runtime_error(toString() + " must have 2D 'data' array using the first two dimensions.");
  if (texData->size() > 0xffffffffull) throw std::runtime_error(toString() + " too large (over 4B texels, indexing is limited to 32bit");
  const int sx = (int)texData->numItems.x, sy = (int)texData->numItems.y;
  const uint32_t format = getParam<uint32_t>("format", (uint32_t)OSP_TEXTURE_FORMAT_INVALID);
  const uint32_t filter = getParam<uint32_t>("filter", (uint32_t)OSP_TEXTURE_FILTER_LINEAR);
  const uint32_t wrap1 = getParam<uint32_t>("wrapMode", (uint32_t)OSP_TEXTURE_WRAP_REPEAT);
  vec2ui wrapDef; wrapDef.x = wrap1; wrapDef.y = wrap1;
  const vec2ui wrap = getParam<vec2ui>("wrapMode", wrapDef);
  if (format == (uint32_t)OSP_TEXTURE_FORMAT_INVALID || texelBytes(format) == 0) throw std::runtime_error(toString() + ": invalid 'format'");
  const size_t tb = texelBytes(format);
  if (tb != sizeOf(texData->type))
    throw std::runtime_error(toString() + ": 'format'='" + std::to_string(format) + "' does not match type of 'data'!");
  // level 0, compacted (strided input is copied, Texture2D.cpp:71-80)
  std::vector<char> all((size_t)sx * sy * tb);
  for (int y = 0; y < sy; y++)
    for (int x = 0; x < sx; x++) {
      vec3ul idx; idx.x = (uint64_t)x; idx.y = (uint64_t)y; idx.z = 0;
      std::memcpy(all.data() + ((size_t)y * sx + x) * tb, texData->data(idx), tb);
    }
  std::vector<size_t> levelOffset{0};
  if (!device.disableMipMapGeneration) {
    int lx = sx, ly = sy;
    size_t srcOff = 0;
    while (!(lx == 1 && ly == 1)) {
      const int nx = std::max(lx / 2, 1), ny = std::max(ly / 2, 1);
      const size_t dstOff = all.size();
      all.resize(all.size() + (size_t)nx * ny * tb);
      mipGenerate(format, all.data() + srcOff, lx, ly, all.data() + dstOff, nx, ny);
      levelOffset.push_back(dstOff);
      srcOff = dstOff; lx = nx; ly = ny;
      if (levelOffset.size() >= OC_MAX_MIPMAP_LEVEL) break;
    }
  }
  // 16-byte texels are fetched with 128-bit loads: keep every level 16-byte aligned in the device copy
  std::vector<char> packed;
  std::vector<size_t> devOffset;
  for (size_t l = 0; l < levelOffset.size(); l++) {
    const size_t end = l + 1 < levelOffset.size() ? levelOffset[l + 1] : all.size();
    const size_t o = (packed.size() + 15) & ~size_t(15);
    packed.resize(o + (end - levelOffset[l]));
    std::memcpy(packed.data() + o, all.data() + levelOffset[l], end - levelOffset[l]);
    devOffset.push_back(o);
  }
  texels.upload(packed.data(), packed.size());
  sh = OcTexture2D{};
  for (size_t l = 0; l < devOffset.size(); l++) sh.data[l] = texels.as<char>() + devOffset[l];
  sh.sizeX = sx; sh.sizeY = sy;
  sh.maxLevel = (int32_t)devOffset.size() - 1;
  sh.format = format; sh.filter = filter; sh.wrapS = wrap.x; sh.wrapT = wrap.y;
  sh.hasAlpha = format == OSP_TEXTURE_RGBA8 || format == OSP_TEXTURE_SRGBA || format == OSP_TEXTURE_RA8 || format == OSP_TEXTURE_LA8
      || format == OSP_TEXTURE_RGBA32F || format == OSP_TEXTURE_RGBA16 || format == OSP_TEXTURE_RA16 || format == OSP_TEXTURE_RA16F
      || format == OSP_TEXTURE_RA32F || format == OSP_TEXTURE_RGBA16F;
  shDev.upload(&sh, sizeof(sh));
}

// Material::getTextureParam + getTextureTransform2f (render/Material.cpp:73-165): translation, scale and rotation (about the
// texture centre) and a free linear part compose right to left into one 2D affine map of the texture coordinates
OcTextureParam Material::getTextureParam(const std::string &name, Ref<Texture2D> &keep)
{
  struct A2 { float l00, l10, l01, l11, px, py; }; // l.vx = (l00, l10), l.vy = (l01, l11)
  auto mul = [](const A2 &a, const A2 &b) { // a * b: (a.l * b.l, a.l * b.p + a.p)
    A2 r;
    r.l00 = a.l00 * b.l00 + a.l01 * b.l10; r.l10 = a.l10 * b.l00 + a.l11 * b.l10;
    r.l01 = a.l00 * b.l01 + a.l01 * b.l11; r.l11 = a.l10 * b.l01 + a.l11 * b.l11;
    r.px = a.l00 * b.px + a.l01 * b.py + a.px; r.py = a.l10 * b.px + a.l11 * b.py + a.py;
    return r;
  };
  auto translate = [](float x, float y) { return A2{1.f, 0.f, 0.f, 1.f, x, y}; };
  const A2 one{1.f, 0.f, 0.f, 1.f, 0.f, 0.f};
  A2 x = one;
  bool have = false;
  const std::string n = name + ".";
  if (hasParam(n + "translation")) {
    const vec2f t = getParam<vec2f>(n + "translation", make_vec2f(0.f, 0.f));
    x = translate(-t.x, -t.y);
    have = true;
  }
  if (hasParam(n + "scale")) {
    const vec2f sc = getParam<vec2f>(n + "scale", make_vec2f(1.f, 1.f));
    x = mul(x, translate(0.5f, 0.5f));
    x = mul(x, A2{1.f / sc.x, 0.f, 0.f, 1.f / sc.y, 0.f, 0.f});
    x = mul(x, translate(-0.5f, -0.5f));
    have = true;
  }
  if (hasParam(n + "rotation")) {
    const float r = -(getParam<float>(n + "rotation", 0.f) * (float)(M_PI / 180.0)); // affine2f::rotate(-deg2rad(r))
    const float s = std::sin(r), c = std::cos(r);
    x = mul(x, translate(0.5f, 0.5f));
    x = mul(x, A2{c, s, -s, c, 0.f, 0.f});
    x = mul(x, translate(-0.5f, -0.5f));
    have = true;
  }
  {
    auto it = params.find(n + "transform");
    if (it != params.end() && it->second.type == OSP_LINEAR2F) {
      const linear2f l = getParam<linear2f>(n + "transform", linear2f{});
      x = mul(x, A2{l.vx.x, l.vx.y, l.vy.x, l.vy.y, 0.f, 0.f});
      have = true;
    }
  }
  OcTextureParam p{};
  Texture2D *t = getParamObject<Texture2D>(name);
  keep = t;
  p.tex = t ? t->devPtr() : nullptr;
  p.transformFlags = have ? OC_TEX_TRANSFORM_2D : OC_TEX_TRANSFORM_NONE;
  p.xform2f[0] = x.l00; p.xform2f[1] = x.l10; p.xform2f[2] = x.l01; p.xform2f[3] = x.l11; p.xform2f[4] = x.px; p.xform2f[5] = x.py;
  p.scale2f = std::sqrt(x.l00 * x.l11 - x.l01 * x.l10);
  return p;
}

} // namespace cuda
} // namespace ospray
