// Framebuffer pixel ops: streaming, HBM-bound kernels (vectorised 128-bit accesses, grid sized
// to a multiple of the SM count).
// reference: fb/frame_ops/Variance.ispc:11-59, ToneMapper.ispc:12-41, ColorConversion.ispc:11-36,
// fb/LocalFB.ispc:14-78 (accumulate), fb/LocalFB.cpp:286-353 (writeTiles);
// rkcommon cvt_uint32 / linear_to_srgba8 (un-vendored, restated).
#include "runtime.h"
#include "../common/osp_math.h"

namespace oc {
using namespace ospm;

__device__ __forceinline__ uint32_t cvt_uint32(float f) { return (uint32_t)rintf(255.f * fminf(fmaxf(f, 0.f), 1.f)); }
__device__ __forceinline__ uint32_t cvt_uint32(float4 v)
{
  return cvt_uint32(v.x) | (cvt_uint32(v.y) << 8) | (cvt_uint32(v.z) << 16) | (cvt_uint32(v.w) << 24);
}
__device__ __forceinline__ float linear_to_srgb(float f)
{
  const float c = fmaxf(f, 0.f);
  return c <= 0.0031308f ? 12.92f * c : powf(c, 1.f / 2.4f) * 1.055f - 0.055f;
}

// cvt_uint32(linear_to_srgb(f)), bit for bit, without the full-precision powf for all but ~0.2 % of the values: the byte is
// round(255 * s). s is first evaluated with the hardware exp2 / log2 (__powf: relative error < 1e-6, i.e. < 3e-4 in 255 * s);
// unless that estimate lands within 5e-4 of a rounding boundary (x.5) it rounds to the same byte as the exact evaluation,
// and the few values that do land there take the exact path. The 4K tonemap + conversion pass goes from ALU-bound (0.12 of
// the HBM roofline) towards the memory bound.
__device__ __forceinline__ uint32_t srgb8(float f)
{
  const float c = fmaxf(f, 0.f);
  if (c <= 0.0031308f) return cvt_uint32(12.92f * c);
  const float v = 255.f * fminf(fmaxf(__powf(c, 1.f / 2.4f) * 1.055f - 0.055f, 0.f), 1.f);
  const float r = rintf(v);
  if (fabsf(fabsf(v - r) - 0.5f) < 5e-4f) return cvt_uint32(linear_to_srgb(f));
  return (uint32_t)r;
}
// x^a / (x^(a d) b + c) with the hardware exp2 / log2: relative error < 3e-6 (the reference's own ISPC pow is no tighter)
__device__ __forceinline__ float toneCurve(float x, const OcToneMapper &tm)
{
  if (!(x > 0.f)) return powf(x, tm.a) / (powf(x, tm.a * tm.d) * tm.b + tm.c); // 0, negative, NaN: the library's special cases
  const float l = __log2f(x);
  return __fdividef(exp2f(tm.a * l), exp2f(tm.a * tm.d * l) * tm.b + tm.c);
}

__device__ __forceinline__ float4 toneMap(float4 col, const OcToneMapper &tm)
{
  float x = col.x * tm.exposure, y = col.y * tm.exposure, z = col.z * tm.exposure;
  if (tm.acesColor) {
    const float nx = 0.5972782409f * x + 0.3545713181f * y + 0.0482176639f * z;
    const float ny = 0.0760130499f * x + 0.9083220973f * y + 0.0156579968f * z;
    const float nz = 0.0284085382f * x + 0.1338243154f * y + 0.8375684636f * z;
    x = nx; y = ny; z = nz;
  }
  x = toneCurve(x, tm);
  y = toneCurve(y, tm);
  z = toneCurve(z, tm);
  if (tm.acesColor) {
    const float nx = 1.6047539945f * x - 0.5310794927f * y - 0.0736720338f * z;
    const float ny = -0.1020831870f * x + 1.1081322801f * y - 0.0060518756f * z;
    const float nz = -0.0032670420f * x - 0.0727552477f * y + 1.0760219533f * z;
    x = nx; y = ny; z = nz;
  }
  return make_float4(fminf(fmaxf(x, 0.f), 1.f), fminf(fmaxf(y, 0.f), 1.f), fminf(fmaxf(z, 0.f), 1.f), col.w);
}

// one pass: optional tonemap -> optional float4 store (post-processed colour) -> format conversion
// numRanks > 1: only the pixels of the 64x64 tiles this rank owns (tile % numRanks == rank) are converted; the other
// pixels of the display buffer belong to the tile exchange (peers store their tiles there, possibly while this runs)
template <int FORMAT, bool TONEMAP>
// (`in` and `ppOut` may be the same buffer: chained tonemappers run in place, every pixel is read before it is written)
__global__ void __launch_bounds__(256) postprocessKernel(const float4 *in, float4 *ppOut,
    void *__restrict__ out, uint32_t n, OcToneMapper tm, uint32_t sizeX, uint32_t numRanks, uint32_t rank)
{
  const uint32_t tilesX = (sizeX + 63u) >> 6;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (numRanks > 1) {
      const uint32_t y = i / sizeX, x = i - y * sizeX;
      if (((y >> 6) * tilesX + (x >> 6)) % numRanks != rank) continue;
    }
    float4 v = __ldcs(in + i);
    if (TONEMAP) {
      v = toneMap(v, tm);
      if (ppOut) __stcs(ppOut + i, v);
    }
    if (FORMAT == 1) { // OSP_FB_RGBA8
      reinterpret_cast<uint32_t *>(out)[i] = cvt_uint32(v);
    } else if (FORMAT == 2) { // OSP_FB_SRGBA
      reinterpret_cast<uint32_t *>(out)[i] = srgb8(v.x) | (srgb8(v.y) << 8) | (srgb8(v.z) << 16) | (cvt_uint32(v.w) << 24);
    } else if (FORMAT == 3) { // OSP_FB_RGBA32F
      __stcs(reinterpret_cast<float4 *>(out) + i, make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fminf(fmaxf(v.w, 0.f), 1.f)));
    }
  }
}

// per-8x8-task error: one warp per task, lanes stride the 64 pixels
__global__ void __launch_bounds__(256) varianceKernel(const float4 *__restrict__ color, const float4 *__restrict__ variance,
    int sizeX, int sizeY, int taskSize, int tasksX, int tasksY, float *__restrict__ taskError)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= tasksX * tasksY) return;
  const int tX = warp % tasksX, tY = warp / tasksX;
  float v = 0.f;
  int cnt = 0;
  for (int p = lane; p < taskSize * taskSize; p += 32) {
    const int x = tX * taskSize + p % taskSize, y = tY * taskSize + p / taskSize;
    if (x >= sizeX || y >= sizeY) continue;
    const size_t id = (size_t)y * sizeX + x;
    const float4 c = __ldg(color + id), w = __ldg(variance + id);
    const float den2 = (c.x + c.y + c.z) + (1.f - c.w);
    if (den2 > 0.f) {
      cnt++;
      v += (fabsf(c.x - w.x) + fabsf(c.y - w.y) + fabsf(c.z - w.z) + fabsf(c.w - w.w)) * rsqrtf(den2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    if (cnt) v *= 1.f / (float)cnt;
    taskError[warp] = v * 64.f;
  }
}

// running-mean accumulation of an externally produced sample image (colour [+depth])
__global__ void __launch_bounds__(256) accumulateKernel(OcFrameBuffer fb, const float4 *__restrict__ sc, const float *__restrict__ sz,
    uint32_t n)
{
  const int accumID = fb.frameID;
  const float f = 1.f / (float)(accumID + 1);
  const float fv = 1.f / (float)(accumID / 2 + 1);
  float4 *color = reinterpret_cast<float4 *>(fb.color), *var = reinterpret_cast<float4 *>(fb.variance);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 s = __ldcs(sc + i);
    if (color) {
      float4 o = s;
      if (!(f > 0.8f)) {
        const float4 c = color[i];
        o.x = lerpf(f, c.x, s.x); o.y = lerpf(f, c.y, s.y); o.z = lerpf(f, c.z, s.z); o.w = lerpf(f, c.w, s.w);
      }
      color[i] = o;
    }
    if (fb.accumulateVariance && var) {
      float4 o = s;
      if (!(fv > 0.8f)) {
        const float4 c = var[i];
        o.x = lerpf(fv, c.x, s.x); o.y = lerpf(fv, c.y, s.y); o.z = lerpf(fv, c.z, s.z); o.w = lerpf(fv, c.w, s.w);
      }
      var[i] = o;
    }
    if (fb.depth && sz) {
      const float z = __ldcs(sz + i);
      if (accumID == 0 || fabsf(z) < fabsf(fb.depth[i])) fb.depth[i] = z;
    }
  }
}

// copy the 64x64 tiles with tile % numRanks == srcRank; one CTA per tile row-chunk, 16-byte lanes
__global__ void __launch_bounds__(256) gatherTilesKernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int sizeX,
    int sizeY, uint32_t bpp, uint32_t numRanks, uint32_t srcRank, uint32_t tilesX, uint32_t ownedTiles)
{
  const uint32_t ownedTile = blockIdx.x;
  if (ownedTile >= ownedTiles) return;
  const uint32_t tileID = srcRank + ownedTile * numRanks;
  const uint32_t tx = tileID % tilesX, ty = tileID / tilesX;
  const int x0 = tx * 64, y0 = ty * 64;
  const int w = min(64, sizeX - x0), h = min(64, sizeY - y0);
  const uint32_t rowBytes = (uint32_t)w * bpp;
  for (int r = threadIdx.x >> 4; r < h; r += blockDim.x >> 4) { // 16 threads per row
    const size_t off = ((size_t)(y0 + r) * sizeX + x0) * bpp;
    if (((off | rowBytes) & 15u) == 0) {
      for (uint32_t b = (threadIdx.x & 15) * 16; b < rowBytes; b += 256)
        *reinterpret_cast<uint4 *>(dst + off + b) = *reinterpret_cast<const uint4 *>(src + off + b);
    } else {
      for (uint32_t b = (threadIdx.x & 15) * 4; b < rowBytes; b += 64)
        *reinterpret_cast<uint32_t *>(dst + off + b) = *reinterpret_cast<const uint32_t *>(src + off + b);
    }
  }
}

// tile-major staging for the NCCL gather: owned tile k of `rank` <-> packed + k * 64*64*bpp (tiles on the
// right / top image border are stored padded to 64x64 so that every rank's shard has a fixed pitch).
// PACK: image -> staging (before ncclSend); !PACK: staging -> image on GPU 0 (after ncclRecv), the device
// analogue of LocalFrameBuffer::writeTiles (fb/LocalFB.cpp:286-353). 16-byte lanes, one CTA per tile.
template <bool PACK>
__global__ void __launch_bounds__(256) packTilesKernel(uint8_t *__restrict__ packed, uint8_t *__restrict__ image, int sizeX, int sizeY,
    uint32_t bpp, uint32_t numRanks, uint32_t rank, uint32_t tilesX, uint32_t ownedTiles)
{
  const uint32_t ownedTile = blockIdx.x;
  if (ownedTile >= ownedTiles) return;
  const uint32_t tileID = rank + ownedTile * numRanks;
  const uint32_t tx = tileID % tilesX, ty = tileID / tilesX;
  const int x0 = tx * 64, y0 = ty * 64;
  const int w = min(64, sizeX - x0), h = min(64, sizeY - y0);
  const uint32_t rowBytes = (uint32_t)w * bpp, tilePitch = 64u * bpp;
  uint8_t *tile = packed + (size_t)ownedTile * 64u * tilePitch;
  for (int r = threadIdx.x >> 4; r < h; r += blockDim.x >> 4) { // 16 threads per row
    const size_t off = ((size_t)(y0 + r) * sizeX + x0) * bpp;
    uint8_t *a = tile + (size_t)r * tilePitch, *b = image + off;
    if (((off | rowBytes) & 15u) == 0) {
      for (uint32_t k = (threadIdx.x & 15) * 16; k < rowBytes; k += 256) {
        if (PACK) *reinterpret_cast<uint4 *>(a + k) = *reinterpret_cast<const uint4 *>(b + k);
        else *reinterpret_cast<uint4 *>(b + k) = *reinterpret_cast<const uint4 *>(a + k);
      }
    } else {
      for (uint32_t k = (threadIdx.x & 15) * 4; k < rowBytes; k += 64) {
        if (PACK) *reinterpret_cast<uint32_t *>(a + k) = *reinterpret_cast<const uint32_t *>(b + k);
        else *reinterpret_cast<uint32_t *>(b + k) = *reinterpret_cast<const uint32_t *>(a + k);
      }
    }
  }
}

static int numSMs()
{
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
static uint32_t streamGrid(uint32_t n, int perSM = 8)
{
  const uint32_t want = (n + 255) / 256;
  const uint32_t cap = (uint32_t)numSMs() * perSM; // multiple of the SM count
  return want < cap ? (want ? want : 1) : cap;
}

} // namespace oc

using namespace oc;

extern "C" int ospcuda_fb_postprocess_tiles(const OcVec4f *d_in, OcVec4f *d_ppOut, void *d_out, int sizeX, int sizeY,
    uint32_t colorFormat, const OcToneMapper *tm, uint32_t numRanks, uint32_t rank)
{
  const uint32_t numPixels = (uint32_t)sizeX * (uint32_t)sizeY;
  if (numPixels == 0) return 0;
  if (numRanks == 0 || rank >= numRanks) return oc_set_error(cudaErrorInvalidValue, "ospcuda_fb_postprocess_tiles: bad rank");
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  const uint32_t grid = streamGrid(numPixels);
  const float4 *in = reinterpret_cast<const float4 *>(d_in);
  float4 *pp = reinterpret_cast<float4 *>(d_ppOut);
  OcToneMapper t{};
  if (tm) t = *tm;
#define OC_PP(F)                                                                                   \
  if (tm) postprocessKernel<F, true><<<grid, 256, 0, st>>>(in, pp, d_out, numPixels, t, (uint32_t)sizeX, numRanks, rank); \
  else postprocessKernel<F, false><<<grid, 256, 0, st>>>(in, pp, d_out, numPixels, t, (uint32_t)sizeX, numRanks, rank);
  switch (colorFormat) {
  case 0: if (!tm) return 0; OC_PP(0) break;
  case 1: OC_PP(1) break;
  case 2: OC_PP(2) break;
  case 3: OC_PP(3) break;
  default: return oc_set_error(cudaErrorInvalidValue, "ospcuda_fb_postprocess: bad colour format");
  }
#undef OC_PP
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_fb_postprocess");
}

extern "C" int ospcuda_fb_postprocess(const OcVec4f *d_in, OcVec4f *d_ppOut, void *d_out, uint32_t numPixels,
    uint32_t colorFormat, const OcToneMapper *tm)
{
  return ospcuda_fb_postprocess_tiles(d_in, d_ppOut, d_out, (int)numPixels, 1, colorFormat, tm, 1, 0);
}

extern "C" int ospcuda_fb_color_convert(const OcVec4f *d_in, void *d_out, uint32_t numPixels, uint32_t colorFormat)
{
  return ospcuda_fb_postprocess(d_in, nullptr, d_out, numPixels, colorFormat, nullptr);
}

extern "C" int ospcuda_fb_tonemap(const OcVec4f *d_in, OcVec4f *d_out, uint32_t numPixels, const OcToneMapper *tm)
{
  return ospcuda_fb_postprocess(d_in, d_out, nullptr, numPixels, 0, tm);
}

extern "C" int ospcuda_fb_variance(const OcVec4f *d_color, const OcVec4f *d_variance, int sizeX, int sizeY, int taskSize,
    float *d_taskError)
{
  const int tasksX = (sizeX + taskSize - 1) / taskSize, tasksY = (sizeY + taskSize - 1) / taskSize;
  const int warps = tasksX * tasksY;
  varianceKernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)ospcuda_stream()>>>(reinterpret_cast<const float4 *>(d_color),
      reinterpret_cast<const float4 *>(d_variance), sizeX, sizeY, taskSize, tasksX, tasksY, d_taskError);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_fb_variance");
}

extern "C" int ospcuda_fb_accumulate(const OcFrameBuffer *fb, const OcVec4f *d_sampleColor, const float *d_sampleDepth)
{
  const uint32_t n = (uint32_t)fb->sizeX * (uint32_t)fb->sizeY;
  if (n == 0) return 0;
  accumulateKernel<<<streamGrid(n), 256, 0, (cudaStream_t)ospcuda_stream()>>>(*fb, reinterpret_cast<const float4 *>(d_sampleColor),
      d_sampleDepth, n);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_fb_accumulate");
}

extern "C" int ospcuda_fb_gather_tiles(void *d_dst, const void *d_src, int sizeX, int sizeY, uint32_t bytesPerPixel,
    uint32_t numRanks, uint32_t srcRank)
{
  const uint32_t tilesX = (sizeX + 63) / 64, tilesY = (sizeY + 63) / 64, total = tilesX * tilesY;
  if (srcRank >= total) return 0;
  const uint32_t owned = (total - srcRank + numRanks - 1) / numRanks;
  gatherTilesKernel<<<owned, 256, 0, (cudaStream_t)ospcuda_stream()>>>((uint8_t *)d_dst, (const uint8_t *)d_src, sizeX, sizeY,
      bytesPerPixel, numRanks, srcRank, tilesX, owned);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_fb_gather_tiles");
}

static int packTiles(bool pack, void *d_packed, void *d_image, int sizeX, int sizeY, uint32_t bpp, uint32_t numRanks, uint32_t rank)
{
  if (bpp % 4 != 0 || numRanks == 0) return oc_set_error(cudaErrorInvalidValue, "ospcuda_fb_pack_tiles: bytesPerPixel must be a multiple of 4");
  const uint32_t tilesX = (sizeX + 63) / 64, tilesY = (sizeY + 63) / 64, total = tilesX * tilesY;
  if (rank >= total) return 0;
  const uint32_t owned = (total - rank + numRanks - 1) / numRanks;
  cudaStream_t st = (cudaStream_t)ospcuda_stream();
  if (pack) packTilesKernel<true><<<owned, 256, 0, st>>>((uint8_t *)d_packed, (uint8_t *)d_image, sizeX, sizeY, bpp, numRanks, rank, tilesX, owned);
  else packTilesKernel<false><<<owned, 256, 0, st>>>((uint8_t *)d_packed, (uint8_t *)d_image, sizeX, sizeY, bpp, numRanks, rank, tilesX, owned);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_fb_pack_tiles");
}
extern "C" int ospcuda_fb_pack_tiles(void *d_packed, const void *d_image, int sizeX, int sizeY, uint32_t bytesPerPixel, uint32_t numRanks,
    uint32_t rank)
{
  return packTiles(true, d_packed, (void *)d_image, sizeX, sizeY, bytesPerPixel, numRanks, rank);
}
extern "C" int ospcuda_fb_unpack_tiles(void *d_image, const void *d_packed, int sizeX, int sizeY, uint32_t bytesPerPixel, uint32_t numRanks,
    uint32_t srcRank)
{
  return packTiles(false, (void *)d_packed, d_image, sizeX, sizeY, bytesPerPixel, numRanks, srcRank);
}
