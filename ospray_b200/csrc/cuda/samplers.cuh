// Integer-exact sampler sequences of the reference, as __device__ code.
// reference: modules/cpu/math/random.ih (hashes :15-78, PCG32 :116-149, TEA :159-190,
// Owen hash :234-266, LDSampler :283-394), math/sobol.ih:38-80, math/halton.ih:38-86,
// render/util.ih:22-60 (Z-order).  u32 -> float uses rkcommon's to_float_unorm
// (un-vendored; restated as round-to-nearest u32->f32 times 2^-32).
#pragma once
#include <cstdint>
#include "../common/osp_math.h"
#include "tables.inc"

namespace oc {
using namespace ospm;

__device__ __forceinline__ float to_float_unorm(uint32_t a) { return __uint2float_rn(a) * 0x1.0p-32f; }
__device__ __forceinline__ uint32_t reverseBits(uint32_t x) { return __brev(x); }

__device__ __forceinline__ uint32_t murmur3_mix(uint32_t hash, uint32_t k)
{
  k *= 0xcc9e2d51u;
  k = (k << 15) | (k >> 17);
  k *= 0x1b873593u;
  hash ^= k;
  hash = ((hash << 13) | (hash >> 19)) * 5u + 0xe6546b64u;
  return hash;
}
__device__ __forceinline__ uint32_t murmur3_finalize(uint32_t h)
{
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t tripleHash(uint32_t x)
{
  x ^= x >> 17; x *= 0xed5ad4bbu; x ^= x >> 11; x *= 0xac4c1b51u;
  x ^= x >> 15; x *= 0x31848babu; x ^= x >> 14;
  return x;
}
__device__ __forceinline__ float CranleyPattersonRotation(float x, float dx)
{
  x += dx;
  if (x >= 1.f) x -= 1.f;
  return x;
}

// ---- PCG32 (random.ih:116-149) ------------------------------------------------------
struct RandomSampler { uint64_t state; uint32_t stream; };
__device__ __forceinline__ uint32_t RandomSampler_pcg32(RandomSampler &s)
{
  const uint64_t old = s.state;
  s.state = old * 6364136223846793005ULL + (uint64_t)(s.stream | 1u);
  const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  const uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}
__device__ __forceinline__ float RandomSampler_getFloat(RandomSampler &s) { return to_float_unorm(RandomSampler_pcg32(s)); }
__device__ __forceinline__ void RandomSampler_init(RandomSampler &s, uint32_t seed, uint32_t stream)
{
  s.state = 0;
  s.stream = (stream << 1u) | 1u;
  // the reference hashes through a 32-bit function on the (zero) 64-bit state
  s.state = murmur3_mix((uint32_t)s.state, seed);
  s.state = murmur3_finalize((uint32_t)s.state);
  RandomSampler_pcg32(s);
  s.state += seed;
  RandomSampler_pcg32(s);
}

// ---- TEA (random.ih:159-190) ---------------------------------------------------------
__device__ __forceinline__ void tea8(uint32_t &v0, uint32_t &v1)
{
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sum += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
  }
}

// ---- Owen hashing (random.ih:234-266) ------------------------------------------------
__device__ __forceinline__ uint32_t OwenHash2(uint32_t x, uint32_t seed)
{
  x ^= x * 0x3d20adeau;
  x += seed;
  x *= (seed >> 16) | 1u;
  x ^= x * 0x05526c56u;
  x ^= x * 0x53a22864u;
  return x;
}
__device__ __forceinline__ uint32_t OwenScramble4(uint32_t x, uint32_t seed)
{
  x = reverseBits(x);
  x ^= x * 0x3d20adeau;
  x ^= (x >> 1) & (x << 1) & 0x55555555u;
  x += seed;
  x *= (seed >> 16) | 1u;
  x ^= (x >> 1) & (x << 1) & 0x55555555u;
  x ^= x * 0x05526c56u;
  x ^= x * 0x53a22864u;
  return reverseBits(x);
}
__device__ __forceinline__ uint32_t partitionZOrder8(uint32_t n)
{
  n &= 0x000000FFu;
  n = (n | (n << 4)) & 0x0F0F0F0Fu;
  n = (n | (n << 2)) & 0x33333333u;
  n = (n | (n << 1)) & 0x55555555u;
  return n;
}
__device__ __forceinline__ uint32_t interleaveZOrder16(uint32_t x, uint32_t y)
{
  return partitionZOrder8(x) | (partitionZOrder8(y) << 1);
}

// ---- Sobol' (sobol.ih:38-80) -----------------------------------------------------------
template <int DIMS>
__device__ __forceinline__ void Sobol_revSample(uint32_t revIndex, uint32_t (&r)[4])
{
  // The reference XORs one matrix column per set bit of the index (bit-serial loop). The map is linear over GF(2): the
  // same value is the XOR of four table entries, one per byte of the index (kSobolRevByte, tools/gen_tables.py) -- no
  // data-dependent trip count, so the lanes of a warp stay together. Zero bytes (indices masked to 16 bits) are skipped.
  const uint4 *tab = reinterpret_cast<const uint4 *>(kSobolRevByte);
  uint4 a = __ldg(tab + (revIndex >> 24));
  const uint4 b = __ldg(tab + 256 + ((revIndex >> 16) & 0xffu));
  a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w;
  if (revIndex & 0xffffu) {
    const uint4 c = __ldg(tab + 512 + ((revIndex >> 8) & 0xffu)), d = __ldg(tab + 768 + (revIndex & 0xffu));
    a.x ^= c.x ^ d.x; a.y ^= c.y ^ d.y; a.z ^= c.z ^ d.z; a.w ^= c.w ^ d.w;
  }
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
}

// ---- LDSampler (random.ih:283-394) -----------------------------------------------------
#define OC_SCRAMBLE0 0x9374b0d3u
#define OC_SCRAMBLE1 0x1ec38426u
#define OC_SCRAMBLE2 0xdb81c29cu
#define OC_SCRAMBLE3 0x5628372bu
#define OC_SCRAMBLE4 0xbcb745bdu

struct LDSampler
{
  uint32_t revIndex, scramble, baseIndexLight, idxMask, idxMaskLight;
};
__device__ __forceinline__ void LDSampler_init(LDSampler &s, uint32_t seed, uint32_t sampleIndex,
    uint32_t sampleOffsetLight, uint32_t bits, uint32_t bitsLight)
{
  s.revIndex = reverseBits(sampleIndex);
  s.scramble = seed;
  s.baseIndexLight = sampleOffsetLight;
  s.idxMask = bits > 31 ? 0xffffffffu : ~(0xffffffffu >> bits);
  s.idxMaskLight = bitsLight > 31 ? 0xffffffffu : ~(0xffffffffu >> bitsLight);
}
__device__ __forceinline__ void LDSampler_nextGroup(LDSampler &s) { s.scramble = tripleHash(s.scramble); }
__device__ __forceinline__ float LDSamplerScramble(uint32_t sample, uint32_t scramble)
{
  return to_float_unorm(reverseBits(OwenHash2(sample, scramble)));
}
__device__ __forceinline__ vec2f LDSampler_get3LightSamples(const LDSampler &s, uint32_t idx, bool split, float &ss)
{
  uint32_t revIndex = split ? reverseBits(s.baseIndexLight + idx) : s.revIndex;
  revIndex = OwenHash2(revIndex, s.scramble);
  revIndex &= split ? s.idxMaskLight : s.idxMask;
  const uint32_t r0 = reverseBits(revIndex);
  uint32_t r[4];
  Sobol_revSample<2>(revIndex, r);
  vec2f res;
  res.x = LDSamplerScramble(r0, s.scramble ^ OC_SCRAMBLE0);
  res.y = LDSamplerScramble(r[0], s.scramble ^ OC_SCRAMBLE1);
  ss = LDSamplerScramble(r[1], s.scramble ^ OC_SCRAMBLE2);
  return res;
}
__device__ __forceinline__ vec2f LDSampler_getNext4Samples(LDSampler &s, float &ss, uint32_t &u)
{
  LDSampler_nextGroup(s);
  const uint32_t revIndex = OwenHash2(s.revIndex, s.scramble) & s.idxMask;
  const uint32_t r0 = reverseBits(revIndex);
  uint32_t r[4];
  Sobol_revSample<3>(revIndex, r);
  vec2f res;
  res.x = LDSamplerScramble(r0, s.scramble ^ OC_SCRAMBLE0);
  res.y = LDSamplerScramble(r[0], s.scramble ^ OC_SCRAMBLE1);
  ss = LDSamplerScramble(r[1], s.scramble ^ OC_SCRAMBLE2);
  u = r[2];
  return res;
}
__device__ __forceinline__ float LDSampler_finalizeDim3(const LDSampler &s, uint32_t revRes)
{
  return LDSamplerScramble(revRes, s.scramble ^ OC_SCRAMBLE3);
}
__device__ __forceinline__ vec2f LDSampler_getNext5Samples(LDSampler &s, uint32_t (&ss)[3])
{
  LDSampler_nextGroup(s);
  const uint32_t revIndex = OwenHash2(s.revIndex, s.scramble) & s.idxMask;
  const uint32_t r0 = reverseBits(revIndex);
  uint32_t r[4];
  Sobol_revSample<4>(revIndex, r);
  vec2f res;
  res.x = LDSamplerScramble(r0, s.scramble ^ OC_SCRAMBLE0);
  res.y = LDSamplerScramble(r[0], s.scramble ^ OC_SCRAMBLE1);
  ss[0] = r[1]; ss[1] = r[2]; ss[2] = r[3];
  return res;
}
__device__ __forceinline__ vec2f LDSampler_finalizeDim23(const LDSampler &s, const uint32_t (&rr)[3])
{
  vec2f res;
  res.x = LDSamplerScramble(rr[0], s.scramble ^ OC_SCRAMBLE2);
  res.y = LDSamplerScramble(rr[1], s.scramble ^ OC_SCRAMBLE3);
  return res;
}
__device__ __forceinline__ float LDSampler_finalizeDim4(const LDSampler &s, const uint32_t (&rr)[3])
{
  return LDSamplerScramble(rr[2], s.scramble ^ OC_SCRAMBLE4);
}

// ---- Halton (halton.ih:38-86) ----------------------------------------------------------
__device__ __forceinline__ float Halton_sample2(uint32_t index) { return to_float_unorm(reverseBits(index)); }
__device__ __forceinline__ float Halton_sample3(uint32_t index)
{
  return __uint2float_rn(kHaltonPerm3[index % 243u] * 14348907u + kHaltonPerm3[(index / 243u) % 243u] * 59049u
             + kHaltonPerm3[(index / 59049u) % 243u] * 243u + kHaltonPerm3[(index / 14348907u) % 243u])
      * (float)(1.0 / 3486784401u);
}
__device__ __forceinline__ float PermHalton_sample5(uint32_t index)
{
  return __uint2float_rn(kHaltonPerm5[index % 125u] * 1953125u + kHaltonPerm5[(index / 125u) % 125u] * 15625u
             + kHaltonPerm5[(index / 15625u) % 125u] * 125u + kHaltonPerm5[(index / 1953125u) % 125u])
      * (float)(1.0 / 244140625u);
}
__device__ __forceinline__ float PermHalton_sample7(uint32_t index)
{
  return __uint2float_rn(kHaltonPerm7[index % 343u] * 117649u + kHaltonPerm7[(index / 343u) % 343u] * 343u
             + kHaltonPerm7[(index / 117649u) % 343u])
      * (float)(1.0 / 40353607u);
}

} // namespace oc
