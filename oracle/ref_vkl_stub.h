// Stand-in for Open VKL 2.0.1's scalar C++ API (openvkl/openvkl.h + openvkl/device/openvkl.h, NOT vendored in
// /root/reference): only what the reference's volume kernels call (SURVEY.md 8c lists the call sites), for
// structuredRegular float volumes.  TEST INFRASTRUCTURE ONLY.
//
// Restated from Open VKL's published behaviour, identical to oracle.cpp's restatement: vertex-centred trilinear
// sampling that returns NaN outside the grid, forward-difference gradient with a one-cell step, and an interval
// iterator that yields the clipped ray range as ONE interval with nominalDeltaT = min_i spacing_i / |dir_i|
// (any macrocell subdivision of that range is mathematically the same march: skipped cells have zero opacity and the
// integrators compensate dt through `emptySpace`).
#pragma once
#include <cstdint>
#include "../include/ospcuda.h"

struct vkl_vec3f { float x, y, z; };
struct vkl_range1f { float lower, upper; };
struct VKLObjectHandle { void *host = nullptr; void *device = nullptr; };
typedef VKLObjectHandle VKLVolume;
typedef VKLObjectHandle VKLSampler;
typedef VKLObjectHandle VKLIntervalIteratorContext;
typedef VKLObjectHandle VKLHitIteratorContext;
typedef void *VKLIntervalIterator;
typedef void *VKLHitIterator;
struct VKLInterval { vkl_range1f tRange; vkl_range1f valueRange; float nominalDeltaT; };
struct VKLHit { float t; float sample; float epsilon; };
typedef uint32_t VKLFeatureFlags;
#define VKL_FEATURE_FLAGS_NONE 0u
#define VKL_FEATURE_FLAGS_DEFAULT 0xffffffffu
#define VKL_MAX_INTERVAL_ITERATOR_SIZE 64
#define VKL_MAX_HIT_ITERATOR_SIZE 64

// what the handles point to (VKLObjectHandle::host)
struct RefVklVolume { const OcVolume *v; };
typedef RefVklVolume RefVklSampler;
typedef RefVklVolume RefVklIntervalContext;

float vklComputeSample(const VKLSampler *sampler, const vkl_vec3f *objectCoordinates, unsigned int attributeIndex = 0, float time = 0.f,
    VKLFeatureFlags featureFlags = VKL_FEATURE_FLAGS_DEFAULT);
vkl_vec3f vklComputeGradient(const VKLSampler *sampler, const vkl_vec3f *objectCoordinates, unsigned int attributeIndex = 0,
    float time = 0.f, VKLFeatureFlags featureFlags = VKL_FEATURE_FLAGS_DEFAULT);
VKLIntervalIterator vklInitIntervalIterator(const VKLIntervalIteratorContext *context, const vkl_vec3f *origin, const vkl_vec3f *direction,
    const vkl_range1f *tRange, float time, void *buffer, VKLFeatureFlags featureFlags = VKL_FEATURE_FLAGS_DEFAULT);
inline VKLIntervalIterator vklInitIntervalIterator(const VKLIntervalIteratorContext *context, const vkl_vec3f *origin,
    const vkl_vec3f *direction, const vkl_range1f *tRange, const float *time, void *buffer,
    VKLFeatureFlags featureFlags = VKL_FEATURE_FLAGS_DEFAULT)
{
  return vklInitIntervalIterator(context, origin, direction, tRange, time ? *time : 0.f, buffer, featureFlags);
}
int vklIterateInterval(VKLIntervalIterator iterator, VKLInterval *interval, VKLFeatureFlags featureFlags = VKL_FEATURE_FLAGS_DEFAULT);
inline size_t vklGetIntervalIteratorSize(const VKLIntervalIteratorContext *) { return VKL_MAX_INTERVAL_ITERATOR_SIZE; }
#define vklGetIntervalIteratorSizeV vklGetIntervalIteratorSize
// iso-surface hit iterators: declared for the isosurface geometry kernel, never hit by the tested scenes
inline VKLHitIterator vklInitHitIterator(const VKLHitIteratorContext *, const vkl_vec3f *, const vkl_vec3f *, const vkl_range1f *, float,
    void *buffer, VKLFeatureFlags = VKL_FEATURE_FLAGS_DEFAULT) { return buffer; }
inline int vklIterateHit(VKLHitIterator, VKLHit *, VKLFeatureFlags = VKL_FEATURE_FLAGS_DEFAULT) { return 0; }
