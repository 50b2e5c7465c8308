/* Extensions of the `cuda` OSPRay device that are not part of the OSPRay API: device-pointer
 * access for zero-copy consumers (the NCCL tile gather in ospray_b200/multigpu.py), ray
 * statistics for the Mrays/s metric, and BVH build timing. Exported by
 * libospray_module_cuda.so; the handles are ordinary OSPRay handles of the cuda device.    */
#ifndef OSPRAY_CUDA_EXT_H
#define OSPRAY_CUDA_EXT_H
#include "ospray/ospray.h"
#ifdef __cplusplus
extern "C" {
#endif
/* device pointer of a framebuffer channel (OSP_FB_COLOR = the converted colour buffer that
 * ospMapFrameBuffer would copy; OSP_FB_ACCUM = the float4 accumulation buffer); NULL if absent */
OSPRAY_INTERFACE void *ospcudaFrameBufferDevicePointer(OSPFrameBuffer, OSPFrameBufferChannel, uint64_t *bytesPerPixel);
/* rays handed to traversal during the last finished frame: [0] primary, [1] continuation /
 * transparency, [2] shadow + occlusion segments */
OSPRAY_INTERFACE void ospcudaGetRayCounts(OSPFrameBuffer, uint64_t counts[3]);
/* of those, the rays the pathtracer's fused tail kernel traced in the last finished frame */
OSPRAY_INTERFACE uint64_t ospcudaGetFusedRayCount(OSPFrameBuffer);
/* milliseconds spent in the last BVH build of a group */
OSPRAY_INTERFACE float ospcudaGetGroupBuildTime(OSPGroup);
/* number of kernels launched so far by this process through libospcuda */
OSPRAY_INTERFACE uint64_t ospcudaKernelLaunchCount(void);
/* cudaStream_t the device renders on (so that external consumers can order after a frame) */
OSPRAY_INTERFACE void *ospcudaRenderStream(void);
/* device time of a finished frame in ms: [0] the renderer kernel (incl. fused accumulation), [1] post-processing
 * (variance / tone mapping / colour conversion); CUDA events on the render stream */
OSPRAY_INTERFACE void ospcudaGetFrameStageTimes(OSPFuture, float ms[2]);
/* render on an external cudaStream_t (e.g. torch's current stream, so NCCL and event timing order with the
 * frame); NULL returns to the device's own stream */
OSPRAY_INTERFACE void ospcudaSetRenderStream(void *stream);
/* device pointer of the committed world's OcWorld (include/ospcuda.h), valid until the next commit;
 * lets a consumer call the batch traversal entry point ospcuda_trace directly */
OSPRAY_INTERFACE const void *ospcudaWorldDevicePointer(OSPWorld);
/* device pointer of a committed texture2d's OcTexture2D (include/ospcuda.h), valid until the next commit of the texture;
 * lets a consumer call the batch lookup entry point ospcuda_texture_sample directly */
OSPRAY_INTERFACE const void *ospcudaTextureDevicePointer(OSPTexture);
#ifdef __cplusplus
}
#endif
#endif
