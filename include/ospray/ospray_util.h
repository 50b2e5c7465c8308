/* ospray-cuda: convenience wrappers of the OSPRay C API, restated
 * (reference: ospray/include/ospray/ospray_util.h:12-95, implemented in
 * ospray/api/ospray_util_impl.cpp:7-192). All of them are thin compositions of
 * the 51 core entry points in ospray.h.                                        */
#pragma once
#include "ospray.h"
#ifdef __cplusplus
extern "C" {
#endif
OSPRAY_INTERFACE OSPData ospNewSharedData1D(const void *sharedData, OSPDataType, uint64_t numItems);
OSPRAY_INTERFACE OSPData ospNewAssignedData1D(const void *sharedData, OSPDataType, uint64_t numItems,
    OSPDeleterCallback, const void *userPtr);
OSPRAY_INTERFACE OSPData ospNewSharedData1DStride(const void *sharedData, OSPDataType, uint64_t numItems, int64_t byteStride);
OSPRAY_INTERFACE OSPData ospNewSharedData2D(const void *sharedData, OSPDataType, uint64_t numItems1, uint64_t numItems2);
OSPRAY_INTERFACE OSPData ospNewSharedData2DStride(const void *sharedData, OSPDataType, uint64_t numItems1,
    int64_t byteStride1, uint64_t numItems2, int64_t byteStride2);
OSPRAY_INTERFACE OSPData ospNewSharedData3D(const void *sharedData, OSPDataType, uint64_t numItems1,
    uint64_t numItems2, uint64_t numItems3);
OSPRAY_INTERFACE OSPData ospNewData1D(OSPDataType, uint64_t numItems);
OSPRAY_INTERFACE OSPData ospNewData2D(OSPDataType, uint64_t numItems1, uint64_t numItems2);
OSPRAY_INTERFACE void ospCopyData1D(const OSPData source, OSPData destination, uint64_t destinationIndex);
OSPRAY_INTERFACE void ospCopyData2D(const OSPData source, OSPData destination, uint64_t destinationIndex1, uint64_t destinationIndex2);

OSPRAY_INTERFACE void ospSetString(OSPObject, const char *n, const char *s);
OSPRAY_INTERFACE void ospSetObject(OSPObject, const char *n, OSPObject obj);
OSPRAY_INTERFACE void ospSetBool(OSPObject, const char *n, int x);
OSPRAY_INTERFACE void ospSetFloat(OSPObject, const char *n, float x);
OSPRAY_INTERFACE void ospSetInt(OSPObject, const char *n, int x);
OSPRAY_INTERFACE void ospSetUInt(OSPObject, const char *n, unsigned int x);
OSPRAY_INTERFACE void ospSetBox1f(OSPObject, const char *n, float x, float y);
OSPRAY_INTERFACE void ospSetLinear2f(OSPObject, const char *n, float x, float y, float z, float w);
OSPRAY_INTERFACE void ospSetVec2f(OSPObject, const char *n, float x, float y);
OSPRAY_INTERFACE void ospSetVec3f(OSPObject, const char *n, float x, float y, float z);
OSPRAY_INTERFACE void ospSetVec4f(OSPObject, const char *n, float x, float y, float z, float w);
OSPRAY_INTERFACE void ospSetVec2i(OSPObject, const char *n, int x, int y);
OSPRAY_INTERFACE void ospSetVec3i(OSPObject, const char *n, int x, int y, int z);
OSPRAY_INTERFACE void ospSetVec4i(OSPObject, const char *n, int x, int y, int z, int w);
OSPRAY_INTERFACE void ospSetObjectAsData(OSPObject target, const char *n, OSPDataType, OSPObject obj);
OSPRAY_INTERFACE float ospRenderFrameBlocking(OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld);
#ifdef __cplusplus
}
#endif
