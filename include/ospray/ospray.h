/* ospray-cuda: the OSPRay 3.3 public C API (51 entry points), restated.
 *
 * This is the application-facing boundary: signatures and handle semantics equal
 * the reference's ospray/include/ospray/ospray.h:151-346 so that ospTutorial,
 * ospBenchmark-style drivers and the header-only C++ wrappers compile and link
 * unchanged against our libospray.so + libospray_module_cuda.so.               */
#pragma once
#ifdef __cplusplus
#include <cstddef>
#include <cstdint>
#else
#include <stddef.h>
#include <stdint.h>
#endif
#include "OSPEnums.h"

#define OSPRAY_VERSION_MAJOR 3
#define OSPRAY_VERSION_MINOR 3
#define OSPRAY_VERSION_PATCH 0
#define OSPRAY_SOVERSION 3
#define OSPRAY_VERSION 30300

#ifndef OSPRAY_INTERFACE
#define OSPRAY_INTERFACE __attribute__((visibility("default")))
#endif
#ifdef __cplusplus
#define OSP_DEFAULT_VAL(a) = a
#else
#define OSP_DEFAULT_VAL(a)
#endif

#ifdef __cplusplus
/* typed opaque handles so that C++ callers get overload checking */
namespace osp {
struct Device;
struct ManagedObject {};
#define OSP_HANDLE_KIND(K) struct K : public ManagedObject {};
OSP_HANDLE_KIND(Camera) OSP_HANDLE_KIND(Data) OSP_HANDLE_KIND(FrameBuffer) OSP_HANDLE_KIND(Future)
OSP_HANDLE_KIND(GeometricModel) OSP_HANDLE_KIND(Geometry) OSP_HANDLE_KIND(Group)
OSP_HANDLE_KIND(ImageOperation) OSP_HANDLE_KIND(Instance) OSP_HANDLE_KIND(Light)
OSP_HANDLE_KIND(Material) OSP_HANDLE_KIND(Renderer) OSP_HANDLE_KIND(Texture)
OSP_HANDLE_KIND(TransferFunction) OSP_HANDLE_KIND(Volume) OSP_HANDLE_KIND(VolumetricModel)
OSP_HANDLE_KIND(World)
#undef OSP_HANDLE_KIND
} // namespace osp
typedef osp::Device *OSPDevice;
typedef osp::ManagedObject *OSPObject;
typedef osp::Camera *OSPCamera;
typedef osp::Data *OSPData;
typedef osp::FrameBuffer *OSPFrameBuffer;
typedef osp::Future *OSPFuture;
typedef osp::GeometricModel *OSPGeometricModel;
typedef osp::Geometry *OSPGeometry;
typedef osp::Group *OSPGroup;
typedef osp::ImageOperation *OSPImageOperation;
typedef osp::Instance *OSPInstance;
typedef osp::Light *OSPLight;
typedef osp::Material *OSPMaterial;
typedef osp::Renderer *OSPRenderer;
typedef osp::Texture *OSPTexture;
typedef osp::TransferFunction *OSPTransferFunction;
typedef osp::Volume *OSPVolume;
typedef osp::VolumetricModel *OSPVolumetricModel;
typedef osp::World *OSPWorld;
extern "C" {
#else
typedef void *OSPDevice;
typedef void *OSPObject, *OSPManagedObject, *OSPCamera, *OSPData, *OSPFrameBuffer, *OSPFuture,
    *OSPGeometricModel, *OSPGeometry, *OSPGroup, *OSPImageOperation, *OSPInstance, *OSPLight,
    *OSPMaterial, *OSPRenderer, *OSPTexture, *OSPTransferFunction, *OSPVolume, *OSPVolumetricModel,
    *OSPWorld;
#endif

typedef void (*OSPStatusCallback)(void *userData, const char *messageText);
typedef void (*OSPErrorCallback)(void *userData, OSPError, const char *errorDetails);
typedef void (*OSPDeleterCallback)(const void *userData, const void *sharedData);
typedef struct { float lower[3]; float upper[3]; } OSPBounds;
typedef struct {
  int hasHit;
  float worldPosition[3];
  OSPInstance instance;
  OSPGeometricModel model;
  uint32_t primID;
} OSPPickResult;

/* -- device lifetime ---------------------------------------------------------- */
OSPRAY_INTERFACE OSPError ospInit(int *argc OSP_DEFAULT_VAL(NULL), const char **argv OSP_DEFAULT_VAL(NULL));
OSPRAY_INTERFACE int64_t ospDeviceGetProperty(OSPDevice, OSPDeviceProperty);
OSPRAY_INTERFACE void ospShutdown(void);
OSPRAY_INTERFACE OSPDevice ospNewDevice(const char *deviceType OSP_DEFAULT_VAL("cpu"));
OSPRAY_INTERFACE void ospSetCurrentDevice(OSPDevice);
OSPRAY_INTERFACE OSPDevice ospGetCurrentDevice(void);
OSPRAY_INTERFACE void ospDeviceSetParam(OSPDevice, const char *id, OSPDataType, const void *mem);
OSPRAY_INTERFACE void ospDeviceRemoveParam(OSPDevice, const char *id);
OSPRAY_INTERFACE void ospDeviceSetStatusCallback(OSPDevice, OSPStatusCallback, void *userData);
OSPRAY_INTERFACE void ospDeviceSetErrorCallback(OSPDevice, OSPErrorCallback, void *userData);
OSPRAY_INTERFACE OSPError ospDeviceGetLastErrorCode(OSPDevice);
OSPRAY_INTERFACE const char *ospDeviceGetLastErrorMsg(OSPDevice);
OSPRAY_INTERFACE void ospDeviceCommit(OSPDevice);
OSPRAY_INTERFACE void ospDeviceRelease(OSPDevice);
OSPRAY_INTERFACE void ospDeviceRetain(OSPDevice);
OSPRAY_INTERFACE OSPError ospLoadModule(const char *name);

/* -- data arrays -------------------------------------------------------------- */
OSPRAY_INTERFACE OSPData ospNewSharedData(const void *sharedData, OSPDataType,
    uint64_t numItems1, int64_t byteStride1 OSP_DEFAULT_VAL(0),
    uint64_t numItems2 OSP_DEFAULT_VAL(1), int64_t byteStride2 OSP_DEFAULT_VAL(0),
    uint64_t numItems3 OSP_DEFAULT_VAL(1), int64_t byteStride3 OSP_DEFAULT_VAL(0),
    OSPDeleterCallback OSP_DEFAULT_VAL(NULL), const void *userPtr OSP_DEFAULT_VAL(NULL));
OSPRAY_INTERFACE OSPData ospNewData(OSPDataType, uint64_t numItems1,
    uint64_t numItems2 OSP_DEFAULT_VAL(1), uint64_t numItems3 OSP_DEFAULT_VAL(1));
OSPRAY_INTERFACE void ospCopyData(const OSPData source, OSPData destination,
    uint64_t destinationIndex1 OSP_DEFAULT_VAL(0), uint64_t destinationIndex2 OSP_DEFAULT_VAL(0),
    uint64_t destinationIndex3 OSP_DEFAULT_VAL(0));

/* -- scene objects ------------------------------------------------------------ */
OSPRAY_INTERFACE OSPLight ospNewLight(const char *type);
OSPRAY_INTERFACE OSPCamera ospNewCamera(const char *type);
OSPRAY_INTERFACE OSPGeometry ospNewGeometry(const char *type);
OSPRAY_INTERFACE OSPVolume ospNewVolume(const char *type);
OSPRAY_INTERFACE OSPGeometricModel ospNewGeometricModel(OSPGeometry OSP_DEFAULT_VAL(NULL));
OSPRAY_INTERFACE OSPVolumetricModel ospNewVolumetricModel(OSPVolume OSP_DEFAULT_VAL(NULL));
OSPRAY_INTERFACE OSPMaterial ospNewMaterial(const char *materialType);
OSPRAY_INTERFACE OSPTransferFunction ospNewTransferFunction(const char *type);
OSPRAY_INTERFACE OSPTexture ospNewTexture(const char *type);
OSPRAY_INTERFACE OSPGroup ospNewGroup(void);
OSPRAY_INTERFACE OSPInstance ospNewInstance(OSPGroup OSP_DEFAULT_VAL(NULL));
OSPRAY_INTERFACE OSPWorld ospNewWorld(void);
OSPRAY_INTERFACE OSPBounds ospGetBounds(OSPObject);

/* -- parameters and lifetime -------------------------------------------------- */
OSPRAY_INTERFACE void ospSetParam(OSPObject, const char *id, OSPDataType, const void *mem);
OSPRAY_INTERFACE void ospRemoveParam(OSPObject, const char *id);
OSPRAY_INTERFACE void ospCommit(OSPObject);
OSPRAY_INTERFACE void ospRelease(OSPObject);
OSPRAY_INTERFACE void ospRetain(OSPObject);

/* -- framebuffer -------------------------------------------------------------- */
OSPRAY_INTERFACE OSPFrameBuffer ospNewFrameBuffer(int size_x, int size_y,
    OSPFrameBufferFormat format OSP_DEFAULT_VAL(OSP_FB_SRGBA),
    uint32_t frameBufferChannels OSP_DEFAULT_VAL(OSP_FB_COLOR));
OSPRAY_INTERFACE OSPImageOperation ospNewImageOperation(const char *type);
OSPRAY_INTERFACE const void *ospMapFrameBuffer(OSPFrameBuffer, OSPFrameBufferChannel OSP_DEFAULT_VAL(OSP_FB_COLOR));
OSPRAY_INTERFACE void ospUnmapFrameBuffer(const void *mapped, OSPFrameBuffer);
OSPRAY_INTERFACE float ospGetVariance(OSPFrameBuffer);
OSPRAY_INTERFACE void ospResetAccumulation(OSPFrameBuffer);

/* -- rendering ---------------------------------------------------------------- */
OSPRAY_INTERFACE OSPRenderer ospNewRenderer(const char *type);
OSPRAY_INTERFACE OSPFuture ospRenderFrame(OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld);
OSPRAY_INTERFACE int ospIsReady(OSPFuture, OSPSyncEvent OSP_DEFAULT_VAL(OSP_TASK_FINISHED));
OSPRAY_INTERFACE void ospWait(OSPFuture, OSPSyncEvent OSP_DEFAULT_VAL(OSP_TASK_FINISHED));
OSPRAY_INTERFACE void ospCancel(OSPFuture);
OSPRAY_INTERFACE float ospGetProgress(OSPFuture);
OSPRAY_INTERFACE float ospGetTaskDuration(OSPFuture);
OSPRAY_INTERFACE void ospPick(OSPPickResult *result, OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld,
    float screenPos_x, float screenPos_y);

#ifdef __cplusplus
} // extern "C"
#endif
