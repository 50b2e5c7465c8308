// ospray::api::Device -- the abstract device every OSPRay module implements, restated.
// reference: ospray/api/Device.h:23-202 (same virtual names / argument meaning), Device.cpp:33-148.
// This base + the registry + the C API in API.cpp form our libospray.so; rkcommon is absent from
// the build image, so the reference's own libospray cannot be compiled here.
#pragma once
#include <functional>
#include <map>
#include <string>
#include "../common/Managed.h"

namespace ospray {
namespace api {

struct __attribute__((visibility("default"))) Device : public RefCounted, public ParameterizedObject
{
  using FactoryFcn = Device *(*)();
  template <typename D>
  static void registerType(const char *type) { registry()[type] = []() -> Device * { return new D; }; }
  static Device *createDevice(const char *type);
  static std::map<std::string, FactoryFcn> &registry();
  static Ref<Device> current;

  ~Device() override = default;
  virtual int64_t getProperty(const OSPDeviceProperty prop);

  virtual int loadModule(const char *name) = 0;
  virtual OSPData newSharedData(const void *sharedData, OSPDataType, const vec3ul &numItems, const vec3l &byteStride,
      OSPDeleterCallback, const void *userPtr) = 0;
  virtual OSPData newData(OSPDataType, const vec3ul &numItems) = 0;
  virtual void copyData(const OSPData source, OSPData destination, const vec3ul &destinationIndex) = 0;
  virtual OSPLight newLight(const char *type) = 0;
  virtual OSPCamera newCamera(const char *type) = 0;
  virtual OSPGeometry newGeometry(const char *type) = 0;
  virtual OSPVolume newVolume(const char *type) = 0;
  virtual OSPGeometricModel newGeometricModel(OSPGeometry geom) = 0;
  virtual OSPVolumetricModel newVolumetricModel(OSPVolume volume) = 0;
  virtual OSPMaterial newMaterial(const char *material_type) = 0;
  virtual OSPTransferFunction newTransferFunction(const char *type) = 0;
  virtual OSPTexture newTexture(const char *type) = 0;
  virtual OSPGroup newGroup() = 0;
  virtual OSPInstance newInstance(OSPGroup group) = 0;
  virtual OSPWorld newWorld() = 0;
  virtual box3f getBounds(OSPObject) = 0;
  virtual void setObjectParam(OSPObject object, const char *name, OSPDataType type, const void *mem) = 0;
  virtual void removeObjectParam(OSPObject object, const char *name) = 0;
  virtual void commit(OSPObject object) = 0;
  virtual void release(OSPObject obj) = 0;
  virtual void retain(OSPObject obj) = 0;
  virtual OSPFrameBuffer frameBufferCreate(const vec2i &size, const OSPFrameBufferFormat mode, const uint32_t channels) = 0;
  virtual OSPImageOperation newImageOp(const char *type) = 0;
  virtual const void *frameBufferMap(OSPFrameBuffer fb, const OSPFrameBufferChannel) = 0;
  virtual void frameBufferUnmap(const void *mapped, OSPFrameBuffer fb) = 0;
  virtual float getVariance(OSPFrameBuffer) = 0;
  virtual void resetAccumulation(OSPFrameBuffer fb) = 0;
  virtual OSPRenderer newRenderer(const char *type) = 0;
  virtual OSPFuture renderFrame(OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld) = 0;
  virtual int isReady(OSPFuture, OSPSyncEvent) = 0;
  virtual void wait(OSPFuture, OSPSyncEvent) = 0;
  virtual void cancel(OSPFuture) = 0;
  virtual float getProgress(OSPFuture) = 0;
  virtual float getTaskDuration(OSPFuture) = 0;
  virtual OSPPickResult pick(OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld, const vec2f &)
  {
    throw std::runtime_error("pick() not implemented for this device");
  }
  virtual void commit(); // device parameters + environment (Device.cpp:72-148)
  bool isCommitted() const { return committed; }

  int numThreads = -1;
  bool debugMode = false;
  static uint32_t logLevel;
  bool warningsAreErrors = false;
  std::function<void(void *, const char *)> msg_fcn = [](void *, const char *) {};
  void *statusUserData = nullptr;
  std::function<void(void *, OSPError, const char *)> error_fcn = [](void *, OSPError, const char *) {};
  void *errorUserData = nullptr;
  OSPError lastErrorCode = OSP_NO_ERROR;
  std::string lastErrorMsg = "no error";

 private:
  bool committed = false;
};

__attribute__((visibility("default"))) bool deviceIsSet();
__attribute__((visibility("default"))) Device &currentDevice();

} // namespace api

// OSPCommon.cpp:749-782: dlopen libospray_module_<name>.so and call ospray_module_init_<name>
__attribute__((visibility("default"))) OSPError loadLocalModule(const std::string &name);
// OSPCommon.h:126-133
inline OSPError moduleVersionCheck(int16_t versionMajor, int16_t versionMinor)
{
  if (versionMajor == OSPRAY_VERSION_MAJOR && versionMinor == OSPRAY_VERSION_MINOR) return OSP_NO_ERROR;
  return OSP_VERSION_MISMATCH;
}

} // namespace ospray
