// CudaDevice : api::Device -- device type "cuda", registered by ospray_module_init_cuda.
// reference (behavioural template): modules/cpu/ISPCDevice.h:35-194, ISPCDevice.cpp:207-600,
// modules/cpu/ModuleInit.cpp:24-54; tile partition parameters mirror
// modules/multiDevice/MultiDevice.cpp:32-57 (numSubdevices) for one-process-per-GPU launches.
#pragma once
#include "../api/Device.h"
#include "Objects.h"

namespace ospray {
namespace cuda {

struct CudaDevice : public api::Device
{
  CudaDevice() = default;
  ~CudaDevice() override;
  void commit() override;

  int loadModule(const char *name) override;
  OSPData newSharedData(const void *sharedData, OSPDataType, const vec3ul &numItems, const vec3l &byteStride, OSPDeleterCallback,
      const void *userPtr) override;
  OSPData newData(OSPDataType, const vec3ul &numItems) override;
  void copyData(const OSPData source, OSPData destination, const vec3ul &destinationIndex) override;
  OSPLight newLight(const char *type) override;
  OSPCamera newCamera(const char *type) override;
  OSPGeometry newGeometry(const char *type) override;
  OSPVolume newVolume(const char *type) override;
  OSPGeometricModel newGeometricModel(OSPGeometry geom) override;
  OSPVolumetricModel newVolumetricModel(OSPVolume volume) override;
  OSPMaterial newMaterial(const char *material_type) override;
  OSPTransferFunction newTransferFunction(const char *type) override;
  OSPTexture newTexture(const char *type) override;
  OSPGroup newGroup() override;
  OSPInstance newInstance(OSPGroup group) override;
  OSPWorld newWorld() override;
  box3f getBounds(OSPObject) override;
  void setObjectParam(OSPObject object, const char *name, OSPDataType type, const void *mem) override;
  void removeObjectParam(OSPObject object, const char *name) override;
  void commit(OSPObject object) override;
  void release(OSPObject obj) override;
  void retain(OSPObject obj) override;
  OSPFrameBuffer frameBufferCreate(const vec2i &size, const OSPFrameBufferFormat mode, const uint32_t channels) override;
  OSPImageOperation newImageOp(const char *type) override;
  const void *frameBufferMap(OSPFrameBuffer fb, const OSPFrameBufferChannel) override;
  void frameBufferUnmap(const void *mapped, OSPFrameBuffer fb) override;
  float getVariance(OSPFrameBuffer) override;
  OSPPickResult pick(OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld, const vec2f &screenPos) override;
  void resetAccumulation(OSPFrameBuffer fb) override;
  OSPRenderer newRenderer(const char *type) override;
  OSPFuture renderFrame(OSPFrameBuffer, OSPRenderer, OSPCamera, OSPWorld) override;
  int isReady(OSPFuture, OSPSyncEvent) override;
  void wait(OSPFuture, OSPSyncEvent) override;
  void cancel(OSPFuture) override;
  float getProgress(OSPFuture) override;
  float getTaskDuration(OSPFuture) override;

  // device-global material table: slots are stable for the life of a Material object
  uint32_t allocMaterialSlot();
  void freeMaterialSlot(uint32_t slot);
  void setMaterial(uint32_t slot, const OcMaterial &m);
  const OcMaterial *materialTable(); // uploads when dirty

  int cudaOrdinal = 0;
  uint32_t numRanks = 1, rank = 0; // image-tile partition (tile % numRanks == rank)
  bool initialized = false;
  bool disableMipMapGeneration = false; // device parameter of the same name (ISPCDevice.cpp:232-233)
  uint32_t texturedMaterials = 0;       // materials whose last commit referenced a texture
  uint32_t emissiveMaterials = 0;       // materials whose last commit made them emit
  uint64_t materialVersion = 0;         // bumped by every material commit

 private:
  std::vector<OcMaterial> materials;
 public:
  const std::vector<OcMaterial> &materialTableHost() const { return materials; }
 private:
  std::vector<uint32_t> freeSlots;
  DevMem materialsDev;
  bool materialsDirty = true;
};

void registerMultiDevice();                          // MultiDevice.cpp
ManagedObject *subObject(void *handle, size_t i);    // sub-object i of a multidevice handle; the handle itself for i == 0 otherwise

} // namespace cuda
} // namespace ospray
