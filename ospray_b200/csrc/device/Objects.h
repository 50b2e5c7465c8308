// Host-side object graph of the `cuda` device: one ManagedObject subclass per API object kind.
// On commit() each object fills the POD struct of include/ospcuda.h that the kernels read.
// reference (behavioural template): modules/cpu/common/{Data,Group,Instance,World}.cpp,
// camera/{Camera,PerspectiveCamera}.cpp, geometry/{Mesh,Spheres,Boxes,GeometricModel}.cpp,
// lights/{Light,AmbientLight,DirectionalLight,QuadLight}.cpp, render/{Renderer}.cpp,
// render/pathtracer/PathTracer.cpp, render/scivis/{SciVis,SciVisData}.cpp, render/ao/AORenderer.cpp,
// render/materials/OBJ.cpp, volume/{Volume,VolumetricModel}.cpp,
// volume/transferFunction/LinearTransferFunction.cpp, fb/{FrameBuffer,LocalFB}.cpp,
// fb/frame_ops/{ToneMapper,Variance}.cpp, render/RenderingFuture.h.
#pragma once
#include <memory>
#include <random>
#include <set>
#include <map>
#include "../common/Managed.h"
#include "ospcuda.h"

namespace ospray {
namespace cuda {

struct CudaDevice;
struct Renderer;

void checkCuda(int rc, const char *what); // throws std::runtime_error with ospcuda_last_error()

inline OcVec3f toOc(const vec3f &v) { OcVec3f r; r.x = v.x; r.y = v.y; r.z = v.z; return r; }
inline OcAffine3f toOc(const affine3f &a)
{
  OcAffine3f r; r.vx = toOc(a.l.vx); r.vy = toOc(a.l.vy); r.vz = toOc(a.l.vz); r.p = toOc(a.p); return r;
}

// device allocation with value semantics disabled
struct DevMem
{
  void *p = nullptr;
  size_t bytes = 0;
  DevMem() = default;
  DevMem(const DevMem &) = delete;
  DevMem &operator=(const DevMem &) = delete;
  ~DevMem() { release(); }
  void release();
  void reserve(size_t n);                    // (re)allocate when n > bytes
  void upload(const void *src, size_t n);    // reserve + H2D
  template <typename T> T *as() const { return (T *)p; }
};

struct Data : public ManagedObject
{
  Data(const void *sharedData, OSPDataType, const vec3ul &numItems, const vec3l &byteStride, OSPDeleterCallback, const void *userPtr);
  Data(OSPDataType, const vec3ul &numItems);
  ~Data() override;
  std::string toString() const override { return "ospray::Data"; }
  void commit() override { deviceDirty = true; }
  void copy(const Data &source, const vec3ul &destinationIndex);
  size_t size() const { return numItems.x * numItems.y * numItems.z; }
  char *data(const vec3ul &idx) const { return addr + idx.x * byteStride.x + idx.y * byteStride.y + idx.z * byteStride.z; }
  char *data(size_t i) const { return addr + (int64_t)i * stride1D; } // 1-D view along the dominant axis
  bool compact() const;
  template <typename T> T &at(size_t i) const { return *(T *)data(i); }
  ManagedObject *objectAt(size_t i) const { return *(ManagedObject **)data(i); }
  // device mirrors (uploaded lazily, refreshed after ospCommit(data))
  OcData1D deviceView1D();
  const void *deviceCompact(); // tightly packed copy of all items (x fastest)

  char *addr = nullptr;
  bool shared = false;
  OSPDataType type;
  vec3ul numItems;
  vec3l byteStride;
  int dimensions = 1;
  int64_t stride1D = 0;
  size_t numItems1D = 0;

 private:
  void init();
  std::vector<char> owned;
  OSPDeleterCallback freeFunction = nullptr;
  const void *userData = nullptr;
  DevMem dev;
  bool deviceDirty = true;
  bool devIsCompact = false;
};

// fetch a data parameter of a given element type; `promoteScalar` wraps a single value/object
// (ospray/common/Managed.h:60-123 getParamDataT)
Data *getParamData(ManagedObject &o, const std::string &name, OSPDataType elemType, bool required = false,
    bool promoteScalar = false, Ref<Data> *promoted = nullptr);

struct Camera : public ManagedObject
{
  explicit Camera(const std::string &type);
  std::string toString() const override { return type == "orthographic" ? "ospray::OrthographicCamera" : type == "panoramic" ? "ospray::PanoramicCamera" : "ospray::PerspectiveCamera"; }
  void commit() override;
  std::string type;
  OcCamera sh{};
};

struct Geometry : public ManagedObject
{
  explicit Geometry(const std::string &type);
  std::string toString() const override { return "ospray::Geometry(" + type + ")"; }
  void commit() override;
  uint32_t numPrimitives() const { return sh.numPrims; }
  const OcGeometry *devicePtr() const { return dev.as<OcGeometry>(); }
  // geometry lights: meshes and spheres can be sampled by area (Geometry::supportAreaLighting); world-space areas of the
  // given primitives under `xfm` (Mesh_getAreas geometry/Mesh.ispc:385-420, Spheres_getAreas geometry/Spheres.ispc:76-104)
  bool supportAreaLighting() const { return type == "mesh" || type == "sphere"; }
  void getAreas(const std::vector<int32_t> &primIDs, const affine3f &xfm, std::vector<float> &areas) const;
  std::string type;
  OcGeometry sh{};

 private:
  Ref<Data> vertexData, indexData, normalData, colorData, texcoordData, sphereData, radiusData, boxData;
  float sphereRadius = 0.01f;
  DevMem dev, generated; // generated: soup index / interleaved spheres
};

// texture2d (texture/Texture2D.cpp:56-146): level 0 = the app's 2D data array, MIP levels generated at commit
// (texture/MipMapGeneration.ispc:8-95) unless the device parameter disableMipMapGeneration is set
struct Texture2D : public ManagedObject
{
  explicit Texture2D(CudaDevice &d) : device(d) { managedObjectType = OSP_TEXTURE; }
  std::string toString() const override { return "ospray::Texture2D"; }
  void commit() override;
  const OcTexture2D *devPtr() const { return shDev.as<OcTexture2D>(); }
  CudaDevice &device;
  OcTexture2D sh{};

 private:
  Ref<Data> texData;
  DevMem texels, shDev; // all levels back to back; the OcTexture2D itself
};

struct Material : public ManagedObject
{
  Material(CudaDevice &device, const std::string &type);
  ~Material() override;
  std::string toString() const override { return "ospray::Material(" + type + ")"; }
  void commit() override;
  CudaDevice &device;
  std::string type;
  uint32_t slot;

 private:
  bool wasEmissive = false, wasExtended = false;
  OcTextureParam getTextureParam(const std::string &name, Ref<Texture2D> &keep); // render/Material.cpp:73-165
  Ref<Texture2D> texD, texKd, texKs, texNs, texBump;
  std::vector<Ref<Texture2D>> texPrincipled; // the principled material's maps (OC_PRMAP_*)
  DevMem principledMaps;                     // OcPrincipledMaps / OcCarPaint on the device
};

uint64_t commitEpoch();     // counts the commits of objects whose state parents snapshot
uint64_t nextCommitEpoch();

struct GeometricModel : public ManagedObject
{
  explicit GeometricModel(Geometry *g);
  std::string toString() const override { return "ospray::GeometricModel"; }
  void commit() override;
  Ref<Geometry> geomAPI, geom;
  OcModel sh{};
  uint64_t epoch = 0; // commitEpoch() at the last commit (parents compare it with the epoch of their snapshot)
  // material table slot of primitive `primID` (GeometricModel_getMaterial, geometry/GeometricModel.ih:21-46), -1 = none;
  // rendererSlots = the renderer's material list for models that index it
  int materialSlotOf(uint32_t primID, const std::vector<uint32_t> &rendererSlots) const;
  std::vector<uint32_t> hostMaterials; // what sh.material holds (slots or renderer indices)

 private:
  Ref<Data> materialData, colorData, indexData, promotedMat, promotedColor;
  DevMem matList;
};

struct TransferFunction : public ManagedObject
{
  std::string toString() const override { return "ospray::LinearTransferFunction"; }
  void commit() override;
  std::vector<box1f> positiveOpacityValueRanges() const;
  OcTransferFunction sh{};
  box1f valueRange{0.f, 1.f};
  uint64_t epoch = 0;

 private:
  Ref<Data> colorData, opacityData, promC, promO;
  std::vector<float> opacities;
};

struct Volume : public ManagedObject
{
  explicit Volume(const std::string &type);
  std::string toString() const override { return "ospray::Volume(" + type + ")"; }
  void commit() override;
  box3f getBounds() const override { return bounds; }
  std::string type;
  box3f bounds = box_empty();
  vec3f gridOrigin{0, 0, 0}, gridSpacing{1, 1, 1};
  uint32_t dims[3] = {0, 0, 0}, mcDims[3] = {0, 0, 0};
  DevMem voxels, mcRange;

 private:
  Ref<Data> data;
};

struct VolumetricModel : public ManagedObject
{
  explicit VolumetricModel(Volume *v);
  std::string toString() const override { return "ospray::VolumetricModel"; }
  void commit() override;
  box3f getBounds() const override { return volume ? volume->getBounds() : box_empty(); }
  Ref<Volume> volumeAPI, volume;
  Ref<TransferFunction> tf;
  OcVolume sh{};
  uint64_t epoch = 0;
  bool refreshFromChildren(); // the transfer function was committed after this model: take its new state over
};

struct Light : public ManagedObject
{
  explicit Light(const std::string &type);
  std::string toString() const override { return "ospray::Light(" + type + ")"; }
  void commit() override;
  std::string type;
  OcLight sh{};
  vec3f ambientRadiance{0, 0, 0};

 private:
  Ref<Texture2D> hdriMap;      // hdri: the environment map and the CDFs of its importance distribution
  DevMem hdriCdfX, hdriCdfY;
};

struct Group : public ManagedObject
{
  ~Group() override;
  std::string toString() const override { return "ospray::Group"; }
  void commit() override;
  box3f getBounds() const override { return bounds; }
  const OcGroup *devicePtr() const { return dev.as<OcGroup>(); }
  bool hasVolumes() const { return sh.numVolumes > 0; }
  Ref<Data> geometricModels, volumetricModels, lights;
  OcGroup sh{};
  box3f bounds = box_empty();
  float lastBuildMs = 0.f;
  uint32_t stackNeed = 0; // deepest traversal stack the group's BVH can ask for (ospcuda_bvh_last_stack_need)
  // The device arrays of this group are SNAPSHOTS of its models' structs (the reference's Group keeps pointers to them,
  // Group.cpp:139/157). A model or transfer function committed later -- new material list, colours, ids, opacities -- must
  // show in the next frame without the group being committed again (OSPRay semantics), and a snapshot must never keep a
  // pointer into a buffer the child has since replaced: World::prepareFrame calls this whenever anything was committed
  // since the last frame. Geometry changes still need the group's commit (the hierarchy is built there).
  void refreshSnapshots();
  uint64_t snapshotEpoch = 0;

 private:
  DevMem dev, modelsDev, volumesDev;
};

struct Instance : public ManagedObject
{
  explicit Instance(Group *g);
  std::string toString() const override { return "ospray::Instance"; }
  void commit() override;
  box3f getBounds() const override;
  Ref<Group> groupAPI, group;
  affine3f xfm = affine_identity(), rcp_xfm = affine_identity();
  uint32_t userID = 0xFFFFFFFFu;
  bool identity = true;
};

struct World : public ManagedObject
{
  explicit World(CudaDevice &d) : device(d) { managedObjectType = OSP_WORLD; }
  ~World() override;
  std::string toString() const override { return "ospray::World"; }
  void commit() override;
  box3f getBounds() const override { return bounds; }
  // refresh material table pointer, upload the struct when dirty; for a pathtracer with geometryLights the emissive
  // geometry is scanned and its proxy lights put in front of the light list (PathTracerData.cpp:32-160)
  const OcWorld *prepareFrame(const Renderer *renderer = nullptr);
  CudaDevice &device;
  Ref<Data> instances, lights;
  OcWorld sh{};
  bool hasHdriLight = false;   // frames of this world run the kernels with texture support
  box3f bounds = box_empty();

 private:
  DevMem dev, instDev, lightsDev, cdfDev, scivisLightsDev;
  // geometry lights of the last frame that asked for them
  std::vector<OcInstance> hostInstances;
  std::vector<OcLight> hostLights;
  std::vector<std::unique_ptr<DevMem>> geoBuffers;
  DevMem geoLightsDev, geoCdfDev, geoInstDev;
  uint64_t geoKey[3] = {~0ull, ~0ull, ~0ull};
  bool geoActive = false, geoInUse = false;

  uint32_t geoNumLights = 0, geoNumGeo = 0;
  void buildGeometryLights(const Renderer &renderer);
  const OcInstance *instDevGeo();
  uint32_t computeShadowOpaque(const Renderer *renderer) const;
  const void *uploadedMaterials = nullptr;
  bool dirty = true;
  uint64_t seenEpoch = 0;
};

struct ImageOp : public ManagedObject
{
  explicit ImageOp(const std::string &type) : type(type) { managedObjectType = OSP_IMAGE_OPERATION; }
  std::string toString() const override { return "ospray::ImageOp(" + type + ")"; }
  void commit() override;
  std::string type;
  OcToneMapper tm{};
};

struct Renderer : public ManagedObject
{
  Renderer(CudaDevice &d, const std::string &type);
  std::string toString() const override { return "ospray::Renderer(" + type + ")"; }
  void commit() override;
  CudaDevice &device;
  std::string type;
  OcRenderer sh{};
  float errorThreshold = 0.f;
  std::vector<uint32_t> materialSlots; // the renderer's material list (material table slots)
  uint64_t commitCount = 0;

 private:
  Ref<Data> materialData;
  Ref<Texture2D> backplate, maxDepthTexture;
  DevMem pfCdfX, pfCdfY, matList;
};

struct FrameBuffer : public ManagedObject
{
  FrameBuffer(CudaDevice &d, const vec2i &size, OSPFrameBufferFormat format, uint32_t channels);
  ~FrameBuffer() override;
  std::string toString() const override { return "ospray::LocalFrameBuffer"; }
  void commit() override;
  void clear();        // ospResetAccumulation (FrameBuffer.cpp:95-103)
  void beginFrame();   // FrameBuffer.cpp:115-137
  void postProcess();  // LocalFB.cpp:459-478 (enqueue on the render stream)
  const void *map(OSPFrameBufferChannel channel);
  void unmap(const void *mapped);
  float getVariance();
  void *devicePointer(OSPFrameBufferChannel channel) const;
  size_t bytesPerPixel(OSPFrameBufferChannel channel) const;

  CudaDevice &device;
  vec2i size;
  OSPFrameBufferFormat colorFormat;
  uint32_t channels;
  bool doAccum, hasVariance;
  OcFrameBuffer sh{};
  unsigned long long rayCounts[4] = {0, 0, 0, 0}; // last finished frame (host); [3]: traced by the PT fused tail kernel
  float errorThreshold = 0.f;

 private:
  std::vector<std::unique_ptr<DevMem>> buffers;
  DevMem *colorBuf = nullptr, *varianceBuf = nullptr, *depthBuf = nullptr, *normalBuf = nullptr, *firstNormalBuf = nullptr,
         *albedoBuf = nullptr, *positionBuf = nullptr, *primBuf = nullptr, *objBuf = nullptr, *instBuf = nullptr, *convBuf = nullptr,
         *ppBuf = nullptr, *ppStore = nullptr, *taskErrBuf = nullptr, *counterBuf = nullptr;
  Ref<Data> imageOpData;
  std::vector<Ref<ImageOp>> toneMappers;
  struct Staging { void *host; bool inUse; };
  std::map<int, Staging> staging;  // pinned read-back buffer per channel, reused across maps
  std::set<void *> extraMapped;    // one-off buffers when a channel is mapped twice
  std::vector<float> taskErrors;
  bool taskErrorsValid = false, taskErrorsPending = false;
  float frameVariance = INFINITY;
  std::mt19937 mtGen{43};
  bool pickOdd = false;
  friend struct Future;
  friend struct CudaDevice;
  friend struct MultiCudaDevice;
};

struct Future : public ManagedObject
{
  Future(FrameBuffer *fb, Renderer *r, Camera *c, World *w);
  ~Future() override; // blocks until the frame is done (RenderingFuture.h:46-53)
  std::string toString() const override { return "ospray::RenderingFuture"; }
  bool isFinished(OSPSyncEvent);
  void wait(OSPSyncEvent);
  float duration();
  void finalizeHost(); // after the finished event: fetch ray counters / task errors
  Ref<FrameBuffer> fb;
  Ref<Renderer> renderer;
  Ref<Camera> camera;
  Ref<World> world;
  void *evStart = nullptr, *evRendered = nullptr, *evFinished = nullptr;
  bool finalized = false;
};

} // namespace cuda
} // namespace ospray
