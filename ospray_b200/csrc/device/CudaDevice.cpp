// CudaDevice implementation + module entry point. See CudaDevice.h for the reference map.
#include <algorithm>
#include "CudaDevice.h"
#include <cstdlib>
#include "ospray_cuda_ext.h"

namespace ospray {
namespace cuda {

CudaDevice::~CudaDevice() {}

void CudaDevice::commit()
{
  api::Device::commit();
  // which GPU: device param "cudaDevice" <- env OSPRAY_CUDA_DEVICE <- LOCAL_RANK (torchrun)
  int ord = getParam<int>("cudaDevice", -1);
  if (const char *e = getenv("OSPRAY_CUDA_DEVICE")) ord = atoi(e);
  if (ord < 0) {
    const char *lr = getenv("LOCAL_RANK");
    ord = lr ? atoi(lr) : 0;
    const int n = ospcuda_device_count();
    if (n > 0) ord %= n;
  }
  // image-tile partition, the multidevice analogue (MultiDevice.cpp:32-57: numSubdevices)
  int nr = getParam<int>("numSubdevices", getParam<int>("numRanks", 1));
  int rk = getParam<int>("subdeviceRank", getParam<int>("rank", 0));
  if (const char *e = getenv("OSPRAY_NUM_SUBDEVICES")) nr = atoi(e);
  if (const char *e = getenv("OSPRAY_SUBDEVICE_RANK")) rk = atoi(e);
  // a sub-device of the in-process multidevice (MultiDevice.cpp): its share is fixed by the parent, whatever the environment says
  if (hasParam("multiNumSubdevices")) {
    nr = getParam<int>("multiNumSubdevices", 1);
    rk = getParam<int>("multiSubdeviceRank", 0);
    ord = getParam<int>("cudaDevice", 0);
  }
  disableMipMapGeneration = getParam<bool>("disableMipMapGeneration", false);
  numRanks = (uint32_t)std::max(1, nr);
  rank = (uint32_t)std::max(0, rk) % numRanks;
  cudaOrdinal = ord;
  checkCuda(ospcuda_init(ord), "CUDA initialisation (the cuda device has no CPU fallback)");
  initialized = true;
  postStatusMsg(OSP_LOG_INFO, "cuda device: using CUDA device " + std::to_string(ord) + ", tile partition " + std::to_string(rank) + "/"
          + std::to_string(numRanks));
}

int CudaDevice::loadModule(const char *name) { return loadLocalModule(name); }

OSPData CudaDevice::newSharedData(const void *sharedData, OSPDataType t, const vec3ul &n, const vec3l &s, OSPDeleterCallback cb,
    const void *up)
{
  return (OSPData) new Data(sharedData, t, n, s, cb, up);
}
OSPData CudaDevice::newData(OSPDataType t, const vec3ul &n) { return (OSPData) new Data(t, n); }
void CudaDevice::copyData(const OSPData source, OSPData destination, const vec3ul &di)
{
  Data *dst = (Data *)destination;
  Data *src = (Data *)source;
  dst->copy(*src, di);
}

OSPLight CudaDevice::newLight(const char *type)
{
  const std::string t(type);
  if (t != "ambient" && t != "distant" && t != "quad" && t != "sphere" && t != "spot" && t != "hdri" && t != "cylinder") return nullptr;
  return (OSPLight) new Light(t);
}
OSPCamera CudaDevice::newCamera(const char *type)
{
  if (std::string(type) != "perspective" && std::string(type) != "orthographic" && std::string(type) != "panoramic") return nullptr;
  return (OSPCamera) new Camera(type);
}
OSPGeometry CudaDevice::newGeometry(const char *type)
{
  const std::string t(type);
  if (t != "mesh" && t != "sphere" && t != "box" && t != "curve") return nullptr;
  return (OSPGeometry) new Geometry(t);
}
OSPVolume CudaDevice::newVolume(const char *type)
{
  const std::string t(type);
  if (t != "structuredRegular" && t != "structured_regular") return nullptr;
  return (OSPVolume) new Volume(t);
}
OSPGeometricModel CudaDevice::newGeometricModel(OSPGeometry g) { return (OSPGeometricModel) new GeometricModel((Geometry *)g); }
OSPVolumetricModel CudaDevice::newVolumetricModel(OSPVolume v) { return (OSPVolumetricModel) new VolumetricModel((Volume *)v); }
OSPMaterial CudaDevice::newMaterial(const char *type) { return (OSPMaterial) new Material(*this, type); }
OSPTransferFunction CudaDevice::newTransferFunction(const char *type)
{
  if (std::string(type) != "piecewiseLinear") return nullptr;
  return (OSPTransferFunction) new TransferFunction;
}
OSPTexture CudaDevice::newTexture(const char *type)
{
  if (std::string(type) != "texture2d") return nullptr; // volume textures are not built
  return (OSPTexture) new Texture2D(*this);
}
OSPGroup CudaDevice::newGroup() { return (OSPGroup) new Group; }
OSPInstance CudaDevice::newInstance(OSPGroup g) { return (OSPInstance) new Instance((Group *)g); }
OSPWorld CudaDevice::newWorld() { return (OSPWorld) new World(*this); }

box3f CudaDevice::getBounds(OSPObject o)
{
  const box3f b = ((ManagedObject *)o)->getBounds();
  if (box_is_empty(b)) {
    // rkcommon's empty box: lower = +inf, upper = -inf
    return b;
  }
  return b;
}

void CudaDevice::setObjectParam(OSPObject object, const char *name, OSPDataType type, const void *mem)
{
  if (type == OSP_UNKNOWN) throw std::runtime_error("cannot set OSP_UNKNOWN parameter type");
  ((ManagedObject *)object)->setParam(name, type, mem);
}
void CudaDevice::removeObjectParam(OSPObject object, const char *name) { ((ManagedObject *)object)->removeParam(name); }
void CudaDevice::commit(OSPObject object)
{
  ManagedObject *o = (ManagedObject *)object;
  o->commit();
  o->checkUnused();
  o->resetAllParamQueryStatus();
}
void CudaDevice::release(OSPObject o) { ((ManagedObject *)o)->refDec(); }
void CudaDevice::retain(OSPObject o) { ((ManagedObject *)o)->refInc(); }

OSPFrameBuffer CudaDevice::frameBufferCreate(const vec2i &size, const OSPFrameBufferFormat mode, const uint32_t channels)
{
  return (OSPFrameBuffer) new FrameBuffer(*this, size, mode, channels);
}
OSPImageOperation CudaDevice::newImageOp(const char *type)
{
  if (std::string(type) != "tonemapper") return nullptr; // blur / debug / denoiser: out of scope
  return (OSPImageOperation) new ImageOp(type);
}
const void *CudaDevice::frameBufferMap(OSPFrameBuffer fb, const OSPFrameBufferChannel ch) { return ((FrameBuffer *)fb)->map(ch); }
void CudaDevice::frameBufferUnmap(const void *mapped, OSPFrameBuffer fb) { ((FrameBuffer *)fb)->unmap(mapped); }
float CudaDevice::getVariance(OSPFrameBuffer fb) { return ((FrameBuffer *)fb)->getVariance(); }
void CudaDevice::resetAccumulation(OSPFrameBuffer fb) { ((FrameBuffer *)fb)->clear(); }

OSPRenderer CudaDevice::newRenderer(const char *type)
{
  const std::string t(type);
  if (t != "pathtracer" && t != "scivis" && t != "ao" && t != "debug") return nullptr;
  return (OSPRenderer) new Renderer(*this, t);
}

// ISPCDevice.cpp:568-600 + LoadBalancer.cpp:13-32, enqueued on the render stream
OSPFuture CudaDevice::renderFrame(OSPFrameBuffer _fb, OSPRenderer _r, OSPCamera _c, OSPWorld _w)
{
  FrameBuffer *fb = (FrameBuffer *)_fb;
  Renderer *r = (Renderer *)_r;
  Camera *c = (Camera *)_c;
  World *w = (World *)_w;
  Future *f = new Future(fb, r, c, w);
  try {
    fb->beginFrame();
    fb->errorThreshold = r->errorThreshold;
    // LocalFrameBuffer::getRenderTaskIDs (fb/LocalFB.cpp:238-261): once minimumAdaptiveFrames (FrameBuffer.h:183-186) have
    // been accumulated, only the render tasks whose variance estimate is above the renderer's varianceThreshold go on
    fb->sh.taskError = nullptr;
    fb->sh.errorThreshold = r->errorThreshold;
    if (r->errorThreshold > 0.f && fb->hasVariance && fb->taskErrBuf
        && fb->sh.frameID >= (int32_t)std::max(2u, 16u / (uint32_t)std::max(1, r->sh.spp)))
      fb->sh.taskError = fb->taskErrBuf->as<float>();
    const OcWorld *dw = w->prepareFrame(r);
    checkCuda(ospcuda_event_record(f->evStart), "event record");
    const bool finished = fb->sh.targetFrames && fb->doAccum && fb->sh.frameID >= fb->sh.targetFrames; // accumulationFinished()
    OcRenderer rsh = r->sh;
    rsh.worldNumLights = w->sh.numLights;
    rsh.worldHasVolumes = w->sh.hasVolumes;
    rsh.worldNumLights = w->sh.numLights; // (with the geometry lights prepareFrame added)
    rsh.worldShadowOpaque = w->sh.shadowOpaque;
    // which kernel instantiation the frame needs: texture maps concern every renderer; emitters and the microfacet materials only
    // the pathtracer (scivis / ao shade non-OBJ materials with the default surface, like the reference); hdri / cylinder lights
    // the pathtracer and scivis (ao ignores lights)
    const bool maps = texturedMaterials > 0 || rsh.backplate || rsh.maxDepthTexture;
    const bool pt = rsh.type == OC_RENDERER_PATHTRACER, lit = pt || rsh.type == OC_RENDERER_SCIVIS;
    rsh.hasTextures = (maps || (pt && emissiveMaterials > 0) || (lit && w->hasHdriLight)) ? 1u : 0u;
    if (!finished) checkCuda(ospcuda_render_frame(dw, &c->sh, &rsh, &fb->sh), "render kernel");
    checkCuda(ospcuda_event_record(f->evRendered), "event record");
    fb->postProcess();
    checkCuda(ospcuda_event_record(f->evFinished), "event record");
  } catch (...) {
    f->refDec();
    throw;
  }
  return (OSPFuture)f;
}

// ospPick: Renderer::pick (render/Renderer.cpp:94-137) + Renderer_pick (render/Renderer.ispc:43-84): the camera ray
// through screenPos (centre of the lens, no pixel filter), closest hit, retained instance / model handles
OSPPickResult CudaDevice::pick(OSPFrameBuffer, OSPRenderer, OSPCamera _c, OSPWorld _w, const vec2f &screenPos)
{
  Camera *c = (Camera *)_c;
  World *w = (World *)_w;
  OSPPickResult res{};
  res.instance = nullptr; res.model = nullptr; res.primID = 0xFFFFFFFFu; res.hasHit = 0;
  const OcCamera &cam = c->sh;
  const auto lerpf = [](float f, float a, float b) { return (1.f - f) * a + f * b; };
  const float sx = lerpf(screenPos.x, cam.subImage[0], cam.subImage[2]), sy = lerpf(screenPos.y, cam.subImage[1], cam.subImage[3]);
  const vec3f du = make_vec3f(cam.du_size.x, cam.du_size.y, cam.du_size.z), dv = make_vec3f(cam.dv_up.x, cam.dv_up.y, cam.dv_up.z);
  vec3f org = make_vec3f(cam.org.x, cam.org.y, cam.org.z), dir = make_vec3f(cam.dir_00.x, cam.dir_00.y, cam.dir_00.z);
  if (cam.type == OC_CAMERA_PANORAMIC) { // PanoramicCamera_initRay (camera/PanoramicCamera.ispc:14-92); du / dv / dir_00 = the frame
    // (the eye of the per-pixel stereo modes depends on CameraSample::pixel_center, which Renderer_pick leaves unset: mono there)
    const float offset = cam.stereoMode == OSP_STEREO_LEFT ? -cam.ipdOffset.x : cam.stereoMode == OSP_STEREO_RIGHT ? cam.ipdOffset.x : 0.f;
    const float phi = 6.28318530717958647692f * sx, theta = 3.14159265358979323846f * sy;
    const float sinTheta = std::sin(theta), cosTheta = std::cos(theta);
    const vec3f local = make_vec3f(std::cos(phi) * sinTheta, std::sin(phi) * sinTheta, cosTheta);
    const vec3f vz = dir;
    dir = (-local.y) * du + (-local.z) * dv + local.x * vz;
    org = org + offset * cross(dir, dv);
  } else if (cam.type == OC_CAMERA_ORTHOGRAPHIC) org = org + sx * du + sy * dv;
  else dir = normalize(dir + sx * du + sy * dv);
  OcRayHit rh{};
  rh.org[0] = org.x; rh.org[1] = org.y; rh.org[2] = org.z; rh.t0 = cam.nearClip;
  rh.dir[0] = dir.x; rh.dir[1] = dir.y; rh.dir[2] = dir.z; rh.t = INFINITY;
  rh.primID = rh.geomID = rh.instID = 0xFFFFFFFFu;
  const OcWorld *dw = w->prepareFrame();
  DevMem ray;
  ray.upload(&rh, sizeof(rh));
  checkCuda(ospcuda_trace(dw, ray.as<OcRayHit>(), 1, 0), "pick trace");
  checkCuda(ospcuda_memcpy_d2h(&rh, ray.p, sizeof(rh)), "pick read-back");
  if (rh.geomID == 0xFFFFFFFFu || !w->instances) return res;
  Instance *instance = dynamic_cast<Instance *>(w->instances->objectAt(rh.instID));
  Group *group = instance ? instance->group.ptr : nullptr;
  if (!group || !group->geometricModels) return res;
  ManagedObject *model = group->geometricModels->objectAt(rh.geomID);
  instance->refInc();
  model->refInc();
  res.hasHit = 1;
  res.worldPosition[0] = org.x + dir.x * rh.t; res.worldPosition[1] = org.y + dir.y * rh.t; res.worldPosition[2] = org.z + dir.z * rh.t;
  res.instance = (OSPInstance)instance;
  res.model = (OSPGeometricModel)model;
  res.primID = rh.primID;
  return res;
}

int CudaDevice::isReady(OSPFuture f, OSPSyncEvent e) { return ((Future *)f)->isFinished(e); }
void CudaDevice::wait(OSPFuture f, OSPSyncEvent e) { ((Future *)f)->wait(e); }
// ospCancel (api/Device.h; fb/FrameBuffer.cpp:180-187 cancelFrame): the frame's kernels are already queued on the render
// stream, so the request travels as a flag in device memory, set on a side stream: every kernel of the frame reads it when
// it starts and the persistent loops when they fetch work; what has not started yet returns at once. The frame then ends
// early with whatever was accumulated, like the reference's cancelled frame.
void CudaDevice::cancel(OSPFuture _f)
{
  Future *f = (Future *)_f;
  if (f->isFinished(OSP_TASK_FINISHED) || !f->fb || !f->fb->sh.control) return;
  checkCuda(ospcuda_aux_poke_u32(f->fb->sh.control, 1u), "cancel flag");
}
// ospGetProgress (fb/FrameBuffer.ih:27-50 reportProgress): work units finished / work units of the frame, counted on the
// device (render tasks of the megakernels, launch stages of the wavefront renderers) and read on the side stream
float CudaDevice::getProgress(OSPFuture _f)
{
  Future *f = (Future *)_f;
  if (f->isFinished(OSP_TASK_FINISHED)) return 1.f;
  if (!f->fb || !f->fb->sh.control) return 0.f;
  uint32_t done = 0, total = 0;
  checkCuda(ospcuda_aux_peek_u32(f->fb->sh.control + 1, &done), "progress read");
  checkCuda(ospcuda_aux_peek_u32(f->fb->sh.control + 2, &total), "progress read");
  if (total == 0) return 0.f;
  return std::min(0.99f, (float)done / (float)total);
}
float CudaDevice::getTaskDuration(OSPFuture f) { return ((Future *)f)->duration(); }

uint32_t CudaDevice::allocMaterialSlot()
{
  uint32_t s;
  if (!freeSlots.empty()) { s = freeSlots.back(); freeSlots.pop_back(); }
  else { s = (uint32_t)materials.size(); materials.push_back(OcMaterial{}); }
  materials[s] = OcMaterial{};
  materials[s].type = OC_MATERIAL_OTHER;
  materialsDirty = true;
  return s;
}
void CudaDevice::freeMaterialSlot(uint32_t slot) { freeSlots.push_back(slot); }
void CudaDevice::setMaterial(uint32_t slot, const OcMaterial &m)
{
  materials[slot] = m;
  materialsDirty = true;
  materialVersion++;
}
const OcMaterial *CudaDevice::materialTable()
{
  if (materialsDirty && !materials.empty()) {
    materialsDev.upload(materials.data(), materials.size() * sizeof(OcMaterial));
    materialsDirty = false;
  }
  return materialsDev.as<OcMaterial>();
}

} // namespace cuda
} // namespace ospray

using namespace ospray;

// modules/cpu/ModuleInit.cpp:24-54
extern "C" __attribute__((visibility("default"))) OSPError ospray_module_init_cuda(int16_t versionMajor, int16_t versionMinor,
    int16_t /*versionPatch*/)
{
  const OSPError status = moduleVersionCheck(versionMajor, versionMinor);
  if (status == OSP_NO_ERROR) {
    api::Device::registerType<cuda::CudaDevice>("cuda");
    cuda::registerMultiDevice(); // "multidevice": N cuda sub-devices in one process (modules/multiDevice analogue)
  }
  return status;
}

extern "C" {
// the handles may come from the cuda device or from the multidevice (sub-object 0 holds the composed image; counters add up)
void *ospcudaFrameBufferDevicePointer(OSPFrameBuffer fb, OSPFrameBufferChannel ch, uint64_t *bpp)
{
  cuda::FrameBuffer *f = (cuda::FrameBuffer *)cuda::subObject(fb, 0);
  if (bpp) *bpp = f->bytesPerPixel(ch);
  return f->devicePointer(ch);
}
uint64_t ospcudaGetFusedRayCount(OSPFrameBuffer fb)
{
  uint64_t n = 0;
  for (size_t i = 0; cuda::subObject(fb, i); i++) n += ((cuda::FrameBuffer *)cuda::subObject(fb, i))->rayCounts[3];
  return n;
}
void ospcudaGetRayCounts(OSPFrameBuffer fb, uint64_t counts[3])
{
  counts[0] = counts[1] = counts[2] = 0;
  for (size_t i = 0; cuda::subObject(fb, i); i++)
    for (int k = 0; k < 3; k++) counts[k] += ((cuda::FrameBuffer *)cuda::subObject(fb, i))->rayCounts[k];
}
float ospcudaGetGroupBuildTime(OSPGroup g) { return ((cuda::Group *)cuda::subObject(g, 0))->lastBuildMs; }
uint64_t ospcudaKernelLaunchCount(void) { return ospcuda_kernel_launch_count(); }
void *ospcudaRenderStream(void) { return ospcuda_stream(); }
void ospcudaGetFrameStageTimes(OSPFuture f, float ms[2])
{
  api::currentDevice().wait(f, OSP_TASK_FINISHED);
  cuda::Future *fu = (cuda::Future *)cuda::subObject(f, 0);
  ms[0] = ms[1] = 0.f;
  ospcuda_event_elapsed_ms(fu->evStart, fu->evRendered, &ms[0]);
  ospcuda_event_elapsed_ms(fu->evRendered, fu->evFinished, &ms[1]);
}
void ospcudaSetRenderStream(void *stream) { ospray::cuda::checkCuda(ospcuda_set_stream(stream), "set stream"); }
const void *ospcudaTextureDevicePointer(OSPTexture t) { return ((cuda::Texture2D *)cuda::subObject(t, 0))->devPtr(); }
const void *ospcudaWorldDevicePointer(OSPWorld w) { return ((cuda::World *)cuda::subObject(w, 0))->prepareFrame(); }
}
