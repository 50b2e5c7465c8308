// MultiCudaDevice : api::Device -- device type "multidevice": N cuda sub-devices in ONE process, one per GPU of the box.
//
// reference: modules/multiDevice/MultiDevice.cpp:32-57 (numSubdevices <- OSPRAY_NUM_SUBDEVICES, one sub-device each),
// :60-393 (every object of the application is a MultiDeviceObject holding one object per sub-device; every API call is
// forwarded to each of them, object-valued parameters and data arrays of handles are translated per sub-device),
// :395-418 + MultiDeviceLoadBalancer.cpp:19-60 (the frame: every sub-device renders the image tiles with
// tileID % N == its index into the one framebuffer), :472-508 (renderFrame / futures).
//
// Here: the scene is replicated on every GPU (each sub-device commits its own copy and builds its own hierarchy), sub-device
// i renders the 64x64 tiles with tile % N == i on GPU i, and GPU 0 then PULLS the finished tiles of every other GPU into
// its own image over NVLink peer memory (one small kernel per peer on GPU 0's render stream, ordered behind the peer's
// frame by its event). ospMapFrameBuffer reads GPU 0's image. No host staging, no NCCL, nothing outside ospRenderFrame:
// an application written for the OSPRay API runs unchanged (--osp:device=multidevice or OSPRAY_DEVICE=multidevice).
// With fewer GPUs than sub-devices several sub-devices share a GPU (ordinal = index % device count): same result.
#include <algorithm>
#include <cstdlib>
#include <unordered_map>
#include "CudaDevice.h"

namespace ospray {
namespace cuda {

struct MultiObject : public ManagedObject
{
  ~MultiObject() override
  {
    for (ManagedObject *s : subs)
      if (s) s->refDec();
    if (deleter) deleter(deleterUser, deleterData);
  }
  std::string toString() const override { return "ospray::MultiDeviceObject"; }
  std::vector<ManagedObject *> subs;                  // one per sub-device, one reference each
  std::vector<std::vector<OSPObject>> handles;        // shared data of handles: the per-sub-device translations
  std::vector<Ref<MultiObject>> children;             // data of handles: the application's objects it holds (kept alive like
                                                      // the sub-device data keeps the sub-objects alive: ospPick hands them back)
  OSPDeleterCallback deleter = nullptr;               // ospNewSharedData's callback runs once, not once per sub-device
  const void *deleterUser = nullptr, *deleterData = nullptr;
};

struct MultiFuture : public MultiObject
{
  ~MultiFuture() override
  {
    if (evComposed) {
      ospcuda_event_synchronize(evComposed);
      ospcuda_event_destroy(evComposed);
    }
  }
  void *evComposed = nullptr; // on GPU 0's stream, behind the tile pulls
  Ref<MultiObject> fb;
};

struct MultiCudaDevice : public api::Device
{
  std::vector<Ref<CudaDevice>> subdevices;
  std::unordered_map<ManagedObject *, MultiObject *> fromSub0; // instances and geometric models (ospPick hands them back)

  ~MultiCudaDevice() override = default;

  size_t N() const { return subdevices.size(); }
  void use(size_t i) { checkCuda(ospcuda_make_current(subdevices[i]->cudaOrdinal), "device switch"); }
  static MultiObject *M(const void *h) { return (MultiObject *)h; }

  void commit() override
  {
    api::Device::commit();
    for (auto &d : subdevices) d->disableMipMapGeneration = getParam<bool>("disableMipMapGeneration", false);
    if (!subdevices.empty()) return; // the sub-devices are created once
    int n = getParam<int>("numSubdevices", 0);
    if (const char *e = getenv("OSPRAY_NUM_SUBDEVICES")) n = atoi(e);
    const int gpus = ospcuda_device_count();
    if (gpus <= 0) throw std::runtime_error("multidevice: no CUDA device (the cuda device has no CPU fallback)");
    if (n <= 0) n = gpus; // default: every GPU of the box
    for (int i = 0; i < n; i++) {
      CudaDevice *d = new CudaDevice;
      const int ord = i % gpus, nn = n, rk = i;
      d->setParam("cudaDevice", OSP_INT, &ord);
      d->setParam("multiNumSubdevices", OSP_INT, &nn);
      d->setParam("multiSubdeviceRank", OSP_INT, &rk);
      d->msg_fcn = [this](void *, const char *m) { msg_fcn(statusUserData, m); };
      d->error_fcn = [this](void *, OSPError e, const char *m) { error_fcn(errorUserData, e, m); };
      d->warningsAreErrors = warningsAreErrors;
      d->commit();
      d->disableMipMapGeneration = getParam<bool>("disableMipMapGeneration", false);
      subdevices.emplace_back(d);
      d->refDec();
    }
    // GPU 0 reads the other GPUs' images
    use(0);
    for (size_t i = 1; i < N(); i++)
      if (subdevices[i]->cudaOrdinal != subdevices[0]->cudaOrdinal)
        checkCuda(ospcuda_enable_peer(subdevices[i]->cudaOrdinal), "peer access GPU 0 -> sub-device GPU");
    postStatusMsg(OSP_LOG_INFO, "multidevice: " + std::to_string(n) + " cuda sub-devices on " + std::to_string(std::min(n, gpus)) + " GPU(s)");
  }

  int loadModule(const char *name) override { return loadLocalModule(name); }

  template <typename F>
  MultiObject *create(OSPDataType kind, F &&make)
  {
    MultiObject *m = new MultiObject;
    m->managedObjectType = kind;
    try {
      for (size_t i = 0; i < N(); i++) {
        use(i);
        ManagedObject *s = (ManagedObject *)make(*subdevices[i], i);
        if (!s) { // unknown type on the sub-device: same answer from the multi-device
          m->refDec();
          return nullptr;
        }
        m->subs.push_back(s);
      }
    } catch (...) {
      m->refDec();
      throw;
    }
    return m;
  }

  OSPData newSharedData(const void *sharedData, OSPDataType t, const vec3ul &n, const vec3l &s, OSPDeleterCallback cb, const void *up) override
  {
    MultiObject *m = new MultiObject;
    m->managedObjectType = OSP_DATA;
    m->deleter = cb; m->deleterUser = up; m->deleterData = sharedData;
    try {
      if (isObjectType(t)) {
        // an array of handles: every sub-device gets an array of ITS objects (MultiDevice.cpp:84-118)
        const size_t count = (size_t)n.x * n.y * n.z;
        if ((s.x != 0 && s.x != (int64_t)sizeof(OSPObject)) || n.y != 1 || n.z != 1)
          throw std::runtime_error("multidevice: shared data of object handles must be a compact 1D array");
        m->handles.resize(N());
        for (size_t i = 0; i < N(); i++) {
          m->handles[i].resize(count);
          for (size_t k = 0; k < count; k++) {
            MultiObject *o = ((MultiObject *const *)sharedData)[k];
            m->handles[i][k] = o ? (OSPObject)o->subs[i] : nullptr;
            if (i == 0 && o) m->children.emplace_back(o);
          }
          use(i);
          m->subs.push_back((ManagedObject *)subdevices[i]->newSharedData(m->handles[i].data(), t, n, vec3l{0, 0, 0}, nullptr, nullptr));
        }
      } else
        for (size_t i = 0; i < N(); i++) {
          use(i);
          m->subs.push_back((ManagedObject *)subdevices[i]->newSharedData(sharedData, t, n, s, nullptr, nullptr));
        }
    } catch (...) {
      m->deleter = nullptr; // the API layer reports the error; the application still owns its memory
      m->refDec();
      throw;
    }
    return (OSPData)m;
  }
  OSPData newData(OSPDataType t, const vec3ul &n) override
  {
    return (OSPData)create(OSP_DATA, [&](CudaDevice &d, size_t) { return d.newData(t, n); });
  }
  void copyData(const OSPData source, OSPData destination, const vec3ul &di) override
  {
    for (size_t i = 0; i < N(); i++) {
      use(i);
      subdevices[i]->copyData((OSPData)M(source)->subs[i], (OSPData)M(destination)->subs[i], di);
    }
    for (const Ref<MultiObject> &c : M(source)->children) M(destination)->children.push_back(c);
  }
  OSPLight newLight(const char *type) override { return (OSPLight)create(OSP_LIGHT, [&](CudaDevice &d, size_t) { return d.newLight(type); }); }
  OSPCamera newCamera(const char *type) override { return (OSPCamera)create(OSP_CAMERA, [&](CudaDevice &d, size_t) { return d.newCamera(type); }); }
  OSPGeometry newGeometry(const char *type) override { return (OSPGeometry)create(OSP_GEOMETRY, [&](CudaDevice &d, size_t) { return d.newGeometry(type); }); }
  OSPVolume newVolume(const char *type) override { return (OSPVolume)create(OSP_VOLUME, [&](CudaDevice &d, size_t) { return d.newVolume(type); }); }
  OSPGeometricModel newGeometricModel(OSPGeometry g) override
  {
    MultiObject *m = create(OSP_GEOMETRIC_MODEL, [&](CudaDevice &d, size_t i) { return d.newGeometricModel(g ? (OSPGeometry)M(g)->subs[i] : nullptr); });
    if (m) fromSub0[m->subs[0]] = m;
    return (OSPGeometricModel)m;
  }
  OSPVolumetricModel newVolumetricModel(OSPVolume v) override
  {
    return (OSPVolumetricModel)create(OSP_VOLUMETRIC_MODEL, [&](CudaDevice &d, size_t i) { return d.newVolumetricModel(v ? (OSPVolume)M(v)->subs[i] : nullptr); });
  }
  OSPMaterial newMaterial(const char *type) override { return (OSPMaterial)create(OSP_MATERIAL, [&](CudaDevice &d, size_t) { return d.newMaterial(type); }); }
  OSPTransferFunction newTransferFunction(const char *type) override
  {
    return (OSPTransferFunction)create(OSP_TRANSFER_FUNCTION, [&](CudaDevice &d, size_t) { return d.newTransferFunction(type); });
  }
  OSPTexture newTexture(const char *type) override { return (OSPTexture)create(OSP_TEXTURE, [&](CudaDevice &d, size_t) { return d.newTexture(type); }); }
  OSPGroup newGroup() override { return (OSPGroup)create(OSP_GROUP, [&](CudaDevice &d, size_t) { return d.newGroup(); }); }
  OSPInstance newInstance(OSPGroup g) override
  {
    MultiObject *m = create(OSP_INSTANCE, [&](CudaDevice &d, size_t i) { return d.newInstance(g ? (OSPGroup)M(g)->subs[i] : nullptr); });
    if (m) fromSub0[m->subs[0]] = m;
    return (OSPInstance)m;
  }
  OSPWorld newWorld() override { return (OSPWorld)create(OSP_WORLD, [&](CudaDevice &d, size_t) { return d.newWorld(); }); }
  OSPRenderer newRenderer(const char *type) override { return (OSPRenderer)create(OSP_RENDERER, [&](CudaDevice &d, size_t) { return d.newRenderer(type); }); }
  OSPImageOperation newImageOp(const char *type) override
  {
    return (OSPImageOperation)create(OSP_IMAGE_OPERATION, [&](CudaDevice &d, size_t) { return d.newImageOp(type); });
  }
  OSPFrameBuffer frameBufferCreate(const vec2i &size, const OSPFrameBufferFormat mode, const uint32_t channels) override
  {
    return (OSPFrameBuffer)create(OSP_FRAMEBUFFER, [&](CudaDevice &d, size_t) { return d.frameBufferCreate(size, mode, channels); });
  }

  box3f getBounds(OSPObject o) override
  {
    use(0);
    return subdevices[0]->getBounds((OSPObject)M(o)->subs[0]);
  }
  void setObjectParam(OSPObject object, const char *name, OSPDataType type, const void *mem) override
  {
    if (isObjectType(type)) M(object)->setParam(name, type, mem); // the wrapper retains the wrapper, like the objects below them do
    for (size_t i = 0; i < N(); i++) {
      use(i);
      if (isObjectType(type)) { // the value is a handle: hand each sub-device its own object (MultiDevice.cpp:230-262)
        MultiObject *v = *(MultiObject *const *)mem;
        OSPObject h = v ? (OSPObject)v->subs[i] : nullptr;
        subdevices[i]->setObjectParam((OSPObject)M(object)->subs[i], name, type, &h);
      } else
        subdevices[i]->setObjectParam((OSPObject)M(object)->subs[i], name, type, mem);
    }
  }
  void removeObjectParam(OSPObject object, const char *name) override
  {
    M(object)->removeParam(name);
    for (size_t i = 0; i < N(); i++) subdevices[i]->removeObjectParam((OSPObject)M(object)->subs[i], name);
  }
  void commit(OSPObject object) override
  {
    for (size_t i = 0; i < N(); i++) {
      use(i);
      subdevices[i]->commit((OSPObject)M(object)->subs[i]);
    }
  }
  void release(OSPObject o) override
  {
    MultiObject *m = M(o);
    if (m->useCount() == 1 && !m->subs.empty()) fromSub0.erase(m->subs[0]);
    m->refDec();
  }
  void retain(OSPObject o) override { M(o)->refInc(); }

  // ---- framebuffer --------------------------------------------------------------------------------------------------
  // pull the tiles of every other sub-device's channel into sub-device 0's buffer (GPU 0's render stream; `after`: the
  // peers' frames, nullptr = they are known to be finished)
  void composeChannel(MultiObject *mfb, OSPFrameBufferChannel ch, MultiFuture *after)
  {
    FrameBuffer *f0 = (FrameBuffer *)mfb->subs[0];
    void *dst = f0->devicePointer(ch);
    const size_t bpp = f0->bytesPerPixel(ch);
    if (!dst || !bpp) return;
    use(0);
    for (size_t i = 1; i < N(); i++) {
      FrameBuffer *fi = (FrameBuffer *)mfb->subs[i];
      if (after) checkCuda(ospcuda_stream_wait_event(((Future *)after->subs[i])->evFinished), "frame order");
      checkCuda(ospcuda_fb_gather_tiles(dst, fi->devicePointer(ch), f0->size.x, f0->size.y, (uint32_t)bpp, (uint32_t)N(), (uint32_t)i), "tile pull");
    }
  }
  const void *frameBufferMap(OSPFrameBuffer fb, const OSPFrameBufferChannel ch) override
  {
    MultiObject *m = M(fb);
    if (ch != OSP_FB_COLOR) { // the colour image is composed inside ospRenderFrame, the other channels when they are asked for
      for (size_t i = 1; i < N(); i++) {
        use(i);
        checkCuda(ospcuda_synchronize(), "sub-device frame");
      }
      composeChannel(m, ch, nullptr);
    }
    use(0);
    return ((FrameBuffer *)m->subs[0])->map(ch);
  }
  void frameBufferUnmap(const void *mapped, OSPFrameBuffer fb) override
  {
    use(0);
    ((FrameBuffer *)M(fb)->subs[0])->unmap(mapped);
  }
  // the estimate over the whole image: every render task's error comes from the sub-device that owns its tile
  // (LocalFB.cpp:188-236 on the composed buffers in the reference)
  float getVariance(OSPFrameBuffer fb) override
  {
    MultiObject *m = M(fb);
    FrameBuffer *f0 = (FrameBuffer *)m->subs[0];
    std::vector<float> v(N());
    for (size_t i = 0; i < N(); i++) {
      use(i);
      v[i] = ((FrameBuffer *)m->subs[i])->getVariance();
    }
    if (!f0->hasVariance || f0->taskErrors.empty() || !f0->taskErrorsValid) return v[0];
    const int tasksX = (f0->size.x + 7) / 8, tilesX = (f0->size.x + 63) / 64;
    float maxErr = 0.f, sumAct = 0.f;
    int active = 0;
    for (size_t t = 0; t < f0->taskErrors.size(); t++) {
      const int tx = (int)(t % tasksX), ty = (int)(t / tasksX);
      const size_t owner = (size_t)((ty / 8) * tilesX + tx / 8) % N();
      const float e = ((FrameBuffer *)m->subs[owner])->taskErrors[t];
      maxErr = std::max(maxErr, e);
      if (e >= f0->errorThreshold) { sumAct += e; active++; }
    }
    return active ? sumAct / active : maxErr;
  }
  void resetAccumulation(OSPFrameBuffer fb) override
  {
    for (size_t i = 0; i < N(); i++) {
      use(i);
      subdevices[i]->resetAccumulation((OSPFrameBuffer)M(fb)->subs[i]);
    }
  }

  // ---- frame ----------------------------------------------------------------------------------------------------------
  OSPFuture renderFrame(OSPFrameBuffer fb, OSPRenderer r, OSPCamera c, OSPWorld w) override
  {
    MultiFuture *f = new MultiFuture;
    f->managedObjectType = OSP_FUTURE;
    f->fb = M(fb);
    try {
      for (size_t i = 0; i < N(); i++) { // every GPU starts on its tiles (the calls only queue work)
        use(i);
        f->subs.push_back((ManagedObject *)subdevices[i]->renderFrame((OSPFrameBuffer)M(fb)->subs[i], (OSPRenderer)M(r)->subs[i],
            (OSPCamera)M(c)->subs[i], (OSPWorld)M(w)->subs[i]));
      }
      composeChannel(M(fb), OSP_FB_COLOR, f);
      use(0);
      checkCuda(ospcuda_event_create(&f->evComposed), "event");
      checkCuda(ospcuda_event_record(f->evComposed), "event record");
    } catch (...) {
      f->refDec();
      throw;
    }
    return (OSPFuture)f;
  }
  int isReady(OSPFuture _f, OSPSyncEvent e) override
  {
    MultiFuture *f = (MultiFuture *)_f;
    for (size_t i = 0; i < N(); i++)
      if (!subdevices[i]->isReady((OSPFuture)f->subs[i], e)) return 0;
    if (e == OSP_TASK_FINISHED || e == OSP_FRAME_FINISHED) return ospcuda_event_query(f->evComposed) == 1;
    return 1;
  }
  void wait(OSPFuture _f, OSPSyncEvent e) override
  {
    MultiFuture *f = (MultiFuture *)_f;
    for (size_t i = 0; i < N(); i++) {
      use(i);
      subdevices[i]->wait((OSPFuture)f->subs[i], e);
    }
    if (e == OSP_TASK_FINISHED || e == OSP_FRAME_FINISHED) checkCuda(ospcuda_event_synchronize(f->evComposed), "event wait");
  }
  void cancel(OSPFuture _f) override
  {
    for (size_t i = 0; i < N(); i++) {
      use(i);
      subdevices[i]->cancel((OSPFuture)M(_f)->subs[i]);
    }
  }
  float getProgress(OSPFuture _f) override
  {
    float p = 1.f;
    for (size_t i = 0; i < N(); i++) {
      use(i);
      p = std::min(p, subdevices[i]->getProgress((OSPFuture)M(_f)->subs[i]));
    }
    return p;
  }
  float getTaskDuration(OSPFuture _f) override
  {
    float t = 0.f;
    for (size_t i = 0; i < N(); i++) t = std::max(t, subdevices[i]->getTaskDuration((OSPFuture)M(_f)->subs[i]));
    return t;
  }
  OSPPickResult pick(OSPFrameBuffer fb, OSPRenderer r, OSPCamera c, OSPWorld w, const vec2f &p) override
  {
    use(0);
    OSPPickResult res = subdevices[0]->pick((OSPFrameBuffer)M(fb)->subs[0], (OSPRenderer)M(r)->subs[0], (OSPCamera)M(c)->subs[0],
        (OSPWorld)M(w)->subs[0], p);
    if (res.hasHit) { // the application gets (and later releases) ITS handles
      ManagedObject *si = (ManagedObject *)res.instance, *sm = (ManagedObject *)res.model;
      auto ii = fromSub0.find(si), mi = fromSub0.find(sm);
      if (ii == fromSub0.end() || mi == fromSub0.end()) { res.hasHit = 0; res.instance = nullptr; res.model = nullptr; }
      else {
        ii->second->refInc(); mi->second->refInc();
        res.instance = (OSPInstance)ii->second; res.model = (OSPGeometricModel)mi->second;
      }
      si->refDec(); sm->refDec();
    }
    return res;
  }
};

void registerMultiDevice() { api::Device::registerType<MultiCudaDevice>("multidevice"); }

// the cuda-device extension entry points accept the handles of either device: sub-object `i` of a multi-device handle
ManagedObject *subObject(void *handle, size_t i)
{
  ManagedObject *o = (ManagedObject *)handle;
  if (MultiObject *m = dynamic_cast<MultiObject *>(o)) return i < m->subs.size() ? m->subs[i] : nullptr;
  return i == 0 ? o : nullptr;
}

} // namespace cuda
} // namespace ospray
