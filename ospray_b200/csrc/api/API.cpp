// The OSPRay C API: every osp*() call is routed to api::Device::current.
// reference: ospray/api/API.cpp (ospInit :86-154, device calls :156-329, data :333-400,
// objects :404-560, framebuffer :564-620, render :624-733) and
// ospray/api/ospray_util_impl.cpp:7-192. Error convention (API.cpp:45-66): device methods throw,
// the shim converts to handleError() + a neutral return value.
#include <cstdlib>
#include <new>
#include <sstream>
#include "Device.h"
#include "ospray/ospray_util.h"

using namespace ospray;
using ospray::api::Device;

#define OSP_TRY try {
#define OSP_CATCH(ret)                                                                            \
  }                                                                                               \
  catch (const std::bad_alloc &) { handleError(OSP_OUT_OF_MEMORY, "OSPRay was unable to allocate memory"); return ret; } \
  catch (const std::exception &e) { handleError(OSP_UNKNOWN_ERROR, e.what()); return ret; }      \
  catch (...) { handleError(OSP_UNKNOWN_ERROR, "an unrecognized exception was caught"); return ret; }

static Device &dev()
{
  if (!api::deviceIsSet())
    throw std::runtime_error("OSPRay not yet initialized (most likely this means you tried to call an ospray API function "
                             "before first calling ospInit())");
  return api::currentDevice();
}

static std::string argValue(const std::string &a)
{
  const size_t eq = a.find('=');
  return eq == std::string::npos ? std::string() : a.substr(eq + 1);
}
static void removeArg(int &ac, const char **av, int where)
{
  for (int i = where + 1; i < ac; i++) av[i - 1] = av[i];
  --ac;
}
static std::vector<std::string> splitCommas(const std::string &s)
{
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, ','))
    if (!item.empty()) out.push_back(item);
  return out;
}

extern "C" {

OSPError ospInit(int *_ac, const char **_av)
{
  OSP_TRY
  if (api::deviceIsSet()) throw std::runtime_error("device already exists [ospInit() called twice?]");
  if (const char *mods = getenv("OSPRAY_LOAD_MODULES"))
    for (auto &m : splitCommas(mods)) loadLocalModule(m);
  std::vector<std::pair<std::string, std::string>> deviceParams;
  bool debug = false;
  int logLevel = -1, numThreads = -1;
  std::string logOutput, errorOutput;
  if (_ac && _av) {
    for (int i = 1; i < *_ac;) {
      const std::string a(_av[i]);
      if (a.rfind("--osp:load-modules", 0) == 0) {
        const std::string mods = argValue(a);
        if (mods.empty()) throw std::runtime_error("Invalid module name(s) provided to --osp:load-modules.");
        for (auto &m : splitCommas(mods)) loadLocalModule(m);
        removeArg(*_ac, _av, i);
      } else if (a.rfind("--osp:device=", 0) == 0) {
        Device::current = Device::createDevice(argValue(a).c_str());
        Device::current->refDec();
        removeArg(*_ac, _av, i);
      } else if (a == "--osp:debug") { debug = true; removeArg(*_ac, _av, i); }
      else if (a == "--osp:verbose") { logLevel = OSP_LOG_INFO; logOutput = "cout"; errorOutput = "cerr"; removeArg(*_ac, _av, i); }
      else if (a == "--osp:vv") { logLevel = OSP_LOG_DEBUG; logOutput = "cout"; errorOutput = "cerr"; removeArg(*_ac, _av, i); }
      else if (a.rfind("--osp:log-level=", 0) == 0) {
        const std::string v = argValue(a);
        logLevel = v == "debug" ? OSP_LOG_DEBUG : v == "info" ? OSP_LOG_INFO : v == "warning" ? OSP_LOG_WARNING
            : v == "error" ? OSP_LOG_ERROR : OSP_LOG_NONE;
        removeArg(*_ac, _av, i);
      } else if (a.rfind("--osp:log-output=", 0) == 0) { logOutput = argValue(a); removeArg(*_ac, _av, i); }
      else if (a.rfind("--osp:error-output=", 0) == 0) { errorOutput = argValue(a); removeArg(*_ac, _av, i); }
      else if (a.rfind("--osp:num-threads=", 0) == 0) { numThreads = atoi(argValue(a).c_str()); removeArg(*_ac, _av, i); }
      else if (a.rfind("--osp:device-params=", 0) == 0) {
        for (auto &kv : splitCommas(argValue(a))) {
          const size_t c = kv.find(':');
          if (c != std::string::npos) deviceParams.emplace_back(kv.substr(0, c), kv.substr(c + 1));
        }
        removeArg(*_ac, _av, i);
      } else
        ++i;
    }
  }
  if (!api::deviceIsSet()) {
    const char *dn = getenv("OSPRAY_DEVICE");
    const std::string name = dn ? dn : "cpu";
    if (!dn) loadLocalModule("cpu");
    Device::current = Device::createDevice(name.c_str());
    Device::current->refDec();
  }
  Device &d = api::currentDevice();
  if (debug) { bool b = true; d.setParam("debug", OSP_BOOL, &b); }
  if (logLevel >= 0) d.setParam("logLevel", OSP_INT, &logLevel);
  if (numThreads >= 0) d.setParam("numThreads", OSP_INT, &numThreads);
  if (!logOutput.empty()) d.setParam("logOutput", OSP_STRING, logOutput.c_str());
  if (!errorOutput.empty()) d.setParam("errorOutput", OSP_STRING, errorOutput.c_str());
  for (auto &kv : deviceParams) {
    // numbers become ints, everything else strings (OSPCommon.cpp:178-206)
    char *end = nullptr;
    const long v = strtol(kv.second.c_str(), &end, 10);
    if (end && *end == 0 && !kv.second.empty()) { int iv = (int)v; d.setParam(kv.first, OSP_INT, &iv); }
    else d.setParam(kv.first, OSP_STRING, kv.second.c_str());
  }
  d.commit();
  return OSP_NO_ERROR;
  OSP_CATCH(OSP_INVALID_OPERATION)
}

int64_t ospDeviceGetProperty(OSPDevice _d, OSPDeviceProperty p)
{
  OSP_TRY
  return ((Device *)_d)->getProperty(p);
  OSP_CATCH(0)
}

void ospShutdown(void)
{
  OSP_TRY
  Device::current = nullptr;
  OSP_CATCH()
}

OSPDevice ospNewDevice(const char *deviceType)
{
  OSP_TRY
  return (OSPDevice)Device::createDevice(deviceType);
  OSP_CATCH(nullptr)
}

void ospSetCurrentDevice(OSPDevice _d)
{
  OSP_TRY
  Device *d = (Device *)_d;
  if (!d) { Device::current = nullptr; return; }
  if (!d->isCommitted()) throw std::runtime_error("You must commit the device before using it!");
  Device::current = d;
  OSP_CATCH()
}

OSPDevice ospGetCurrentDevice(void)
{
  OSP_TRY
  Device *d = Device::current.get();
  if (d) d->refInc();
  return (OSPDevice)d;
  OSP_CATCH(nullptr)
}

void ospDeviceSetParam(OSPDevice _d, const char *id, OSPDataType type, const void *mem)
{
  OSP_TRY
  ((Device *)_d)->setParam(id, type, mem);
  OSP_CATCH()
}
void ospDeviceRemoveParam(OSPDevice _d, const char *id)
{
  OSP_TRY
  ((Device *)_d)->removeParam(id);
  OSP_CATCH()
}
void ospDeviceSetStatusCallback(OSPDevice _d, OSPStatusCallback cb, void *userData)
{
  OSP_TRY
  Device *d = (Device *)_d;
  if (cb) d->msg_fcn = cb; else d->msg_fcn = [](void *, const char *) {};
  d->statusUserData = userData;
  OSP_CATCH()
}
void ospDeviceSetErrorCallback(OSPDevice _d, OSPErrorCallback cb, void *userData)
{
  OSP_TRY
  Device *d = (Device *)_d;
  if (cb) d->error_fcn = cb; else d->error_fcn = [](void *, OSPError, const char *) {};
  d->errorUserData = userData;
  OSP_CATCH()
}
OSPError ospDeviceGetLastErrorCode(OSPDevice _d)
{
  OSP_TRY
  return ((Device *)_d)->lastErrorCode;
  OSP_CATCH(OSP_NO_ERROR)
}
const char *ospDeviceGetLastErrorMsg(OSPDevice _d)
{
  OSP_TRY
  return ((Device *)_d)->lastErrorMsg.c_str();
  OSP_CATCH(nullptr)
}
void ospDeviceCommit(OSPDevice _d)
{
  Device *d = (Device *)_d;
  try {
    d->commit();
  } catch (const std::exception &e) {
    // the device being committed is usually not current yet: record the failure on it directly
    d->lastErrorCode = OSP_UNKNOWN_ERROR;
    d->lastErrorMsg = std::string("#ospray: ") + e.what();
    d->error_fcn(d->errorUserData, OSP_UNKNOWN_ERROR, e.what());
    if (!api::deviceIsSet() || Device::current.get() != d) handleError(OSP_UNKNOWN_ERROR, e.what());
  }
}
void ospDeviceRelease(OSPDevice _d)
{
  OSP_TRY
  if (_d) ((Device *)_d)->refDec();
  OSP_CATCH()
}
void ospDeviceRetain(OSPDevice _d)
{
  OSP_TRY
  if (_d) ((Device *)_d)->refInc();
  OSP_CATCH()
}

OSPError ospLoadModule(const char *name)
{
  OSP_TRY
  if (api::deviceIsSet()) return (OSPError)dev().loadModule(name);
  return loadLocalModule(name);
  OSP_CATCH(OSP_UNKNOWN_ERROR)
}

OSPData ospNewSharedData(const void *sharedData, OSPDataType type, uint64_t n1, int64_t s1, uint64_t n2, int64_t s2, uint64_t n3,
    int64_t s3, OSPDeleterCallback freeFn, const void *userPtr)
{
  OSP_TRY
  vec3ul n{n1, n2, n3};
  vec3l s{s1, s2, s3};
  return dev().newSharedData(sharedData, type, n, s, freeFn, userPtr);
  OSP_CATCH(nullptr)
}
OSPData ospNewData(OSPDataType type, uint64_t n1, uint64_t n2, uint64_t n3)
{
  OSP_TRY
  vec3ul n{n1, n2, n3};
  return dev().newData(type, n);
  OSP_CATCH(nullptr)
}
void ospCopyData(const OSPData source, OSPData destination, uint64_t i1, uint64_t i2, uint64_t i3)
{
  OSP_TRY
  vec3ul d{i1, i2, i3};
  dev().copyData(source, destination, d);
  OSP_CATCH()
}

#define OSP_NEW_TYPED(FN, RET, CALL)                                                             \
  RET FN(const char *type)                                                                        \
  {                                                                                               \
    OSP_TRY                                                                                       \
    if (!type) throw std::runtime_error(#FN ": NULL type string");                                \
    RET r = dev().CALL(type);                                                                     \
    if (!r) postStatusMsg(OSP_LOG_WARNING, std::string("#ospray: could not create object of type '") + type + "'"); \
    return r;                                                                                     \
    OSP_CATCH(nullptr)                                                                            \
  }
OSP_NEW_TYPED(ospNewLight, OSPLight, newLight)
OSP_NEW_TYPED(ospNewCamera, OSPCamera, newCamera)
OSP_NEW_TYPED(ospNewGeometry, OSPGeometry, newGeometry)
OSP_NEW_TYPED(ospNewVolume, OSPVolume, newVolume)
OSP_NEW_TYPED(ospNewMaterial, OSPMaterial, newMaterial)
OSP_NEW_TYPED(ospNewTransferFunction, OSPTransferFunction, newTransferFunction)
OSP_NEW_TYPED(ospNewTexture, OSPTexture, newTexture)

static std::string normalizedTypeName(const char *type)
{
  // API.cpp:590-598: '-' and ':' become '_'
  std::string t(type);
  for (char &c : t)
    if (c == '-' || c == ':') c = '_';
  return t;
}
OSPImageOperation ospNewImageOperation(const char *type)
{
  OSP_TRY
  if (!type) throw std::runtime_error("ospNewImageOperation: NULL type string");
  return dev().newImageOp(normalizedTypeName(type).c_str());
  OSP_CATCH(nullptr)
}
OSPRenderer ospNewRenderer(const char *type)
{
  OSP_TRY
  if (!type) throw std::runtime_error("ospNewRenderer: NULL type string");
  OSPRenderer r = dev().newRenderer(normalizedTypeName(type).c_str());
  if (!r) postStatusMsg(OSP_LOG_WARNING, std::string("#ospray: could not create renderer '") + type + "'");
  return r;
  OSP_CATCH(nullptr)
}

OSPGeometricModel ospNewGeometricModel(OSPGeometry g)
{
  OSP_TRY
  return dev().newGeometricModel(g);
  OSP_CATCH(nullptr)
}
OSPVolumetricModel ospNewVolumetricModel(OSPVolume v)
{
  OSP_TRY
  return dev().newVolumetricModel(v);
  OSP_CATCH(nullptr)
}
OSPGroup ospNewGroup(void)
{
  OSP_TRY
  return dev().newGroup();
  OSP_CATCH(nullptr)
}
OSPInstance ospNewInstance(OSPGroup g)
{
  OSP_TRY
  return dev().newInstance(g);
  OSP_CATCH(nullptr)
}
OSPWorld ospNewWorld(void)
{
  OSP_TRY
  return dev().newWorld();
  OSP_CATCH(nullptr)
}
OSPBounds ospGetBounds(OSPObject o)
{
  OSPBounds b = {{NAN, NAN, NAN}, {NAN, NAN, NAN}};
  OSP_TRY
  const box3f r = dev().getBounds(o);
  b.lower[0] = r.lower.x; b.lower[1] = r.lower.y; b.lower[2] = r.lower.z;
  b.upper[0] = r.upper.x; b.upper[1] = r.upper.y; b.upper[2] = r.upper.z;
  return b;
  OSP_CATCH(b)
}

void ospSetParam(OSPObject o, const char *id, OSPDataType type, const void *mem)
{
  OSP_TRY
  if (!o) throw std::runtime_error("ospSetParam: NULL object");
  if (!id) throw std::runtime_error("ospSetParam: NULL parameter name");
  dev().setObjectParam(o, id, type, mem);
  OSP_CATCH()
}
void ospRemoveParam(OSPObject o, const char *id)
{
  OSP_TRY
  if (!o || !id) throw std::runtime_error("ospRemoveParam: NULL argument");
  dev().removeObjectParam(o, id);
  OSP_CATCH()
}
void ospCommit(OSPObject o)
{
  OSP_TRY
  if (!o) throw std::runtime_error("ospCommit: NULL object");
  dev().commit(o);
  OSP_CATCH()
}
void ospRelease(OSPObject o)
{
  OSP_TRY
  if (!o) return;
  dev().release(o);
  OSP_CATCH()
}
void ospRetain(OSPObject o)
{
  OSP_TRY
  if (!o) return;
  dev().retain(o);
  OSP_CATCH()
}

OSPFrameBuffer ospNewFrameBuffer(int sx, int sy, OSPFrameBufferFormat mode, uint32_t channels)
{
  OSP_TRY
  uint32_t ch = channels;
  if (mode == OSP_FB_NONE) ch &= ~OSP_FB_COLOR;              // API.cpp:572-574
  if (!(ch & OSP_FB_ACCUM)) ch &= ~OSP_FB_VARIANCE;          // API.cpp:575-578
  vec2i size{sx, sy};
  return dev().frameBufferCreate(size, mode, ch);
  OSP_CATCH(nullptr)
}
const void *ospMapFrameBuffer(OSPFrameBuffer fb, OSPFrameBufferChannel ch)
{
  OSP_TRY
  if (!fb) throw std::runtime_error("ospMapFrameBuffer: NULL framebuffer");
  return dev().frameBufferMap(fb, ch);
  OSP_CATCH(nullptr)
}
void ospUnmapFrameBuffer(const void *mapped, OSPFrameBuffer fb)
{
  OSP_TRY
  if (!fb) throw std::runtime_error("ospUnmapFrameBuffer: NULL framebuffer");
  dev().frameBufferUnmap(mapped, fb);
  OSP_CATCH()
}
float ospGetVariance(OSPFrameBuffer fb)
{
  OSP_TRY
  if (!fb) throw std::runtime_error("ospGetVariance: NULL framebuffer");
  return dev().getVariance(fb);
  OSP_CATCH(INFINITY)
}
void ospResetAccumulation(OSPFrameBuffer fb)
{
  OSP_TRY
  if (!fb) throw std::runtime_error("ospResetAccumulation: NULL framebuffer");
  dev().resetAccumulation(fb);
  OSP_CATCH()
}

OSPFuture ospRenderFrame(OSPFrameBuffer fb, OSPRenderer r, OSPCamera c, OSPWorld w)
{
  OSP_TRY
  if (!fb || !r || !c || !w) throw std::runtime_error("ospRenderFrame: NULL argument");
  return dev().renderFrame(fb, r, c, w);
  OSP_CATCH(nullptr)
}
int ospIsReady(OSPFuture f, OSPSyncEvent e)
{
  OSP_TRY
  if (!f) throw std::runtime_error("ospIsReady: NULL future");
  return dev().isReady(f, e);
  OSP_CATCH(1)
}
void ospWait(OSPFuture f, OSPSyncEvent e)
{
  OSP_TRY
  if (!f) throw std::runtime_error("ospWait: NULL future");
  if (e == OSP_NONE_FINISHED) return;
  dev().wait(f, e);
  OSP_CATCH()
}
void ospCancel(OSPFuture f)
{
  OSP_TRY
  if (!f) throw std::runtime_error("ospCancel: NULL future");
  dev().cancel(f);
  OSP_CATCH()
}
float ospGetProgress(OSPFuture f)
{
  OSP_TRY
  if (!f) throw std::runtime_error("ospGetProgress: NULL future");
  return dev().getProgress(f);
  OSP_CATCH(1.f)
}
float ospGetTaskDuration(OSPFuture f)
{
  OSP_TRY
  if (!f) throw std::runtime_error("ospGetTaskDuration: NULL future");
  return dev().getTaskDuration(f);
  OSP_CATCH(0.f)
}
void ospPick(OSPPickResult *result, OSPFrameBuffer fb, OSPRenderer r, OSPCamera c, OSPWorld w, float x, float y)
{
  OSP_TRY
  if (!result) throw std::runtime_error("ospPick: NULL result");
  result->hasHit = 0; result->instance = nullptr; result->model = nullptr; result->primID = 0xFFFFFFFFu;
  if (!fb || !r || !c || !w) return;
  vec2f p{x, y};
  *result = dev().pick(fb, r, c, w, p);
  OSP_CATCH()
}

// ---- ospray_util (ospray_util_impl.cpp:7-192) ---------------------------------------------
OSPData ospNewSharedData1D(const void *d, OSPDataType t, uint64_t n) { return ospNewSharedData(d, t, n, 0, 1, 0, 1, 0, nullptr, nullptr); }
OSPData ospNewAssignedData1D(const void *d, OSPDataType t, uint64_t n, OSPDeleterCallback cb, const void *up)
{
  return ospNewSharedData(d, t, n, 0, 1, 0, 1, 0, cb, up);
}
OSPData ospNewSharedData1DStride(const void *d, OSPDataType t, uint64_t n, int64_t s) { return ospNewSharedData(d, t, n, s, 1, 0, 1, 0, nullptr, nullptr); }
OSPData ospNewSharedData2D(const void *d, OSPDataType t, uint64_t n1, uint64_t n2) { return ospNewSharedData(d, t, n1, 0, n2, 0, 1, 0, nullptr, nullptr); }
OSPData ospNewSharedData2DStride(const void *d, OSPDataType t, uint64_t n1, int64_t s1, uint64_t n2, int64_t s2)
{
  return ospNewSharedData(d, t, n1, s1, n2, s2, 1, 0, nullptr, nullptr);
}
OSPData ospNewSharedData3D(const void *d, OSPDataType t, uint64_t n1, uint64_t n2, uint64_t n3)
{
  return ospNewSharedData(d, t, n1, 0, n2, 0, n3, 0, nullptr, nullptr);
}
OSPData ospNewData1D(OSPDataType t, uint64_t n) { return ospNewData(t, n, 1, 1); }
OSPData ospNewData2D(OSPDataType t, uint64_t n1, uint64_t n2) { return ospNewData(t, n1, n2, 1); }
void ospCopyData1D(const OSPData s, OSPData d, uint64_t i) { ospCopyData(s, d, i, 0, 0); }
void ospCopyData2D(const OSPData s, OSPData d, uint64_t i1, uint64_t i2) { ospCopyData(s, d, i1, i2, 0); }

void ospSetString(OSPObject o, const char *n, const char *s) { ospSetParam(o, n, OSP_STRING, s); }
void ospSetObject(OSPObject o, const char *n, OSPObject obj) { ospSetParam(o, n, OSP_OBJECT, &obj); }
void ospSetBool(OSPObject o, const char *n, int x) { bool b = x != 0; ospSetParam(o, n, OSP_BOOL, &b); }
void ospSetFloat(OSPObject o, const char *n, float x) { ospSetParam(o, n, OSP_FLOAT, &x); }
void ospSetInt(OSPObject o, const char *n, int x) { ospSetParam(o, n, OSP_INT, &x); }
void ospSetUInt(OSPObject o, const char *n, unsigned int x) { ospSetParam(o, n, OSP_UINT, &x); }
void ospSetBox1f(OSPObject o, const char *n, float x, float y) { float v[] = {x, y}; ospSetParam(o, n, OSP_BOX1F, v); }
void ospSetLinear2f(OSPObject o, const char *n, float x, float y, float z, float w) { float v[] = {x, y, z, w}; ospSetParam(o, n, OSP_LINEAR2F, v); }
void ospSetVec2f(OSPObject o, const char *n, float x, float y) { float v[] = {x, y}; ospSetParam(o, n, OSP_VEC2F, v); }
void ospSetVec3f(OSPObject o, const char *n, float x, float y, float z) { float v[] = {x, y, z}; ospSetParam(o, n, OSP_VEC3F, v); }
void ospSetVec4f(OSPObject o, const char *n, float x, float y, float z, float w) { float v[] = {x, y, z, w}; ospSetParam(o, n, OSP_VEC4F, v); }
void ospSetVec2i(OSPObject o, const char *n, int x, int y) { int v[] = {x, y}; ospSetParam(o, n, OSP_VEC2I, v); }
void ospSetVec3i(OSPObject o, const char *n, int x, int y, int z) { int v[] = {x, y, z}; ospSetParam(o, n, OSP_VEC3I, v); }
void ospSetVec4i(OSPObject o, const char *n, int x, int y, int z, int w) { int v[] = {x, y, z, w}; ospSetParam(o, n, OSP_VEC4I, v); }
void ospSetObjectAsData(OSPObject target, const char *n, OSPDataType type, OSPObject obj)
{
  OSPData tmp = ospNewSharedData(&obj, type, 1, 0, 1, 0, 1, 0, nullptr, nullptr);
  OSPData data = ospNewData(type, 1, 1, 1);
  ospCopyData(tmp, data, 0, 0, 0);
  ospSetParam(target, n, OSP_DATA, &data);
  ospRelease(tmp);
  ospRelease(data);
}
float ospRenderFrameBlocking(OSPFrameBuffer fb, OSPRenderer r, OSPCamera c, OSPWorld w)
{
  OSPFuture f = ospRenderFrame(fb, r, c, w);
  if (!f) return INFINITY;
  ospWait(f, OSP_TASK_FINISHED);
  ospRelease(f);
  return ospGetVariance(fb);
}

} // extern "C"
