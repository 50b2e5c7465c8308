// reference: ospray/api/Device.cpp:33-148 (createDevice, commit: params <- env), ospray/common/
// OSPCommon.cpp:749-838 (module loading, status / error plumbing)
#include "Device.h"
#include <dlfcn.h>
#include <cstdlib>
#include <iostream>
#include <mutex>

namespace ospray {
namespace api {

Ref<Device> Device::current;
uint32_t Device::logLevel = OSP_LOG_NONE;

std::map<std::string, Device::FactoryFcn> &Device::registry()
{
  static std::map<std::string, FactoryFcn> r;
  return r;
}

Device *Device::createDevice(const char *type)
{
  auto &r = registry();
  auto it = r.find(type);
  if (it == r.end())
    throw std::runtime_error(std::string("Could not find device of type: ") + type
        + ".  Make sure you have the correct OSPRay libraries linked and initialized.");
  return it->second();
}

int64_t Device::getProperty(const OSPDeviceProperty prop)
{
  switch (prop) {
  case OSP_DEVICE_VERSION: return OSPRAY_VERSION;
  case OSP_DEVICE_VERSION_MAJOR: return OSPRAY_VERSION_MAJOR;
  case OSP_DEVICE_VERSION_MINOR: return OSPRAY_VERSION_MINOR;
  case OSP_DEVICE_VERSION_PATCH: return OSPRAY_VERSION_PATCH;
  case OSP_DEVICE_SO_VERSION: return OSPRAY_SOVERSION;
  default: handleError(OSP_INVALID_ARGUMENT, "unknown readable device property"); return 0;
  }
}

static int logLevelFromString(const std::string &s, int def)
{
  if (s == "debug") return OSP_LOG_DEBUG;
  if (s == "info") return OSP_LOG_INFO;
  if (s == "warning") return OSP_LOG_WARNING;
  if (s == "error") return OSP_LOG_ERROR;
  if (s == "none") return OSP_LOG_NONE;
  return def;
}

void Device::commit()
{
  debugMode = getParam<bool>("debug", false);
  int ll = getParam<int>("logLevel", (int)getParam<uint32_t>("logLevel", OSP_LOG_NONE));
  numThreads = getParam<int>("numThreads", -1);
  warningsAreErrors = getParam<bool>("warnAsError", false);
  std::string dst = getParamString("logOutput", "");
  std::string edst = getParamString("errorOutput", "");
  getParam<int>("setAffinity", 0);
  getParam<bool>("traceApi", false);
  // environment overrides (Device.cpp:96-128)
  if (const char *e = getenv("OSPRAY_LOG_LEVEL")) ll = logLevelFromString(e, ll);
  if (const char *e = getenv("OSPRAY_DEBUG")) debugMode = atoi(e) != 0;
  if (const char *e = getenv("OSPRAY_NUM_THREADS")) numThreads = atoi(e);
  if (const char *e = getenv("OSPRAY_WARN_AS_ERROR")) warningsAreErrors = atoi(e) != 0;
  if (const char *e = getenv("OSPRAY_LOG_OUTPUT")) dst = e;
  if (const char *e = getenv("OSPRAY_ERROR_OUTPUT")) edst = e;
  if (debugMode) { ll = OSP_LOG_DEBUG; numThreads = 1; }
  logLevel = (uint32_t)ll;
  if (dst == "cout") msg_fcn = [](void *, const char *m) { std::cout << m; };
  else if (dst == "cerr") msg_fcn = [](void *, const char *m) { std::cerr << m; };
  if (edst == "cout") error_fcn = [](void *, OSPError e, const char *m) { std::cout << "#ospray: ERROR " << (int)e << ": " << m << std::endl; };
  else if (edst == "cerr") error_fcn = [](void *, OSPError e, const char *m) { std::cerr << "#ospray: ERROR " << (int)e << ": " << m << std::endl; };
  committed = true;
}

bool deviceIsSet() { return Device::current.get() != nullptr; }
Device &currentDevice() { return *Device::current; }

} // namespace api

void postStatusMsg(uint32_t level, const std::string &msg)
{
  if (!api::deviceIsSet()) {
    if (level >= OSP_LOG_ERROR) std::cerr << msg << std::endl;
    return;
  }
  api::Device &d = api::currentDevice();
  if (level == OSP_LOG_WARNING && d.warningsAreErrors) {
    handleError(OSP_UNKNOWN_ERROR, msg); // OSPCommon.cpp:800-822
    return;
  }
  if (level >= api::Device::logLevel) d.msg_fcn(d.statusUserData, (msg + '\n').c_str());
}

void handleError(OSPError e, const std::string &msg)
{
  if (api::deviceIsSet()) {
    api::Device &d = api::currentDevice();
    d.lastErrorCode = e;
    d.lastErrorMsg = "#ospray: " + msg;
    d.error_fcn(d.errorUserData, e, msg.c_str());
  } else {
    std::cerr << "#ospray: INITIALIZATION ERROR --> " << msg << std::endl;
  }
}

OSPError loadLocalModule(const std::string &name)
{
  static std::mutex m;
  static std::map<std::string, void *> loaded;
  std::lock_guard<std::mutex> lock(m);
  void *lib = nullptr;
  auto it = loaded.find(name);
  if (it != loaded.end()) lib = it->second;
  const std::string base = "libospray_module_" + name + ".so";
  if (!lib) {
    std::vector<std::string> candidates;
    if (const char *dir = getenv("OSPRAY_MODULE_PATH")) candidates.push_back(std::string(dir) + "/" + base);
    // next to libospray.so itself
    Dl_info info;
    if (dladdr((void *)&loadLocalModule, &info) && info.dli_fname) {
      std::string self(info.dli_fname);
      const size_t slash = self.find_last_of('/');
      if (slash != std::string::npos) candidates.push_back(self.substr(0, slash + 1) + base);
    }
    candidates.push_back(base + "." + std::to_string(OSPRAY_VERSION_MAJOR) + "." + std::to_string(OSPRAY_VERSION_MINOR) + "."
        + std::to_string(OSPRAY_VERSION_PATCH));
    candidates.push_back(base);
    std::string errs;
    for (const std::string &c : candidates) {
      lib = dlopen(c.c_str(), RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
      errs += std::string("\n  ") + dlerror();
    }
    if (!lib) {
      handleError(OSP_INVALID_OPERATION, "could not open module lib ospray_module_" + name + ":" + errs);
      return OSP_INVALID_OPERATION;
    }
    loaded[name] = lib;
  }
  const std::string initName = "ospray_module_init_" + name;
  using InitFcn = OSPError (*)(int16_t, int16_t, int16_t);
  InitFcn init = (InitFcn)dlsym(lib, initName.c_str());
  if (!init) {
    handleError(OSP_INVALID_OPERATION, "could not find module initializer " + initName);
    return OSP_INVALID_OPERATION;
  }
  const OSPError r = init(OSPRAY_VERSION_MAJOR, OSPRAY_VERSION_MINOR, OSPRAY_VERSION_PATCH);
  if (r != OSP_NO_ERROR) handleError(r, "initialization of module " + name + " failed");
  return r;
}

} // namespace ospray
