// reference: ospray/common/Managed.cpp:9-60, OSPCommon.cpp:225-373 (sizeOf), :375-560 (stringFor)
#include "Managed.h"
#include <sstream>

namespace ospray {

size_t sizeOf(OSPDataType type)
{
  if (type == OSP_DEVICE || type == OSP_VOID_PTR || type == OSP_STRING || isObjectType(type)) return sizeof(void *);
  if (type == OSP_BOOL) return sizeof(bool);
  struct Fam { uint32_t base; size_t elem; };
  static const Fam fams[] = {{OSP_CHAR, 1}, {OSP_UCHAR, 1}, {OSP_SHORT, 2}, {OSP_USHORT, 2}, {OSP_INT, 4}, {OSP_UINT, 4},
      {OSP_LONG, 8}, {OSP_ULONG, 8}, {OSP_HALF, 2}, {OSP_FLOAT, 4}, {OSP_DOUBLE, 8}};
  for (const Fam &f : fams)
    if (type >= f.base && type < f.base + 4) return f.elem * (type - f.base + 1);
  if (type >= OSP_BOX1I && type <= OSP_BOX4I) return 2 * 4 * (type - OSP_BOX1I + 1);
  if (type >= OSP_BOX1F && type <= OSP_BOX4F) return 2 * 4 * (type - OSP_BOX1F + 1);
  switch (type) {
  case OSP_LINEAR2F: return 4 * 4;
  case OSP_LINEAR3F: return 9 * 4;
  case OSP_AFFINE2F: return 6 * 4;
  case OSP_AFFINE3F: return 12 * 4;
  case OSP_QUATF: return 4 * 4;
  default: break;
  }
  std::stringstream e;
  e << "ospray::sizeOf: unknown OSPDataType " << (uint32_t)type;
  throw std::runtime_error(e.str());
}

std::string stringFor(OSPDataType type)
{
  switch (type) {
  case OSP_DATA: return "OSP_DATA";
  case OSP_CAMERA: return "OSP_CAMERA";
  case OSP_FRAMEBUFFER: return "OSP_FRAMEBUFFER";
  case OSP_GEOMETRIC_MODEL: return "OSP_GEOMETRIC_MODEL";
  case OSP_GEOMETRY: return "OSP_GEOMETRY";
  case OSP_GROUP: return "OSP_GROUP";
  case OSP_INSTANCE: return "OSP_INSTANCE";
  case OSP_LIGHT: return "OSP_LIGHT";
  case OSP_MATERIAL: return "OSP_MATERIAL";
  case OSP_RENDERER: return "OSP_RENDERER";
  case OSP_WORLD: return "OSP_WORLD";
  case OSP_FLOAT: return "OSP_FLOAT";
  case OSP_VEC2F: return "OSP_VEC2F";
  case OSP_VEC3F: return "OSP_VEC3F";
  case OSP_VEC4F: return "OSP_VEC4F";
  case OSP_INT: return "OSP_INT";
  case OSP_UINT: return "OSP_UINT";
  case OSP_VEC3UI: return "OSP_VEC3UI";
  case OSP_VEC4UI: return "OSP_VEC4UI";
  case OSP_BOOL: return "OSP_BOOL";
  case OSP_STRING: return "OSP_STRING";
  default: break;
  }
  std::stringstream s;
  s << "OSPDataType(" << (uint32_t)type << ")";
  return s.str();
}

ParameterizedObject::~ParameterizedObject()
{
  for (auto &p : params)
    if (p.second.obj) p.second.obj->refDec();
}

void ParameterizedObject::setParam(const std::string &name, OSPDataType type, const void *mem)
{
  if (type == OSP_UNKNOWN) throw std::runtime_error("cannot set OSP_UNKNOWN parameter type");
  Param np;
  np.type = type;
  if (type == OSP_STRING) {
    np.str = mem ? std::string((const char *)mem) : std::string();
  } else if (isObjectType(type)) {
    np.obj = mem ? *(ManagedObject *const *)mem : nullptr;
    if (np.obj) np.obj->refInc();
  } else {
    const size_t sz = sizeOf(type);
    if (sz > sizeof(np.bytes)) throw std::runtime_error("parameter too large");
    std::memcpy(np.bytes, mem, sz);
  }
  auto it = params.find(name);
  if (it != params.end()) {
    if (it->second.obj) it->second.obj->refDec(); // Managed.h:132-148
    it->second = np;
  } else
    params.emplace(name, np);
}

void ParameterizedObject::removeParam(const std::string &name)
{
  auto it = params.find(name);
  if (it == params.end()) return;
  if (it->second.obj) it->second.obj->refDec();
  params.erase(it);
}

std::string ParameterizedObject::getParamString(const std::string &name, const std::string &def)
{
  auto it = params.find(name);
  if (it == params.end() || it->second.type != OSP_STRING) return def;
  it->second.queried = true;
  return it->second.str;
}

ManagedObject *ParameterizedObject::getParamObjectRaw(const std::string &name)
{
  auto it = params.find(name);
  if (it == params.end() || !isObjectType(it->second.type)) return nullptr;
  it->second.queried = true;
  return it->second.obj;
}

std::vector<std::string> ParameterizedObject::unusedParams() const
{
  std::vector<std::string> r;
  for (auto &p : params)
    if (!p.second.queried) r.push_back(p.first);
  return r;
}

void ParameterizedObject::resetAllParamQueryStatus()
{
  for (auto &p : params) p.second.queried = false;
}

void ManagedObject::checkUnused()
{
  for (const std::string &n : unusedParams())
    postStatusMsg(OSP_LOG_WARNING,
        toString() + ": found unused (not queried) parameter '" + n + "' of type " + stringFor(params[n].type));
}

} // namespace ospray
