// Reference-counted, parameterised object base used by both libospray (api::Device) and the
// cuda module's object graph.
// reference: ospray/common/Managed.h:21-148, Managed.cpp:31-40 (checkUnused), rkcommon
// ParameterizedObject / RefCountedObject / Any (un-vendored; semantics restated: a parameter is
// found only when the stored OSPDataType equals the type asked for).
#pragma once
#include <atomic>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "ospray/ospray.h"
#include "osp_math.h"

namespace ospray {
using namespace ospm;

size_t sizeOf(OSPDataType type);                 // OSPCommon.cpp:225-373
std::string stringFor(OSPDataType type);
inline bool isObjectType(OSPDataType t) { return (t & OSP_OBJECT) != 0; }

void postStatusMsg(uint32_t level, const std::string &msg); // OSPCommon.cpp:784-822
void handleError(OSPError e, const std::string &msg);       // OSPCommon.cpp:824-838

struct RefCounted
{
  RefCounted() = default;
  virtual ~RefCounted() = default;
  void refInc() const { refs.fetch_add(1, std::memory_order_relaxed); }
  void refDec() const
  {
    if (refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete this;
  }
  long useCount() const { return refs.load(); }

 private:
  mutable std::atomic<long> refs{1};
};

template <typename T>
struct Ref
{
  T *ptr = nullptr;
  Ref() = default;
  Ref(T *p) : ptr(p) { if (ptr) ptr->refInc(); }
  Ref(const Ref &o) : ptr(o.ptr) { if (ptr) ptr->refInc(); }
  Ref &operator=(const Ref &o)
  {
    if (o.ptr) o.ptr->refInc();
    if (ptr) ptr->refDec();
    ptr = o.ptr;
    return *this;
  }
  Ref &operator=(T *p)
  {
    if (p) p->refInc();
    if (ptr) ptr->refDec();
    ptr = p;
    return *this;
  }
  ~Ref() { if (ptr) ptr->refDec(); }
  T *operator->() const { return ptr; }
  T &operator*() const { return *ptr; }
  explicit operator bool() const { return ptr != nullptr; }
  T *get() const { return ptr; }
};

template <typename T> struct OSPTypeFor { static constexpr OSPDataType value = OSP_UNKNOWN; };
#define OSP_TYPEFOR(T, V) template <> struct OSPTypeFor<T> { static constexpr OSPDataType value = V; };
OSP_TYPEFOR(bool, OSP_BOOL)
OSP_TYPEFOR(int, OSP_INT)
OSP_TYPEFOR(uint32_t, OSP_UINT)
OSP_TYPEFOR(int64_t, OSP_LONG)
OSP_TYPEFOR(uint64_t, OSP_ULONG)
OSP_TYPEFOR(float, OSP_FLOAT)
OSP_TYPEFOR(vec2f, OSP_VEC2F)
OSP_TYPEFOR(vec3f, OSP_VEC3F)
OSP_TYPEFOR(vec4f, OSP_VEC4F)
OSP_TYPEFOR(vec2i, OSP_VEC2I)
OSP_TYPEFOR(vec3i, OSP_VEC3I)
OSP_TYPEFOR(box1f, OSP_BOX1F)
OSP_TYPEFOR(box3f, OSP_BOX3F)
OSP_TYPEFOR(linear3f, OSP_LINEAR3F)
OSP_TYPEFOR(affine3f, OSP_AFFINE3F)
OSP_TYPEFOR(void *, OSP_VOID_PTR)
OSP_TYPEFOR(uint8_t, OSP_UCHAR)
OSP_TYPEFOR(vec2ui, OSP_VEC2UI)
OSP_TYPEFOR(linear2f, OSP_LINEAR2F)
OSP_TYPEFOR(vec3ui, OSP_VEC3UI)
OSP_TYPEFOR(vec4ui, OSP_VEC4UI)
#undef OSP_TYPEFOR

struct ManagedObject;

struct Param
{
  OSPDataType type = OSP_UNKNOWN;
  uint8_t bytes[64] = {0};
  ManagedObject *obj = nullptr; // retained
  std::string str;
  bool queried = false;
};

struct ParameterizedObject
{
  virtual ~ParameterizedObject();
  // `mem` as handed to ospSetParam: pointer to the value; for OSP_STRING the char* itself;
  // for object types a pointer to the handle (ISPCDevice.cpp:52-70)
  void setParam(const std::string &name, OSPDataType type, const void *mem);
  void removeParam(const std::string &name);
  bool hasParam(const std::string &name) const { return params.count(name) != 0; }

  template <typename T>
  T getParam(const std::string &name, T def)
  {
    auto it = params.find(name);
    if (it == params.end() || it->second.type != OSPTypeFor<T>::value) return def;
    it->second.queried = true;
    T v;
    std::memcpy(&v, it->second.bytes, sizeof(T));
    return v;
  }
  std::string getParamString(const std::string &name, const std::string &def);
  // object of any handle type; marks queried. Null when absent or not an object.
  ManagedObject *getParamObjectRaw(const std::string &name);
  std::vector<std::string> unusedParams() const;
  void resetAllParamQueryStatus();

  std::map<std::string, Param> params;
};

struct ManagedObject : public RefCounted, public ParameterizedObject
{
  ~ManagedObject() override = default;
  virtual void commit() {}
  virtual std::string toString() const { return "ospray::ManagedObject"; }
  virtual box3f getBounds() const { return box_empty(); }
  void checkUnused(); // Managed.cpp:31-40
  OSPDataType managedObjectType = OSP_OBJECT;

  template <typename T>
  T *getParamObject(const std::string &name, T *def = nullptr)
  {
    ManagedObject *o = getParamObjectRaw(name);
    T *t = o ? dynamic_cast<T *>(o) : nullptr;
    return t ? t : def;
  }
};

} // namespace ospray
