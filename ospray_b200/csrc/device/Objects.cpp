// Object graph implementation (scene side). See Objects.h for the reference map.
#include "Objects.h"
#include <algorithm>
#include <cmath>
#include <numeric>
#include "CudaDevice.h"

namespace ospray {
namespace cuda {

void checkCuda(int rc, const char *what)
{
  if (rc != 0) throw std::runtime_error(std::string("cuda device: ") + what + " failed: " + ospcuda_last_error());
}

// ---- DevMem ---------------------------------------------------------------------------------
void DevMem::release()
{
  if (p) ospcuda_free(p);
  p = nullptr;
  bytes = 0;
}
void DevMem::reserve(size_t n)
{
  if (n <= bytes && p) return;
  release();
  if (n == 0) return;
  checkCuda(ospcuda_malloc(&p, n), "device allocation");
  bytes = n;
}
void DevMem::upload(const void *src, size_t n)
{
  reserve(n);
  if (n) checkCuda(ospcuda_memcpy_h2d(p, src, n), "host-to-device copy");
}

// ---- Data (modules/cpu/common/Data.cpp:12-201) -------------------------------------------------
Data::Data(const void *sharedData, OSPDataType t, const vec3ul &n, const vec3l &s, OSPDeleterCallback cb, const void *up)
    : addr((char *)sharedData), shared(true), type(t), numItems(n), byteStride(s), freeFunction(cb), userData(up)
{
  if (sharedData == nullptr) throw std::runtime_error("OSPData: shared buffer is NULL");
  init();
  if (isObjectType(type))
    for (size_t z = 0; z < numItems.z; z++)
      for (size_t y = 0; y < numItems.y; y++)
        for (size_t x = 0; x < numItems.x; x++) {
          ManagedObject *c = *(ManagedObject **)data(vec3ul{x, y, z});
          if (c) c->refInc();
        }
}

Data::Data(OSPDataType t, const vec3ul &n) : shared(false), type(t), numItems(n), byteStride{0, 0, 0}
{
  owned.resize(size() * sizeOf(type) + 16, 0);
  addr = owned.data();
  init();
}

Data::~Data()
{
  if (isObjectType(type))
    for (size_t z = 0; z < numItems.z; z++)
      for (size_t y = 0; y < numItems.y; y++)
        for (size_t x = 0; x < numItems.x; x++) {
          ManagedObject *c = *(ManagedObject **)data(vec3ul{x, y, z});
          if (c) c->refDec();
        }
  if (freeFunction) freeFunction(userData, addr);
}

void Data::init()
{
  managedObjectType = OSP_DATA;
  if (std::min(numItems.x, std::min(numItems.y, numItems.z)) == 0)
    throw std::out_of_range("OSPData: all numItems must be positive");
  dimensions = std::max(1, (int)(numItems.x > 1) + (int)(numItems.y > 1) + (int)(numItems.z > 1));
  if (byteStride.x == 0) byteStride.x = (int64_t)sizeOf(type);
  if (byteStride.y == 0) byteStride.y = (int64_t)numItems.x * byteStride.x;
  if (byteStride.z == 0) byteStride.z = (int64_t)numItems.y * byteStride.y;
  // dominant axis of 1-D arrays (Data.cpp:139-155)
  stride1D = byteStride.x;
  numItems1D = numItems.x;
  if (numItems.y > 1) { stride1D = byteStride.y; numItems1D = numItems.y; }
  else if (numItems.z > 1) { stride1D = byteStride.z; numItems1D = numItems.z; }
  if (dimensions != 1) { stride1D = byteStride.x; numItems1D = size(); }
}

bool Data::compact() const
{
  return data(vec3ul{0, 0, 0}) + sizeOf(type) * (size() - 1) == data(vec3ul{numItems.x - 1, numItems.y - 1, numItems.z - 1});
}

void Data::copy(const Data &source, const vec3ul &di)
{
  if (type != source.type && !(type == OSP_OBJECT && isObjectType(source.type)))
    throw std::runtime_error(toString() + "::copy: types must match (cannot copy '" + stringFor(source.type) + "' into '"
        + stringFor(type) + "')");
  if (shared && !source.shared) throw std::runtime_error("OSPData::copy: cannot copy opaque (non-shared) data into shared data");
  if (numItems.x < di.x + source.numItems.x || numItems.y < di.y + source.numItems.y || numItems.z < di.z + source.numItems.z)
    throw std::out_of_range("OSPData::copy: source does not fit into destination");
  const size_t es = sizeOf(type);
  for (size_t z = 0; z < source.numItems.z; z++)
    for (size_t y = 0; y < source.numItems.y; y++)
      for (size_t x = 0; x < source.numItems.x; x++) {
        char *dst = data(vec3ul{x + di.x, y + di.y, z + di.z});
        const char *src = source.data(vec3ul{x, y, z});
        if (dst == src) continue;
        if (isObjectType(type)) {
          ManagedObject *s = *(ManagedObject *const *)src;
          if (s) s->refInc();
          ManagedObject **d = (ManagedObject **)dst;
          if (*d) (*d)->refDec();
          *d = s;
        } else
          std::memcpy(dst, src, es);
      }
  deviceDirty = true;
}

OcData1D Data::deviceView1D()
{
  OcData1D v{};
  const size_t es = sizeOf(type);
  if (deviceDirty || !dev.p || devIsCompact) {
    if (stride1D >= (int64_t)es) {
      // upload the strided span as is
      dev.upload(addr, (size_t)stride1D * (numItems1D - 1) + es);
      devIsCompact = false;
    } else {
      // negative / overlapping strides: compact on the host
      std::vector<char> tmp(numItems1D * es);
      for (size_t i = 0; i < numItems1D; i++) std::memcpy(tmp.data() + i * es, data(i), es);
      dev.upload(tmp.data(), tmp.size());
      devIsCompact = true;
    }
    deviceDirty = false;
  }
  v.addr = dev.as<uint8_t>();
  v.byteStride = devIsCompact ? (int64_t)es : stride1D;
  v.numItems = (uint32_t)numItems1D;
  return v;
}

const void *Data::deviceCompact()
{
  const size_t es = sizeOf(type);
  if (deviceDirty || !dev.p || !devIsCompact) {
    if (compact()) {
      dev.upload(addr, size() * es);
    } else {
      std::vector<char> tmp(size() * es);
      size_t o = 0;
      for (size_t z = 0; z < numItems.z; z++)
        for (size_t y = 0; y < numItems.y; y++)
          for (size_t x = 0; x < numItems.x; x++, o += es) std::memcpy(tmp.data() + o, data(vec3ul{x, y, z}), es);
      dev.upload(tmp.data(), tmp.size());
    }
    devIsCompact = true;
    deviceDirty = false;
  }
  return dev.p;
}

Data *getParamData(ManagedObject &o, const std::string &name, OSPDataType elemType, bool required, bool promoteScalar,
    Ref<Data> *promoted)
{
  auto it = o.params.find(name);
  if (it != o.params.end()) {
    Param &p = it->second;
    if (isObjectType(p.type) && p.obj) {
      Data *d = dynamic_cast<Data *>(p.obj);
      if (d && (d->type == elemType || (isObjectType(elemType) && isObjectType(d->type) && (elemType == OSP_OBJECT || d->type == OSP_OBJECT)))
          && d->dimensions == 1) {
        p.queried = true;
        return d;
      }
      // single object promoted to an array of one (GeometricModel.cpp:27)
      if (!d && promoteScalar && promoted && isObjectType(elemType) && (p.type == elemType || p.type == OSP_OBJECT || elemType == OSP_OBJECT)) {
        p.queried = true;
        Data *nd = new Data(elemType, vec3ul{1, 1, 1});
        *(ManagedObject **)nd->addr = p.obj;
        p.obj->refInc();
        *promoted = nd;
        nd->refDec();
        return promoted->get();
      }
    } else if (promoteScalar && promoted && p.type == elemType && !isObjectType(elemType)) {
      p.queried = true;
      Data *nd = new Data(elemType, vec3ul{1, 1, 1});
      std::memcpy(nd->addr, p.bytes, sizeOf(elemType));
      *promoted = nd;
      nd->refDec();
      return promoted->get();
    }
  }
  if (required)
    throw std::runtime_error(o.toString() + " must have '" + name + "' array with element type " + stringFor(elemType));
  return nullptr;
}

// ---- Camera (camera/Camera.cpp:33-126, PerspectiveCamera.cpp:26-117) ------------------------------
Camera::Camera(const std::string &t) : type(t) { managedObjectType = OSP_CAMERA; }

void Camera::commit()
{
  vec3f pos = getParam<vec3f>("position", make_vec3f(0.f));
  vec3f dir = getParam<vec3f>("direction", make_vec3f(0.f, 0.f, 1.f));
  vec3f up = getParam<vec3f>("up", make_vec3f(0.f, 1.f, 0.f));
  const float nearClip = std::max(getParam<float>("nearClip", 1e-6f), 0.f);
  const vec2f imageStart = getParam<vec2f>("imageStart", make_vec2f(0.f, 0.f));
  const vec2f imageEnd = getParam<vec2f>("imageEnd", make_vec2f(1.f, 1.f));
  getParam<box1f>("shutter", box1f{0.5f, 0.5f});
  getParam<uint32_t>("shutterType", 0);
  getParam<float>("rollingShutterDuration", 0.f);
  const affine3f xfm = getParam<affine3f>("transform", getParam<affine3f>("xfm", affine_identity()));
  pos = xfmPoint(xfm, pos);
  dir = normalize(xfmVector(xfm, dir));
  up = normalize(xfmVector(xfm, up));

  if (type == "panoramic") { // camera/PanoramicCamera.cpp:29-53
    sh = OcCamera{};
    sh.type = OC_CAMERA_PANORAMIC;
    const vec3f vz = -normalize(dir);
    const vec3f vx = normalize(cross(up, vz));
    const vec3f vy = cross(vz, vx);
    sh.org = toOc(pos);
    sh.du_size = toOc(vx); sh.dv_up = toOc(vy); sh.dir_00 = toOc(vz); sh.direction = toOc(dir);
    sh.nearClip = nearClip;
    sh.subImage[0] = imageStart.x; sh.subImage[1] = imageStart.y; sh.subImage[2] = imageEnd.x; sh.subImage[3] = imageEnd.y;
    sh.stereoMode = getParam<uint32_t>("stereoMode", OSP_STEREO_NONE);
    float ipd = 0.5f * getParam<float>("interpupillaryDistance", 0.0635f);
    if (sh.stereoMode == OSP_STEREO_TOP_BOTTOM) ipd = -ipd; // left eye at the top (image origin at the lower left)
    sh.ipdOffset = toOc(make_vec3f(ipd, 0.f, 0.f));
    return;
  }
  if (type == "orthographic") { // camera/OrthographicCamera.cpp:25-49
    const float height = getParam<float>("height", 1.f);
    const float oaspect = getParam<float>("aspect", 1.f);
    const vec3f dn = normalize(dir);
    vec3f du = normalize(cross(dn, up));
    const vec3f dv = cross(du, dn) * height;
    du = du * (height * oaspect);
    sh = OcCamera{};
    sh.type = OC_CAMERA_ORTHOGRAPHIC;
    sh.org = toOc(pos - 0.5f * du - 0.5f * dv);
    sh.dir_00 = toOc(dn); sh.du_size = toOc(du); sh.dv_up = toOc(dv); sh.direction = toOc(dir);
    sh.nearClip = nearClip;
    sh.subImage[0] = imageStart.x; sh.subImage[1] = imageStart.y; sh.subImage[2] = imageEnd.x; sh.subImage[3] = imageEnd.y;
    sh.aspect = oaspect; sh.imgPlaneSizeY = height;
    return;
  }
  const float fovy = getParam<float>("fovy", 60.f);
  float aspect = getParam<float>("aspect", 1.f);
  const float apertureRadius = getParam<float>("apertureRadius", 0.f);
  const float focusDistance = getParam<float>("focusDistance", 1.f);
  const bool architectural = getParam<bool>("architectural", false);
  const uint32_t stereoMode = getParam<uint32_t>("stereoMode", OSP_STEREO_NONE);
  const float interpupillaryDistance = getParam<float>("interpupillaryDistance", 0.0635f);
  if (stereoMode == OSP_STEREO_SIDE_BY_SIDE) aspect *= 0.5f; // PerspectiveCamera.cpp:39-48
  else if (stereoMode == OSP_STEREO_TOP_BOTTOM) aspect *= 2.f;

  const float imgPlaneY = 2.f * tanf((0.5f * fovy) * (float)(M_PI / 180.0));
  const float imgPlaneX = imgPlaneY * aspect;
  const float scaledAperture = apertureRadius ? apertureRadius / (imgPlaneX * focusDistance) : 0.0f;
  vec3f dir_00 = normalize(dir);
  vec3f du = normalize(cross(dir_00, up));
  vec3f dv = architectural ? normalize(up) : cross(du, dir_00);
  vec3f ipdOffset = 0.5f * interpupillaryDistance * du; // :85-100
  if (stereoMode == OSP_STEREO_LEFT) pos = pos - ipdOffset;
  else if (stereoMode == OSP_STEREO_RIGHT) pos = pos + ipdOffset;
  else if (stereoMode == OSP_STEREO_TOP_BOTTOM) ipdOffset = -ipdOffset; // left eye at the top (image origin at the lower left)
  du = du * imgPlaneX;
  dv = dv * imgPlaneY;
  dir_00 = dir_00 - 0.5f * du - 0.5f * dv;
  if (scaledAperture > 0.f) {
    du = du * focusDistance;
    dv = dv * focusDistance;
    dir_00 = dir_00 * focusDistance;
  }
  sh.org = toOc(pos);
  sh.dir_00 = toOc(dir_00);
  sh.du_size = toOc(du);
  sh.dv_up = toOc(dv);
  sh.direction = toOc(dir);
  sh.nearClip = nearClip;
  sh.subImage[0] = imageStart.x; sh.subImage[1] = imageStart.y; sh.subImage[2] = imageEnd.x; sh.subImage[3] = imageEnd.y;
  sh.scaledAperture = scaledAperture;
  sh.aspect = aspect;
  sh.imgPlaneSizeY = imgPlaneY;
  sh.needLensSample = scaledAperture > 0.f;
  sh.needTimeSample = 0;
  sh.type = OC_CAMERA_PERSPECTIVE;
  sh.stereoMode = stereoMode;
  sh.ipdOffset = toOc(ipdOffset);
}

// ---- Geometry (geometry/Mesh.cpp:37-204, Spheres.cpp:46-106, Boxes.cpp:33-42) ----------------------
Geometry::Geometry(const std::string &t) : type(t) { managedObjectType = OSP_GEOMETRY; }

void Geometry::commit()
{
  sh = OcGeometry{};
  if (type == "mesh") {
    vertexData = getParamData(*this, "vertex.position", OSP_VEC3F, true);
    bool normalFV = true, colorFV = true, texFV = true;
    normalData = getParamData(*this, "normal", OSP_VEC3F);
    if (!normalData) { normalData = getParamData(*this, "vertex.normal", OSP_VEC3F); normalFV = false; }
    Data *col = getParamObject<Data>("color");
    if (!col) { col = getParamObject<Data>("vertex.color"); colorFV = false; }
    if (col && col->type != OSP_VEC3F && col->type != OSP_VEC4F) throw std::runtime_error("mesh colour must be vec3f or vec4f");
    colorData = col;
    texcoordData = getParamData(*this, "texcoord", OSP_VEC2F);
    if (!texcoordData) { texcoordData = getParamData(*this, "vertex.texcoord", OSP_VEC2F); texFV = false; }
    indexData = getParamData(*this, "index", OSP_VEC3UI);
    if (!indexData) indexData = getParamData(*this, "index", OSP_VEC4UI);
    if (!indexData) {
      const bool quadSoup = getParam<bool>("quadSoup", false);
      const int elements = quadSoup ? 4 : 3;
      Data *gen = new Data(quadSoup ? OSP_VEC4UI : OSP_VEC3UI, vec3ul{vertexData->size() / elements, 1, 1});
      uint32_t *b = (uint32_t *)gen->addr;
      std::iota(b, b + gen->size() * elements, 0u);
      indexData = gen;
      gen->refDec();
    }
    const bool isTri = indexData->type == OSP_VEC3UI;
    sh.type = isTri ? OC_GEOM_TRIANGLES : OC_GEOM_QUADS;
    sh.numPrims = (uint32_t)indexData->size();
    sh.index = indexData->deviceView1D();
    sh.vertex = vertexData->deviceView1D();
    if (normalData) sh.normal = normalData->deviceView1D();
    if (colorData) sh.color = colorData->deviceView1D();
    if (texcoordData) sh.texcoord = texcoordData->deviceView1D();
    sh.flags = (normalFV ? OC_GEOMFLAG_NORMAL_FACEVARYING : 0) | (colorFV ? OC_GEOMFLAG_COLOR_FACEVARYING : 0)
        | (texFV ? OC_GEOMFLAG_TEXCOORD_FACEVARYING : 0)
        | ((colorData && colorData->type == OSP_VEC4F) ? (OC_GEOMFLAG_HAS_ALPHA | OC_GEOMFLAG_COLOR_IS_VEC4) : 0);
    if (colorData && colorData->type == OSP_VEC4F) {
      // lets shadow rays treat such a mesh like one without alpha (pathtracer_wf.cuh ptOccluderIsOpaque)
      bool allOne = true;
      for (size_t i = 0; i < colorData->size() && allOne; i++) allOne = colorData->at<vec4f>(i).w == 1.0f;
      if (allOne) sh.flags |= OC_GEOMFLAG_ALPHA_ALL_ONE;
    }
  } else if (type == "sphere") {
    const uint32_t sphereType = getParam<uint32_t>("type", OSP_SPHERE);
    if (sphereType != OSP_SPHERE) throw std::runtime_error("cuda device: only OSP_SPHERE sphere geometry is supported");
    const float radius = getParam<float>("radius", 0.01f);
    sphereRadius = radius;
    vertexData = getParamData(*this, "sphere.position", OSP_VEC3F, true);
    radiusData = getParamData(*this, "sphere.radius", OSP_FLOAT);
    texcoordData = getParamData(*this, "sphere.texcoord", OSP_VEC2F);
    const size_t n = vertexData->size();
    std::vector<vec4f> inter(n);
    for (size_t i = 0; i < n; i++) {
      const vec3f p = vertexData->at<vec3f>(i);
      inter[i] = make_vec4f(p, radiusData ? radiusData->at<float>(i) : radius);
    }
    generated.upload(inter.data(), n * sizeof(vec4f));
    sh.type = OC_GEOM_SPHERES;
    sh.numPrims = (uint32_t)n;
    sh.sphere.addr = generated.as<uint8_t>();
    sh.sphere.byteStride = sizeof(vec4f);
    sh.sphere.numItems = (uint32_t)n;
    if (texcoordData) sh.texcoord = texcoordData->deviceView1D();
  } else if (type == "box") {
    boxData = getParamData(*this, "box", OSP_BOX3F, true);
    sh.type = OC_GEOM_BOXES;
    sh.numPrims = (uint32_t)boxData->size();
    sh.box = boxData->deviceView1D();
  } else if (type == "curve") {
    // geometry/Curves.cpp:84-194: round curves only (type OSP_ROUND), linear and Catmull-Rom bases; the segments are
    // tessellated into capsules by the BVH builder (include/ospcuda.h OC_CURVE_SUBDIV)
    vertexData = getParamData(*this, "vertex.position_radius", OSP_VEC4F, true);
    indexData = getParamData(*this, "index", OSP_UINT, true);
    colorData = getParamData(*this, "vertex.color", OSP_VEC4F);
    texcoordData = getParamData(*this, "vertex.texcoord", OSP_VEC2F);
    const uint32_t curveType = getParam<uint32_t>("type", OSP_UNKNOWN_CURVE_TYPE);
    const uint32_t basis = getParam<uint32_t>("basis", OSP_UNKNOWN_CURVE_BASIS);
    if (curveType != OSP_ROUND) throw std::runtime_error("cuda device: only OSP_ROUND curves are supported");
    if (basis != OSP_LINEAR && basis != OSP_CATMULL_ROM && basis != OSP_BEZIER && basis != OSP_BSPLINE)
      throw std::runtime_error("cuda device: curve basis must be OSP_LINEAR, OSP_BEZIER, OSP_BSPLINE or OSP_CATMULL_ROM (Hermite curves carry tangents and are not built)");
    const uint32_t need = basis == OSP_LINEAR ? 2u : 4u;
    for (size_t i = 0; i < indexData->size(); i++)
      if ((size_t)indexData->at<uint32_t>(i) + need > vertexData->size())
        throw std::runtime_error("cuda device: curve 'index' refers past the end of 'vertex.position_radius'");
    sh.type = OC_GEOM_CURVES;
    sh.numPrims = (uint32_t)indexData->size();
    sh.curveBasis = basis;
    sh.index = indexData->deviceView1D();
    sh.sphere = vertexData->deviceView1D();
    if (colorData) { sh.color = colorData->deviceView1D(); sh.flags = OC_GEOMFLAG_HAS_ALPHA | OC_GEOMFLAG_COLOR_IS_VEC4; }
    if (texcoordData) sh.texcoord = texcoordData->deviceView1D();
  } else
    throw std::runtime_error("cuda device: geometry type '" + type + "' is not supported");
  dev.upload(&sh, sizeof(sh));
}

// ---- Material (render/materials/OBJ.cpp:30-66) ---------------------------------------------------
Material::Material(CudaDevice &d, const std::string &t) : device(d), type(t), slot(d.allocMaterialSlot())
{
  managedObjectType = OSP_MATERIAL;
}
Material::~Material()
{
  if (wasEmissive) device.emissiveMaterials--;
  if (wasExtended) device.emissiveMaterials--;
  if (texD || texKd || texKs || texNs || texBump) device.texturedMaterials--;
  device.freeMaterialSlot(slot);
}

// LinearSpace2::orthogonal() of rkcommon (un-vendored; restated): the rotation part of a 2x2 matrix by the iteration
// m <- (m + m^-T) / 2, a mirrored x axis flipped first and restored at the end
static void linear2fOrthogonal(float m[4]) // vx = (m[0], m[1]), vy = (m[2], m[3])
{
  float mirror = 1.f;
  if (m[0] * m[3] - m[2] * m[1] < 0.f) { m[0] = -m[0]; m[1] = -m[1]; mirror = -1.f; }
  for (int i = 0; i < 99; i++) {
    // transposed(): rows become columns; inverse of [[a, c], [b, d]] (columns (a,b), (c,d))
    const float ta = m[0], tb = m[2], tc = m[1], td = m[3]; // transposed columns (ta, tb), (tc, td)
    const float det = ta * td - tc * tb, r = 1.f / det;
    const float ia = td * r, ib = -tb * r, ic = -tc * r, id = ta * r; // inverse columns (ia, ib), (ic, id)
    const float n[4] = {0.5f * (m[0] + ia), 0.5f * (m[1] + ib), 0.5f * (m[2] + ic), 0.5f * (m[3] + id)};
    const float d0 = n[0] - m[0], d1 = n[1] - m[1], d2 = n[2] - m[2], d3 = n[3] - m[3];
    for (int k = 0; k < 4; k++) m[k] = n[k];
    if (std::max(d0 * d0 + d1 * d1, d2 * d2 + d3 * d3) < 1e-8f) break;
  }
  m[0] *= mirror; m[1] *= mirror;
}

void Material::commit()
{
  OcMaterial m{};
  bool extended = false; // needs the extended-material kernels and the microfacet albedo tables
  const bool hadTextures = texD || texKd || texKs || texNs || texBump;
  texD = texKd = texKs = texNs = texBump = nullptr;
  texPrincipled.clear();
  if (type == "obj") {
    // render/materials/OBJ.cpp:30-66: a parameter with a map defaults to factor 1 (Material.cpp:110-133)
    m.type = OC_MATERIAL_OBJ;
    m.mapD = getTextureParam("map_d", texD);
    m.mapKd = getTextureParam("map_kd", texKd);
    m.mapKs = getTextureParam("map_ks", texKs);
    m.mapNs = getTextureParam("map_ns", texNs);
    m.mapBump = getTextureParam("map_bump", texBump);
    float d = getParam<float>("d", 1.f);
    vec3f Kd = getParam<vec3f>("kd", make_vec3f(texKd ? 1.f : 0.8f));
    vec3f Ks = getParam<vec3f>("ks", make_vec3f(texKs ? 1.f : 0.f));
    const float Ns = getParam<float>("ns", texNs ? 1.f : 10.f);
    vec3f Tf = getParam<vec3f>("tf", make_vec3f(0.f));
    const float total = reduce_max(Kd + Ks + Tf);
    if (total > 1.f) { Kd = Kd / total; Ks = Ks / total; Tf = Tf / total; }
    float rot[4] = {m.mapBump.xform2f[0], m.mapBump.xform2f[1], m.mapBump.xform2f[2], m.mapBump.xform2f[3]};
    linear2fOrthogonal(rot);
    m.bumpRot[0] = rot[0]; m.bumpRot[1] = rot[2]; m.bumpRot[2] = rot[1]; m.bumpRot[3] = rot[3]; // transposed()
    m.d = d; m.Kd = toOc(Kd); m.Ks = toOc(Ks); m.Ns = Ns; m.Tf = toOc(Tf);
  } else if (type == "thinGlass") {
    // render/materials/ThinGlass.cpp:33-48
    m.type = OC_MATERIAL_THINGLASS;
    const float eta = getParam<float>("eta", 1.5f);
    m.mapKd = getTextureParam("map_attenuationColor", texKd);
    const vec3f acf = getParam<vec3f>("attenuationColor", make_vec3f(1.f));
    const float attenuationDistance = getParam<float>("attenuationDistance", 1.f);
    const float thickness = getParam<float>("thickness", 1.f);
    m.eta = 1.f / eta;
    m.attenuationScale = thickness * (1.f / std::max(attenuationDistance, 1e-5f));
    m.attenuation = toOc(make_vec3f(std::log(acf.x), std::log(acf.y), std::log(acf.z)) * m.attenuationScale);
  } else if (type == "metal") {
    // render/materials/Metal.cpp:30-76, RGB form (eta / k per channel; default aluminium, math/spectrum.h:41-48)
    if (getParamObjectRaw("ior")) throw std::runtime_error("cuda device: metal with spectral 'ior' data is not supported (use eta / k)");
    m.type = OC_MATERIAL_METAL;
    m.mapNs = getTextureParam("map_roughness", texNs);
    m.Kd = toOc(getParam<vec3f>("eta", make_vec3f(1.47f, 0.984f, 0.602f)));
    m.Ks = toOc(getParam<vec3f>("k", make_vec3f(7.64f, 6.55f, 5.36f)));
    m.Ns = getParam<float>("roughness", texNs ? 1.f : 0.1f);
    extended = true;
  } else if (type == "plastic") {
    // render/materials/Plastic.cpp:28-38
    m.type = OC_MATERIAL_PLASTIC;
    m.Kd = toOc(getParam<vec3f>("pigmentColor", make_vec3f(1.f)));
    m.eta = 1.f / getParam<float>("eta", 1.4f);
    m.Ns = getParam<float>("roughness", 0.01f);
    extended = true;
  } else if (type == "glass") {
    // render/materials/Glass.cpp:35-63
    m.type = OC_MATERIAL_GLASS;
    const float etaInside = getParam<float>("etaInside", getParam<float>("eta", 1.5f));
    const float etaOutside = getParam<float>("etaOutside", 1.f);
    const vec3f aci = getParam<vec3f>("attenuationColorInside", getParam<vec3f>("attenuationColor", make_vec3f(1.f)));
    const vec3f aco = getParam<vec3f>("attenuationColorOutside", make_vec3f(1.f));
    const float dist = std::max(getParam<float>("attenuationDistance", 1.0f), 1e-5f);
    m.eta = etaInside;
    m.attenuation = toOc(make_vec3f(std::log(aci.x), std::log(aci.y), std::log(aci.z)) / dist);
    m.principled.outsideIor = etaOutside;
    m.principled.outsideAttenuation = toOc(make_vec3f(std::log(aco.x), std::log(aco.y), std::log(aco.z)) / dist);
    extended = true;
  } else if (type == "alloy") {
    // render/materials/Alloy.cpp:28-40
    m.type = OC_MATERIAL_ALLOY;
    m.mapKd = getTextureParam("map_color", texKd);
    m.mapKs = getTextureParam("map_edgeColor", texKs);
    m.mapNs = getTextureParam("map_roughness", texNs);
    m.Kd = toOc(getParam<vec3f>("color", make_vec3f(texKd ? 1.f : 0.9f)));
    m.Ks = toOc(getParam<vec3f>("edgeColor", make_vec3f(1.f)));
    m.Ns = getParam<float>("roughness", texNs ? 1.f : 0.1f);
    extended = true;
  } else if (type == "carPaint") {
    // render/materials/CarPaint.cpp:27-87; a factor with a map defaults to 1
    static const char *const mapNames[OC_CPMAP_COUNT] = {"baseColor", "roughness", "normal", "flakeColor", "flakeScale", "flakeDensity",
        "flakeSpread", "flakeJitter", "flakeRoughness", "coat", "coatIor", "coatColor", "coatThickness", "coatRoughness", "coatNormal",
        "flipflopColor", "flipflopFalloff"};
    OcCarPaint cp{};
    texPrincipled.assign(OC_CPMAP_COUNT, nullptr);
    for (int i = 0; i < OC_CPMAP_COUNT; i++) cp.map[i] = getTextureParam(std::string("map_") + mapNames[i], texPrincipled[i]);
    auto mapped = [&](int i) { return cp.map[i].tex != nullptr; };
    auto rotOf = [&](int i, float *dst) {
      float rot[4] = {cp.map[i].xform2f[0], cp.map[i].xform2f[1], cp.map[i].xform2f[2], cp.map[i].xform2f[3]};
      linear2fOrthogonal(rot);
      dst[0] = rot[0]; dst[1] = rot[2]; dst[2] = rot[1]; dst[3] = rot[3];
    };
    rotOf(OC_CPMAP_NORMAL, cp.normalRot); rotOf(OC_CPMAP_COATNORMAL, cp.coatNormalRot);
    auto f1 = [&](const char *name, int i, float def) { return getParam<float>(name, mapped(i) ? 1.f : def); };
    auto f3 = [&](const char *name, int i, float def) { return toOc(getParam<vec3f>(name, make_vec3f(mapped(i) ? 1.f : def))); };
    cp.baseColor = f3("baseColor", OC_CPMAP_BASECOLOR, 0.8f);
    cp.roughness = f1("roughness", OC_CPMAP_ROUGHNESS, 0.f);
    cp.normalScale = f1("normal", OC_CPMAP_NORMAL, 1.f);
    cp.useFlakeColor = hasParam("flakeColor") ? 1u : 0u;
    cp.flakeColor = f3("flakeColor", OC_CPMAP_FLAKECOLOR, 0.f);
    cp.flakeScale = f1("flakeScale", OC_CPMAP_FLAKESCALE, 100.f);
    cp.flakeDensity = f1("flakeDensity", OC_CPMAP_FLAKEDENSITY, 0.f);
    cp.flakeSpread = f1("flakeSpread", OC_CPMAP_FLAKESPREAD, 0.3f);
    cp.flakeJitter = f1("flakeJitter", OC_CPMAP_FLAKEJITTER, 0.75f);
    cp.flakeRoughness = f1("flakeRoughness", OC_CPMAP_FLAKEROUGHNESS, 0.3f);
    cp.coat = f1("coat", OC_CPMAP_COAT, 1.f);
    cp.coatIor = f1("coatIor", OC_CPMAP_COATIOR, 1.5f);
    cp.coatColor = f3("coatColor", OC_CPMAP_COATCOLOR, 1.f);
    cp.coatThickness = f1("coatThickness", OC_CPMAP_COATTHICKNESS, 1.f);
    cp.coatRoughness = f1("coatRoughness", OC_CPMAP_COATROUGHNESS, 0.f);
    cp.coatNormalScale = f1("coatNormal", OC_CPMAP_COATNORMAL, 1.f);
    cp.flipflopColor = f3("flipflopColor", OC_CPMAP_FLIPFLOPCOLOR, 1.f);
    cp.flipflopFalloff = f1("flipflopFalloff", OC_CPMAP_FLIPFLOPFALLOFF, 1.f);
    principledMaps.upload(&cp, sizeof(cp));
    m.type = OC_MATERIAL_CARPAINT;
    m.carPaint = principledMaps.as<OcCarPaint>();
    extended = true;
  } else if (type == "velvet") {
    // render/materials/Velvet.cpp:27-39
    m.type = OC_MATERIAL_VELVET;
    m.Kd = toOc(getParam<vec3f>("reflectance", make_vec3f(.4f, 0.f, 0.f)));
    m.Ns = getParam<float>("backScattering", .5f);
    m.Ks = toOc(getParam<vec3f>("horizonScatteringColor", make_vec3f(.75f, .1f, .1f)));
    m.eta = getParam<float>("horizonScatteringFallOff", 10.f);
    extended = true;
  } else if (type == "metallicPaint") {
    // render/materials/MetallicPaint.cpp:27-40 (default flake colour = aluminium's reflectivity, math/spectrum.h:50-53)
    m.type = OC_MATERIAL_METALLICPAINT;
    m.mapKd = getTextureParam("map_baseColor", texKd);
    const vec3f color = getParam<vec3f>("baseColor", make_vec3f(texKd ? 1.f : 0.8f));
    const float flakeAmount = getParam<float>("flakeAmount", 0.3f);
    const vec3f flakeColor = getParam<vec3f>("flakeColor", make_vec3f(0.909f, 0.916f, 0.923f));
    m.Kd = toOc(color * (1.f - flakeAmount));
    m.attenuationScale = flakeAmount;
    m.Ks = toOc(flakeColor * flakeAmount);
    m.Ns = getParam<float>("flakeSpread", 0.5f);
    m.eta = 1.f / getParam<float>("eta", 1.5f);
    extended = true;
  } else if (type == "principled") {
    // render/materials/Principled.cpp:33-150; a factor with a map defaults to 1 (getMaterialParam1f / 3f, Material.cpp:103-126)
    static const char *const mapNames[OC_PRMAP_COUNT] = {"baseColor", "edgeColor", "metallic", "diffuse", "specular", "ior", "transmission",
        "transmissionColor", "transmissionDepth", "roughness", "anisotropy", "rotation", "normal", "baseNormal", "coat", "coatIor", "coatColor",
        "coatThickness", "coatRoughness", "coatNormal", "sheen", "sheenColor", "sheenTint", "sheenRoughness", "opacity", "backlight",
        "thickness", "emissiveColor"};
    OcPrincipledMaps maps{};
    texPrincipled.assign(OC_PRMAP_COUNT, nullptr);
    bool anyMap = false;
    for (int i = 0; i < OC_PRMAP_COUNT; i++) {
      maps.map[i] = getTextureParam(std::string("map_") + mapNames[i], texPrincipled[i]);
      anyMap |= maps.map[i].tex != nullptr;
    }
    auto mapped = [&](int i) { return maps.map[i].tex != nullptr; };
    auto rotOf = [&](int i, float *dst) { // orthogonal().transposed() of the map's 2D linear part
      float rot[4] = {maps.map[i].xform2f[0], maps.map[i].xform2f[1], maps.map[i].xform2f[2], maps.map[i].xform2f[3]};
      linear2fOrthogonal(rot);
      dst[0] = rot[0]; dst[1] = rot[2]; dst[2] = rot[1]; dst[3] = rot[3];
    };
    rotOf(OC_PRMAP_NORMAL, maps.normalRot); rotOf(OC_PRMAP_BASENORMAL, maps.baseNormalRot); rotOf(OC_PRMAP_COATNORMAL, maps.coatNormalRot);
    maps.normalScale = getParam<float>("normal", 1.f);
    maps.baseNormalScale = getParam<float>("baseNormal", 1.f);
    maps.coatNormalScale = getParam<float>("coatNormal", 1.f);
    auto f1 = [&](const char *name, int i, float def) { return getParam<float>(name, mapped(i) ? 1.f : def); };
    auto f3 = [&](const char *name, int i, float def) { return getParam<vec3f>(name, make_vec3f(mapped(i) ? 1.f : def)); };
    m.type = OC_MATERIAL_PRINCIPLED;
    OcPrincipled &q = m.principled;
    q.baseColor = toOc(f3("baseColor", OC_PRMAP_BASECOLOR, 0.8f));
    q.edgeColor = toOc(f3("edgeColor", OC_PRMAP_EDGECOLOR, 1.f));
    q.metallic = f1("metallic", OC_PRMAP_METALLIC, 0.f);
    q.diffuse = f1("diffuse", OC_PRMAP_DIFFUSE, 1.f);
    q.specular = f1("specular", OC_PRMAP_SPECULAR, 1.f);
    q.specularMetallic = getParam<bool>("specularMetallic", true) ? 1u : 0u;
    q.ior = f1("ior", OC_PRMAP_IOR, 1.f);
    q.transmission = f1("transmission", OC_PRMAP_TRANSMISSION, 0.f);
    q.transmissionColor = toOc(f3("transmissionColor", OC_PRMAP_TRANSMISSIONCOLOR, 1.f));
    q.transmissionDepth = f1("transmissionDepth", OC_PRMAP_TRANSMISSIONDEPTH, 1.f);
    q.roughness = f1("roughness", OC_PRMAP_ROUGHNESS, 0.f);
    q.anisotropy = f1("anisotropy", OC_PRMAP_ANISOTROPY, 0.f);
    q.rotation = f1("rotation", OC_PRMAP_ROTATION, 0.f);
    q.coat = f1("coat", OC_PRMAP_COAT, 0.f);
    q.coatIor = f1("coatIor", OC_PRMAP_COATIOR, 1.5f);
    q.coatColor = toOc(f3("coatColor", OC_PRMAP_COATCOLOR, 1.f));
    q.coatThickness = f1("coatThickness", OC_PRMAP_COATTHICKNESS, 1.f);
    q.coatRoughness = f1("coatRoughness", OC_PRMAP_COATROUGHNESS, 0.f);
    q.opacity = f1("opacity", OC_PRMAP_OPACITY, 1.f);
    q.sheen = f1("sheen", OC_PRMAP_SHEEN, 0.f);
    q.sheenColor = toOc(f3("sheenColor", OC_PRMAP_SHEENCOLOR, 1.f));
    q.sheenTint = f1("sheenTint", OC_PRMAP_SHEENTINT, 0.f);
    q.sheenRoughness = f1("sheenRoughness", OC_PRMAP_SHEENROUGHNESS, 0.2f);
    q.thin = getParam<bool>("thin", false) ? 1u : 0u;
    q.backlight = f1("backlight", OC_PRMAP_BACKLIGHT, 0.f);
    q.thickness = f1("thickness", OC_PRMAP_THICKNESS, 1.f);
    const float outsideIor = getParam<float>("outsideIor", 1.f);
    const vec3f otc = getParam<vec3f>("outsideTransmissionColor", make_vec3f(1.f));
    const float otd = getParam<float>("outsideTransmissionDepth", 1.f);
    q.outsideIor = outsideIor >= 1.f ? outsideIor : 1.f / outsideIor;
    q.outsideAttenuation = toOc(make_vec3f(std::log(otc.x), std::log(otc.y), std::log(otc.z)) / otd);
    m.emission = toOc(f3("emissiveColor", OC_PRMAP_EMISSION, 0.f) * getParam<float>("intensity", 1.f));
    m.isEmissive = std::max(m.emission.x, std::max(m.emission.y, m.emission.z)) > 0.f ? 1u : 0u;
    if (anyMap) {
      principledMaps.upload(&maps, sizeof(maps));
      q.maps = principledMaps.as<OcPrincipledMaps>();
    }
    extended = true;
  } else if (type == "luminous") {
    // render/materials/Luminous.cpp:33-45: radiance = color * intensity (times map_color), isEmissive = any of it
    m.type = OC_MATERIAL_LUMINOUS;
    m.mapKd = getTextureParam("map_color", texKd);
    const vec3f color = getParam<vec3f>("color", make_vec3f(1.f));
    m.emission = toOc(color * getParam<float>("intensity", 1.f));
    m.transparency = getParam<float>("transparency", 0.f);
    m.isEmissive = std::max(m.emission.x, std::max(m.emission.y, m.emission.z)) > 0.f ? 1u : 0u;
  } else {
    m.type = OC_MATERIAL_OTHER;
    for (auto &p : params) p.second.queried = true;
    postStatusMsg(OSP_LOG_INFO, "cuda device: material '" + type + "' is not implemented; surfaces using it terminate paths");
  }
  if (extended) {
    uint32_t n = 0;
    checkCuda(ospcuda_microfacet_albedo_tables(&m.microfacetTables, &n), "microfacet albedo tables");
  }
  if (extended && !wasExtended) device.emissiveMaterials++; // (the counter of materials that need the extended kernels)
  if (!extended && wasExtended) device.emissiveMaterials--;
  wasExtended = extended;
  if (m.isEmissive && !wasEmissive) device.emissiveMaterials++;
  if (!m.isEmissive && wasEmissive) device.emissiveMaterials--;
  wasEmissive = m.isEmissive != 0;
  const bool hasTextures = texD || texKd || texKs || texNs || texBump;
  m.hasTextures = hasTextures ? 1u : 0u;
  m.cutout = (m.type == OC_MATERIAL_OBJ && (texD || (texKd && texKd->sh.hasAlpha))) ? 1u : 0u;
  if (hasTextures && !hadTextures) device.texturedMaterials++;
  if (!hasTextures && hadTextures) device.texturedMaterials--;
  device.setMaterial(slot, m);
}

// stack entries a hierarchy may need: OSPCUDA_TRAVERSAL_STACK (what traversal keeps), or less when OSPCUDA_STACK_LIMIT says so
// (lets a test drive the fall-back build with a hierarchy the device could still walk)
static uint32_t stackLimit()
{
  const char *e = getenv("OSPCUDA_STACK_LIMIT");
  const uint32_t v = e ? (uint32_t)atoi(e) : 0u;
  return v > 0 && v < OSPCUDA_TRAVERSAL_STACK ? v : (uint32_t)OSPCUDA_TRAVERSAL_STACK;
}

static std::atomic<uint64_t> g_commitEpoch{0};
uint64_t commitEpoch() { return g_commitEpoch.load(); }
uint64_t nextCommitEpoch() { return g_commitEpoch.fetch_add(1) + 1; }

// ---- GeometricModel (geometry/GeometricModel.cpp:21-87) -------------------------------------------
GeometricModel::GeometricModel(Geometry *g) : geomAPI(g) { managedObjectType = OSP_GEOMETRIC_MODEL; }

void GeometricModel::commit()
{
  geom = getParamObject<Geometry>("geometry", geomAPI.get());
  if (!geom) throw std::runtime_error(toString() + " received NULL 'geometry'");
  sh = OcModel{};
  sh.geom = geom->devicePtr();
  materialData = getParamData(*this, "material", OSP_MATERIAL, false, true, &promotedMat);
  sh.materialMode = OC_MAT_NONE;
  if (materialData) {
    std::vector<uint32_t> slots(materialData->size());
    for (size_t i = 0; i < slots.size(); i++) {
      Material *m = dynamic_cast<Material *>(materialData->objectAt(i));
      if (!m) throw std::runtime_error(toString() + ": 'material' array holds a non-material object");
      slots[i] = m->slot;
    }
    matList.upload(slots.data(), slots.size() * sizeof(uint32_t));
    hostMaterials = slots;
    sh.material = matList.as<uint32_t>();
    sh.numMaterials = (uint32_t)slots.size();
    sh.materialMode = OC_MAT_FROM_MODEL;
  } else {
    materialData = getParamData(*this, "material", OSP_UINT, false, true, &promotedMat);
    if (materialData) {
      std::vector<uint32_t> ids(materialData->size());
      for (size_t i = 0; i < ids.size(); i++) ids[i] = materialData->at<uint32_t>(i);
      matList.upload(ids.data(), ids.size() * sizeof(uint32_t));
      hostMaterials = ids;
      sh.material = matList.as<uint32_t>();
      sh.numMaterials = (uint32_t)ids.size();
      sh.materialMode = OC_MAT_FROM_RENDERER;
    }
  }
  colorData = getParamData(*this, "color", OSP_VEC4F, false, true, &promotedColor);
  if (colorData) sh.color = colorData->deviceView1D();
  indexData = getParamData(*this, "index", OSP_UCHAR);
  if (indexData) sh.index = indexData->deviceView1D();
  sh.invertedNormals = getParam<bool>("invertNormals", false);
  sh.userID = getParam<uint32_t>("id", 0xFFFFFFFFu);
  epoch = nextCommitEpoch();
}

int GeometricModel::materialSlotOf(uint32_t primID, const std::vector<uint32_t> &rendererSlots) const
{
  if (sh.materialMode == OC_MAT_NONE || hostMaterials.empty()) return -1;
  uint32_t idx = primID;
  if (indexData) idx = indexData->at<uint8_t>(std::min<size_t>(idx, indexData->size() - 1));
  uint32_t mat = hostMaterials[std::min<size_t>(idx, hostMaterials.size() - 1)];
  if (sh.materialMode == OC_MAT_FROM_RENDERER) {
    if (rendererSlots.empty()) return -1;
    mat = rendererSlots[std::min<size_t>(mat, rendererSlots.size() - 1)];
  }
  return (int)mat;
}

void Geometry::getAreas(const std::vector<int32_t> &primIDs, const affine3f &xfm, std::vector<float> &areas) const
{
  areas.resize(primIDs.size());
  if (type == "sphere") {
    const vec3f sc = make_vec3f(length(xfm.l.vx), length(xfm.l.vy), length(xfm.l.vz));
    const float s3 = (sc.x + sc.y + sc.z) / 3.f;
    const float scale = 4.0f * (float)M_PI * (s3 * s3);
    const float radius = radiusData ? 0.f : sphereRadius;
    // (the reference indexes the spheres by position in the list, not by primID: Spheres.ispc:99-102)
    for (size_t i = 0; i < primIDs.size(); i++) {
      const float r = radiusData ? radiusData->at<float>(i) : radius;
      areas[i] = scale * (r * r);
    }
    return;
  }
  const bool isTri = indexData->type == OSP_VEC3UI;
  for (size_t i = 0; i < primIDs.size(); i++) {
    const uint32_t *ix = (const uint32_t *)indexData->data((size_t)primIDs[i]);
    const uint32_t i0 = ix[0], i1 = ix[1], i2 = ix[2], i3 = isTri ? ix[2] : ix[3];
    const bool quad = !isTri && i2 != i3;
    const vec3f a = vertexData->at<vec3f>(i0), b = vertexData->at<vec3f>(i1), d = vertexData->at<vec3f>(i3);
    const vec3f e1 = xfmVector(xfm, a - d), e2 = xfmVector(xfm, b - d);
    float area = length(cross(e1, e2));
    if (quad) {
      const vec3f c = vertexData->at<vec3f>(i2);
      const vec3f e3 = xfmVector(xfm, c - d);
      area += length(cross(e2, e3));
    }
    areas[i] = area * 0.5f;
  }
}

// ---- TransferFunction (LinearTransferFunction.cpp:22-155) -------------------------------------------
void TransferFunction::commit()
{
  managedObjectType = OSP_TRANSFER_FUNCTION;
  valueRange = getParam<box1f>("value", box1f{0.f, 1.f});
  colorData = getParamData(*this, "color", OSP_VEC3F, true, true, &promC);
  opacityData = getParamData(*this, "opacity", OSP_FLOAT, true, true, &promO);
  sh.color = (const OcVec3f *)colorData->deviceCompact();
  sh.opacity = (const float *)opacityData->deviceCompact();
  sh.numColors = (uint32_t)colorData->size();
  sh.numOpacities = (uint32_t)opacityData->size();
  sh.valueLower = valueRange.lower;
  sh.valueUpper = valueRange.upper;
  opacities.resize(opacityData->size());
  for (size_t i = 0; i < opacities.size(); i++) opacities[i] = opacityData->at<float>(i);
  epoch = nextCommitEpoch();
}

std::vector<box1f> TransferFunction::positiveOpacityValueRanges() const
{
  // index ranges with positive opacity, widened by one entry on each side (linear interpolation)
  std::vector<std::pair<int, int>> idx;
  bool active = false;
  int lo = 0;
  const int n = (int)opacities.size();
  for (int i = 0; i < n; i++) {
    if (opacities[i] > 0.f && !active) { active = true; lo = i; }
    else if (opacities[i] <= 0.f && active) { active = false; idx.emplace_back(lo, i); }
  }
  if (n && opacities[n - 1] > 0.f && active) idx.emplace_back(lo, n);
  std::vector<box1f> out;
  const float size = valueRange.upper - valueRange.lower;
  for (auto &r : idx) {
    box1f v{-INFINITY, INFINITY};
    const int minIdx = r.first - 1, maxIdx = r.second;
    if (minIdx >= 0) v.lower = valueRange.lower + minIdx * size / (n - 1.f);
    if (maxIdx < n) v.upper = valueRange.lower + maxIdx * size / (n - 1.f);
    out.push_back(v);
  }
  return out;
}

// ---- Volume (volume/Volume.cpp:60-222 -> Open VKL structuredRegular) ------------------------------
Volume::Volume(const std::string &t) : type(t) { managedObjectType = OSP_VOLUME; }

void Volume::commit()
{
  if (type != "structuredRegular" && type != "structured_regular")
    throw std::runtime_error("cuda device: volume type '" + type + "' is not supported (structuredRegular only)");
  gridOrigin = getParam<vec3f>("gridOrigin", make_vec3f(0.f));
  gridSpacing = getParam<vec3f>("gridSpacing", make_vec3f(1.f));
  // what the sampler implements: vertex-centred data, trilinear reconstruction, NaN outside (volume/Volume.cpp:120-180 ->
  // Open VKL structuredRegular); anything else is said, not silently rendered differently
  if (getParam<bool>("cellCentered", false))
    postStatusMsg(OSP_LOG_WARNING, "cuda device: structuredRegular 'cellCentered' is not supported, the data is taken as vertex-centred");
  if (getParam<uint32_t>("filter", OSP_VOLUME_FILTER_LINEAR) != OSP_VOLUME_FILTER_LINEAR)
    postStatusMsg(OSP_LOG_WARNING, "cuda device: volume 'filter' other than OSP_VOLUME_FILTER_LINEAR is not supported, trilinear is used");
  if (getParam<uint32_t>("gradientFilter", OSP_VOLUME_FILTER_LINEAR) != OSP_VOLUME_FILTER_LINEAR)
    postStatusMsg(OSP_LOG_WARNING, "cuda device: volume 'gradientFilter' other than OSP_VOLUME_FILTER_LINEAR is not supported");
  if (!std::isnan(getParam<float>("background", NAN)))
    postStatusMsg(OSP_LOG_WARNING, "cuda device: volume 'background' is not supported, samples outside the grid are NaN");
  Data *d = getParamObject<Data>("data");
  if (!d) throw std::runtime_error("structuredRegular volume needs a 3-D 'data' array");
  data = d;
  dims[0] = (uint32_t)d->numItems.x; dims[1] = (uint32_t)d->numItems.y; dims[2] = (uint32_t)d->numItems.z;
  if (dims[0] < 2 || dims[1] < 2 || dims[2] < 2) // the trilinear sampler reads the +1 neighbours of every cell
    throw std::runtime_error("structuredRegular volume: every dimension of 'data' must be at least 2");
  const size_t n = d->size();
  if (d->type == OSP_FLOAT && d->compact()) {
    voxels.upload(d->addr, n * sizeof(float));
  } else {
    std::vector<float> tmp(n);
    size_t o = 0;
    for (size_t z = 0; z < d->numItems.z; z++)
      for (size_t y = 0; y < d->numItems.y; y++)
        for (size_t x = 0; x < d->numItems.x; x++, o++) {
          const char *p = d->data(vec3ul{x, y, z});
          switch (d->type) {
          case OSP_UCHAR: tmp[o] = *(const uint8_t *)p; break;
          case OSP_SHORT: tmp[o] = *(const int16_t *)p; break;
          case OSP_USHORT: tmp[o] = *(const uint16_t *)p; break;
          case OSP_FLOAT: tmp[o] = *(const float *)p; break;
          case OSP_DOUBLE: tmp[o] = (float)*(const double *)p; break;
          default: throw std::runtime_error("structuredRegular volume: unsupported voxel type");
          }
        }
    voxels.upload(tmp.data(), n * sizeof(float));
  }
  for (int a = 0; a < 3; a++) mcDims[a] = std::max(1u, (dims[a] - 1 + 15) / 16);
  mcRange.reserve(sizeof(float) * 2 * (size_t)mcDims[0] * mcDims[1] * mcDims[2]);
  checkCuda(ospcuda_volume_macrocells(voxels.as<float>(), dims, mcRange.as<float>(), mcDims), "macrocell build");
  bounds.lower = gridOrigin;
  bounds.upper = gridOrigin + make_vec3f((dims[0] - 1) * gridSpacing.x, (dims[1] - 1) * gridSpacing.y, (dims[2] - 1) * gridSpacing.z);
}

VolumetricModel::VolumetricModel(Volume *v) : volumeAPI(v) { managedObjectType = OSP_VOLUMETRIC_MODEL; }

void VolumetricModel::commit()
{
  volume = getParamObject<Volume>("volume", volumeAPI.get());
  if (!volume) throw std::runtime_error(toString() + " received NULL 'volume'");
  tf = getParamObject<TransferFunction>("transferFunction");
  if (!tf) throw std::runtime_error(toString() + " received NULL 'transferFunction'");
  sh = OcVolume{};
  sh.voxels = volume->voxels.as<float>();
  sh.mcRange = volume->mcRange.as<float>();
  for (int a = 0; a < 3; a++) { sh.dims[a] = volume->dims[a]; sh.mcDims[a] = volume->mcDims[a]; }
  sh.gridOrigin = toOc(volume->gridOrigin);
  sh.gridSpacing = toOc(volume->gridSpacing);
  const box3f b = volume->getBounds();
  sh.bounds[0] = b.lower.x; sh.bounds[1] = b.lower.y; sh.bounds[2] = b.lower.z;
  sh.bounds[3] = b.upper.x; sh.bounds[4] = b.upper.y; sh.bounds[5] = b.upper.z;
  sh.densityScale = getParam<float>("densityScale", 1.f);
  sh.anisotropy = getParam<float>("anisotropy", 0.f);
  sh.gradientShadingScale = getParam<float>("gradientShadingScale", 0.f);
  sh.userID = getParam<uint32_t>("id", 0xFFFFFFFFu);
  epoch = 0;
  refreshFromChildren();
}

// the part of the struct that mirrors the transfer function (the reference holds a pointer to it, VolumetricModel.cpp:72)
bool VolumetricModel::refreshFromChildren()
{
  if (!tf || (epoch != 0 && tf->epoch <= epoch)) return false;
  sh.tf = tf->sh;
  const std::vector<box1f> ranges = tf->positiveOpacityValueRanges();
  sh.numRanges = (uint32_t)std::min<size_t>(ranges.size(), 8);
  if (ranges.size() > 8) { // too many disjoint ranges: fall back to "everything"
    sh.numRanges = 1; sh.ranges[0] = -INFINITY; sh.ranges[1] = INFINITY;
  } else
    for (uint32_t i = 0; i < sh.numRanges; i++) { sh.ranges[2 * i] = ranges[i].lower; sh.ranges[2 * i + 1] = ranges[i].upper; }
  epoch = nextCommitEpoch();
  return true;
}

// ---- Light (lights/Light.cpp:20-25, AmbientLight.cpp, DirectionalLight.cpp, QuadLight.cpp:68-108) ---
Light::Light(const std::string &t) : type(t) { managedObjectType = OSP_LIGHT; }

void Light::commit()
{
  sh = OcLight{};
  sh.isVisible = getParam<bool>("visible", true);
  const vec3f coloredIntensity = getParam<vec3f>("color", make_vec3f(1.f)) * getParam<float>("intensity", 1.f);
  ambientRadiance = make_vec3f(0.f);
  if (type == "ambient") {
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_RADIANCE);
    vec3f radiance = make_vec3f(0.f);
    if (iq == OSP_INTENSITY_QUANTITY_IRRADIANCE) radiance = coloredIntensity / (float)M_PI;
    else if (iq == OSP_INTENSITY_QUANTITY_RADIANCE) radiance = coloredIntensity;
    else postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
    sh.type = OC_LIGHT_AMBIENT;
    sh.radiance = toOc(radiance);
    ambientRadiance = radiance;
  } else if (type == "distant") {
    vec3f direction = getParam<vec3f>("direction", make_vec3f(0.f, 0.f, 1.f));
    float angularDiameter = getParam<float>("angularDiameter", .0f);
    direction = -normalize(direction);
    angularDiameter = std::min(std::max(angularDiameter, 0.f), 180.f);
    const float cosAngle = std::cos((0.5f * angularDiameter) * (float)(M_PI / 180.0));
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_IRRADIANCE);
    vec3f irradiance = make_vec3f(0.f);
    if (iq == OSP_INTENSITY_QUANTITY_RADIANCE) irradiance = (2.0f * (float)M_PI * (1.0f - cosAngle)) * coloredIntensity;
    else if (iq == OSP_INTENSITY_QUANTITY_IRRADIANCE) irradiance = coloredIntensity;
    else postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
    const linear3f f = frame(direction);
    sh.type = OC_LIGHT_DISTANT;
    sh.radiance = toOc(irradiance);
    sh.fx = toOc(f.vx); sh.fy = toOc(f.vy); sh.fz = toOc(f.vz);
    sh.cosAngle = cosAngle;
    sh.pdf = cosAngle < 0.99999988f ? 1.f / (2.f * (float)M_PI * (1.0f - cosAngle)) : INFINITY;
  } else if (type == "quad") {
    const vec3f position = getParam<vec3f>("position", make_vec3f(0.f));
    const vec3f edge1 = getParam<vec3f>("edge1", make_vec3f(1.f, 0.f, 0.f));
    const vec3f edge2 = getParam<vec3f>("edge2", make_vec3f(0.f, 1.f, 0.f));
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_RADIANCE);
    const float quadArea = length(cross(edge1, edge2));
    vec3f radiance = make_vec3f(0.f);
    if (iq == OSP_INTENSITY_QUANTITY_INTENSITY) radiance = coloredIntensity / quadArea;
    else if (iq == OSP_INTENSITY_QUANTITY_POWER) radiance = coloredIntensity / ((float)M_PI * quadArea);
    else if (iq == OSP_INTENSITY_QUANTITY_RADIANCE) radiance = coloredIntensity;
    else postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
    const vec3f ndirection = cross(edge2, edge1);
    const float ppdf = 1.f / length(ndirection);
    sh.type = OC_LIGHT_QUAD;
    sh.radiance = toOc(radiance);
    sh.position = toOc(position); sh.edge1 = toOc(edge1); sh.edge2 = toOc(edge2);
    sh.ppdf = ppdf;
    sh.nnormal = toOc(ndirection * ppdf);
  } else if (type == "sphere") { // lights/PointLight.cpp:77-124 (no intensity distribution)
    const vec3f position = getParam<vec3f>("position", make_vec3f(0.f));
    const float radius = getParam<float>("radius", 0.f);
    getParam<vec3f>("direction", make_vec3f(0.f, 0.f, 1.f));
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_INTENSITY);
    const float sphereArea = 4.0f * M_PI * radius * radius;
    vec3f radIntensity = make_vec3f(0.f), radiance = make_vec3f(0.f);
    if (iq == OSP_INTENSITY_QUANTITY_INTENSITY) {
      radIntensity = coloredIntensity;
      if (radius > 0.0f) radiance = coloredIntensity / (sphereArea / 4.0f);
    } else if (iq == OSP_INTENSITY_QUANTITY_POWER) {
      radIntensity = coloredIntensity / (float)(4.0f * M_PI);
      if (radius > 0.0f) radiance = coloredIntensity / (float)(M_PI * sphereArea);
    } else if (iq == OSP_INTENSITY_QUANTITY_RADIANCE) {
      if (radius > 0.0f) radiance = coloredIntensity;
    } else
      postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
    sh.type = OC_LIGHT_SPHERE;
    sh.position = toOc(position);
    sh.radius = radius;
    sh.intensity = toOc(radIntensity);
    sh.radiance = toOc(radiance);
  } else if (type == "spot") { // lights/SpotLight.cpp:22-128 (no intensity distribution)
    const vec3f position = getParam<vec3f>("position", make_vec3f(0.f));
    const vec3f direction = getParam<vec3f>("direction", make_vec3f(0.f, 0.f, 1.f));
    float openingAngle = getParam<float>("openingAngle", 180.f);
    float penumbraAngle = getParam<float>("penumbraAngle", 5.f);
    const float radius = std::max(0.0f, getParam<float>("radius", 0.f));
    const float innerRadius = std::min(std::max(getParam<float>("innerRadius", 0.f), 0.0f), 0.999f * radius);
    const vec3f c0i = std::abs(direction.x) < std::abs(direction.y) ? make_vec3f(0.0f, direction.z, direction.y)
                                                                    : make_vec3f(direction.z, 0.0f, direction.x);
    const vec3f d = normalize(direction);
    const vec3f c90 = normalize(cross(c0i, d));
    const vec3f c0 = cross(d, c90);
    openingAngle = std::min(std::max(openingAngle, 0.f), 360.f);
    penumbraAngle = std::min(std::max(penumbraAngle, 0.f), 0.5f * openingAngle);
    const float deg = 1.74532925199432957692e-2f;
    const float cosAngleMax = std::cos((0.5f * openingAngle) * deg);
    const float cosAngleMin = std::cos((0.5f * openingAngle - penumbraAngle) * deg);
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_INTENSITY);
    const float halfOpeningAngleRad = M_PI * (openingAngle * 0.5f) / 180.0f;
    const float cosHalfOpeningAngle = cos(halfOpeningAngleRad);
    const float ringDiskArea = M_PI * (radius * radius - innerRadius * innerRadius);
    const float sphericalCapCosInt = M_PI * (1.0f - cosHalfOpeningAngle * cosHalfOpeningAngle);
    vec3f radIntensity = make_vec3f(0.f), radiance = make_vec3f(0.f);
    if (iq == OSP_INTENSITY_QUANTITY_INTENSITY) {
      radIntensity = coloredIntensity;
      if (radius > 0.0f) radiance = radIntensity / ringDiskArea;
    } else if (iq == OSP_INTENSITY_QUANTITY_POWER) {
      radIntensity = coloredIntensity / sphericalCapCosInt;
      if (radius > 0.0f) radiance = coloredIntensity / (sphericalCapCosInt * ringDiskArea);
    } else if (iq == OSP_INTENSITY_QUANTITY_RADIANCE) {
      if (radius > 0.0f) radiance = coloredIntensity;
    } else
      postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
    sh.type = OC_LIGHT_SPOT;
    sh.position = toOc(position);
    sh.fx = toOc(c0); sh.fy = toOc(c90); sh.fz = toOc(d);
    sh.cosAngle = cosAngleMax;
    sh.cosAngleScale = 1.0f / (cosAngleMin - cosAngleMax);
    sh.radius = radius; sh.innerRadius = innerRadius;
    sh.ppdf = 1.f / ((float)M_PI * (radius * radius - innerRadius * innerRadius));
    sh.intensity = toOc(radIntensity);
    sh.radiance = toOc(radiance);
  } else if (type == "cylinder") {
    // lights/CylinderLight.cpp:22-97
    const vec3f position0 = getParam<vec3f>("position0", make_vec3f(0.f));
    const vec3f position1 = getParam<vec3f>("position1", make_vec3f(0.f, 0.f, 1.f));
    const float radius = getParam<float>("radius", 1.f);
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_RADIANCE);
    const float cylinderArea = 2.f * (float)M_PI * length(position1 - position0) * radius;
    vec3f radiance = make_vec3f(0.f);
    if (iq == OSP_INTENSITY_QUANTITY_POWER) { if (cylinderArea > 0.f) radiance = coloredIntensity / ((float)M_PI * cylinderArea); }
    else if (iq == OSP_INTENSITY_QUANTITY_INTENSITY) { if (cylinderArea > 0.f) radiance = (float)M_PI * coloredIntensity / cylinderArea; }
    else if (iq == OSP_INTENSITY_QUANTITY_RADIANCE) radiance = coloredIntensity;
    else postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
    if (length(position1 - position0) <= 0.f || radius <= 0.f) radiance = make_vec3f(0.f);
    sh.type = OC_LIGHT_CYLINDER;
    sh.position = toOc(position0); sh.edge1 = toOc(position1); sh.radius = radius;
    sh.radiance = toOc(radiance);
  } else if (type == "hdri") {
    // lights/HDRILight.cpp:73-110: frame from direction / up, intensityQuantity must be SCALE, the importance distribution
    // of the map (HDRILight_initDistribution) rebuilt at every commit
    const vec3f up = getParam<vec3f>("up", make_vec3f(0.f, 1.f, 0.f));
    const vec3f dir = getParam<vec3f>("direction", make_vec3f(0.f, 0.f, 1.f));
    hdriMap = dynamic_cast<Texture2D *>(getParamObjectRaw("map"));
    if (!hdriMap) throw std::runtime_error(toString() + " must have 'map'");
    const uint32_t iq = getParam<uint32_t>("intensityQuantity", OSP_INTENSITY_QUANTITY_SCALE);
    vec3f scale = coloredIntensity;
    if (iq != OSP_INTENSITY_QUANTITY_SCALE) {
      postStatusMsg(OSP_LOG_WARNING, toString() + " unsupported 'intensityQuantity' value");
      scale = make_vec3f(0.f);
    }
    linear3f f;
    f.vx = normalize(-dir);
    f.vy = normalize(cross(f.vx, up));
    f.vz = cross(f.vx, f.vy);
    const linear3f inv = rcp(f);
    const int sx = hdriMap->sh.sizeX, sy = hdriMap->sh.sizeY;
    hdriCdfX.reserve((size_t)sx * sy * sizeof(float));
    hdriCdfY.reserve((size_t)sy * sizeof(float));
    checkCuda(ospcuda_hdri_distribution(hdriMap->devPtr(), sx, sy, hdriCdfX.as<float>(), hdriCdfY.as<float>()), "hdri distribution");
    sh.type = OC_LIGHT_HDRI;
    sh.radiance = toOc(scale);
    sh.fx = toOc(f.vx); sh.fy = toOc(f.vy); sh.fz = toOc(f.vz);
    sh.wx = toOc(inv.vx); sh.wy = toOc(inv.vy); sh.wz = toOc(inv.vz);
    sh.hdriMap = hdriMap->devPtr();
    sh.hdriCdfX = hdriCdfX.as<float>(); sh.hdriCdfY = hdriCdfY.as<float>();
  } else
    throw std::runtime_error("cuda device: light type '" + type + "' is not supported (ambient, cylinder, distant, quad, sphere, spot, hdri)");
}

// ---- Group (common/Group.cpp:69-209) ------------------------------------------------------------------
Group::~Group() { ospcuda_bvh_free(&sh.bvh); }

void Group::commit()
{
  managedObjectType = OSP_GROUP;
  stackNeed = 0;
  geometricModels = getParamData(*this, "geometry", OSP_GEOMETRIC_MODEL);
  volumetricModels = getParamData(*this, "volume", OSP_VOLUMETRIC_MODEL);
  lights = getParamData(*this, "light", OSP_LIGHT);
  if (getParamData(*this, "clippingGeometry", OSP_GEOMETRIC_MODEL))
    postStatusMsg(OSP_LOG_WARNING, "cuda device: clipping geometry is not supported and is ignored");
  getParam<bool>("dynamicScene", false);
  getParam<bool>("compactMode", false);
  getParam<bool>("robustMode", false);

  ospcuda_bvh_free(&sh.bvh);
  sh = OcGroup{};
  bounds = box_empty();
  lastBuildMs = 0.f;
  if (geometricModels) {
    const size_t n = geometricModels->size();
    std::vector<OcModel> models(n);
    std::vector<uint32_t> numPrims(n), types(n);
    for (size_t i = 0; i < n; i++) {
      GeometricModel *m = dynamic_cast<GeometricModel *>(geometricModels->objectAt(i));
      if (!m || !m->geom) throw std::runtime_error("ospray::Group: 'geometry' holds an uncommitted or foreign object");
      models[i] = m->sh;
      numPrims[i] = m->geom->sh.numPrims;
      types[i] = m->geom->sh.type;
    }
    modelsDev.upload(models.data(), n * sizeof(OcModel));
    sh.models = modelsDev.as<OcModel>();
    sh.numModels = (uint32_t)n;
    checkCuda(ospcuda_bvh_build_group(sh.models, sh.numModels, numPrims.data(), types.data(), &sh.bvh, &lastBuildMs), "BVH build");
    stackNeed = ospcuda_bvh_last_stack_need();
    if (stackNeed > stackLimit()) {
      // strongly clustered primitives make a long chain of the Morton hierarchy: build this group again over a coarser
      // grid with larger leaves (equal keys are split by index, a balanced subtree) instead of refusing the scene
      postStatusMsg(OSP_LOG_INFO, "cuda device: hierarchy of a group needs " + std::to_string(stackNeed) + " stack entries, rebuilding it over a coarser grid");
      ospcuda_bvh_free(&sh.bvh);
      ospcuda_bvh_set_build_options(10, 8);
      const int rc = ospcuda_bvh_build_group(sh.models, sh.numModels, numPrims.data(), types.data(), &sh.bvh, &lastBuildMs);
      ospcuda_bvh_set_build_options(0, 0);
      checkCuda(rc, "BVH build");
      stackNeed = ospcuda_bvh_last_stack_need();
    }
    if (stackNeed > stackLimit())
      throw std::runtime_error("ospray::Group: the hierarchy over this geometry is too deep for the traversal stack (needs "
          + std::to_string(stackNeed) + " entries, the device keeps " + std::to_string(OSPCUDA_TRAVERSAL_STACK) + ")");
    if (sh.bvh.numRecords) {
      box3f b;
      b.lower = make_vec3f(sh.bvh.bounds[0], sh.bvh.bounds[1], sh.bvh.bounds[2]);
      b.upper = make_vec3f(sh.bvh.bounds[3], sh.bvh.bounds[4], sh.bvh.bounds[5]);
      box_extend(bounds, b);
    }
  }
  if (volumetricModels) {
    const size_t n = volumetricModels->size();
    std::vector<OcVolume> vols(n);
    for (size_t i = 0; i < n; i++) {
      VolumetricModel *m = dynamic_cast<VolumetricModel *>(volumetricModels->objectAt(i));
      if (!m || !m->volume) throw std::runtime_error("ospray::Group: 'volume' holds an uncommitted or foreign object");
      vols[i] = m->sh;
      box_extend(bounds, m->getBounds());
    }
    volumesDev.upload(vols.data(), n * sizeof(OcVolume));
    sh.volumes = volumesDev.as<OcVolume>();
    sh.numVolumes = (uint32_t)n;
  }
  dev.upload(&sh, sizeof(sh));
  snapshotEpoch = commitEpoch();
}

void Group::refreshSnapshots()
{
  if (snapshotEpoch == commitEpoch()) return;
  bool models = false, vols = false;
  if (geometricModels)
    for (size_t i = 0; i < geometricModels->size() && !models; i++) {
      GeometricModel *m = dynamic_cast<GeometricModel *>(geometricModels->objectAt(i));
      models = m && m->epoch > snapshotEpoch;
    }
  if (volumetricModels)
    for (size_t i = 0; i < volumetricModels->size(); i++) {
      VolumetricModel *m = dynamic_cast<VolumetricModel *>(volumetricModels->objectAt(i));
      if (m && (m->refreshFromChildren() || m->epoch > snapshotEpoch)) vols = true;
    }
  if (models) {
    std::vector<OcModel> snap(geometricModels->size());
    for (size_t i = 0; i < snap.size(); i++) snap[i] = dynamic_cast<GeometricModel *>(geometricModels->objectAt(i))->sh;
    modelsDev.upload(snap.data(), snap.size() * sizeof(OcModel));
    sh.models = modelsDev.as<OcModel>();
  }
  if (vols) {
    std::vector<OcVolume> snap(volumetricModels->size());
    for (size_t i = 0; i < snap.size(); i++) snap[i] = dynamic_cast<VolumetricModel *>(volumetricModels->objectAt(i))->sh;
    volumesDev.upload(snap.data(), snap.size() * sizeof(OcVolume));
    sh.volumes = volumesDev.as<OcVolume>();
  }
  if (models || vols) dev.upload(&sh, sizeof(sh)); // (the array may have moved)
  snapshotEpoch = commitEpoch();
}

// ---- Instance (common/Instance.cpp:20-52) ----------------------------------------------------------------
Instance::Instance(Group *g) : groupAPI(g) { managedObjectType = OSP_INSTANCE; }

void Instance::commit()
{
  group = getParamObject<Group>("group", groupAPI.get());
  if (!group) throw std::runtime_error(toString() + " received NULL 'group'");
  xfm = getParam<affine3f>("transform", getParam<affine3f>("xfm", affine_identity()));
  if (getParamObjectRaw("motion.transform"))
    postStatusMsg(OSP_LOG_WARNING, "cuda device: motion blur transforms are not supported");
  rcp_xfm = rcp(xfm);
  userID = getParam<uint32_t>("id", 0xFFFFFFFFu);
  const affine3f I = affine_identity();
  identity = std::memcmp(&xfm, &I, sizeof(affine3f)) == 0;
}

box3f Instance::getBounds() const
{
  if (!group) return box_empty();
  const box3f g = group->getBounds();
  if (box_is_empty(g)) return g;
  return xfmBounds(xfm, g);
}

// ---- World (common/World.cpp:70-209; PathTracerData.cpp:32-82; SciVisData.cpp:74-105) -----------------------
World::~World() { ospcuda_bvh_free(&sh.tlas); }

static float distribution1D_create(std::vector<float> &cdf) // math/Distribution1D.ispc:14-41
{
  float sum = cdf[0];
  for (size_t i = 1; i < cdf.size(); i++) cdf[i] = sum += cdf[i];
  const float rcpSum = 1.0f / sum;
  for (float &c : cdf) c = c >= sum ? 0x1.000002p+0f : c * rcpSum;
  return sum;
}

void World::commit()
{
  instances = getParamData(*this, "instance", OSP_INSTANCE);
  lights = getParamData(*this, "light", OSP_LIGHT);
  getParam<bool>("dynamicScene", false);
  getParam<bool>("compactMode", false);
  getParam<bool>("robustMode", false);

  ospcuda_bvh_free(&sh.tlas);
  sh = OcWorld{};
  bounds = box_empty();
  std::vector<OcInstance> inst;
  std::vector<float> instBounds;
  if (instances) {
    inst.resize(instances->size());
    for (size_t i = 0; i < inst.size(); i++) {
      Instance *in = dynamic_cast<Instance *>(instances->objectAt(i));
      if (!in || !in->group) throw std::runtime_error("ospray::World: 'instance' holds an uncommitted or foreign object");
      OcInstance &o = inst[i];
      o = OcInstance{};
      o.xfm = toOc(in->xfm);
      o.rcp_xfm = toOc(in->rcp_xfm);
      o.group = in->group->devicePtr();
      o.userID = in->userID;
      o.isIdentity = in->identity;
      const box3f b = in->getBounds();
      o.bounds[0] = b.lower.x; o.bounds[1] = b.lower.y; o.bounds[2] = b.lower.z;
      o.bounds[3] = b.upper.x; o.bounds[4] = b.upper.y; o.bounds[5] = b.upper.z;
      instBounds.insert(instBounds.end(), o.bounds, o.bounds + 6);
      if (!box_is_empty(b)) box_extend(bounds, b);
      if (in->group->hasVolumes()) sh.hasVolumes = 1;
    }
    instDev.upload(inst.data(), inst.size() * sizeof(OcInstance));
    hostInstances = inst;
    sh.instances = instDev.as<OcInstance>();
    sh.numInstances = (uint32_t)inst.size();
    if (!inst.empty()) {
      // the top level is always a hierarchy (one instance per leaf), also over a single instance: traversal is one state
      // machine that switches between it and the group hierarchies on one stack (intersect.cuh Traversal)
      checkCuda(ospcuda_bvh_build_tlas(instBounds.data(), (uint32_t)inst.size(), &sh.tlas), "TLAS build");
      const uint32_t tlasNeed = ospcuda_bvh_last_stack_need();
      uint32_t groupNeed = 0;
      for (size_t i = 0; i < inst.size(); i++) {
        Instance *in = dynamic_cast<Instance *>(instances->objectAt(i));
        groupNeed = std::max(groupNeed, in->group->stackNeed);
      }
      if (tlasNeed + groupNeed > OSPCUDA_TRAVERSAL_STACK)
        throw std::runtime_error("ospray::World: instance hierarchy + group hierarchy are too deep for the traversal stack (need "
            + std::to_string(tlasNeed) + " + " + std::to_string(groupNeed) + " entries, the device keeps "
            + std::to_string(OSPCUDA_TRAVERSAL_STACK) + ")");
    }
  }
  // light lists: world lights, then the lights of the instances' groups, moved into world space with the instance
  // (World.cpp:120-150 + *Light::createSh(instance): QuadLight.ispc:14-33, PointLight.ispc:17-29, SpotLight.ispc:14-24,
  // DirectionalLight.cpp:36-44)
  auto inWorldSpace = [](OcLight l, const affine3f &xfm) {
    auto V = [](const OcVec3f &v) { return make_vec3f(v.x, v.y, v.z); };
    switch (l.type) {
    case OC_LIGHT_QUAD: {
      const vec3f position = xfmPoint(xfm, V(l.position)), e1 = xfmVector(xfm.l, V(l.edge1)), e2 = xfmVector(xfm.l, V(l.edge2));
      const vec3f ndirection = cross(e2, e1);
      const float ppdf = 1.f / length(ndirection);
      l.position = toOc(position); l.edge1 = toOc(e1); l.edge2 = toOc(e2); l.ppdf = ppdf; l.nnormal = toOc(ndirection * ppdf);
      break;
    }
    case OC_LIGHT_SPHERE: l.position = toOc(xfmPoint(xfm, V(l.position))); break;
    case OC_LIGHT_CYLINDER: // CylinderLight.ispc:12-19: both end points
      l.position = toOc(xfmPoint(xfm, V(l.position))); l.edge1 = toOc(xfmPoint(xfm, V(l.edge1)));
      break;
    case OC_LIGHT_SPOT: { // fx / fy / fz = c0 / c90 / direction
      const vec3f direction = normalize(xfmVector(xfm.l, V(l.fz)));
      const vec3f c90 = normalize(cross(xfmVector(xfm.l, V(l.fx)), direction));
      l.position = toOc(xfmPoint(xfm, V(l.position)));
      l.fz = toOc(direction); l.fy = toOc(c90); l.fx = toOc(cross(direction, c90));
      break;
    }
    case OC_LIGHT_DISTANT:
      l.fx = toOc(xfmVector(xfm.l, V(l.fx))); l.fy = toOc(xfmVector(xfm.l, V(l.fy))); l.fz = toOc(xfmVector(xfm.l, V(l.fz)));
      break;
    case OC_LIGHT_HDRI: { // HDRILight.cpp:52-55: light2world = xfm.l * light2world, world2light its inverse
      linear3f f;
      f.vx = xfmVector(xfm.l, V(l.fx)); f.vy = xfmVector(xfm.l, V(l.fy)); f.vz = xfmVector(xfm.l, V(l.fz));
      const linear3f inv = rcp(f);
      l.fx = toOc(f.vx); l.fy = toOc(f.vy); l.fz = toOc(f.vz);
      l.wx = toOc(inv.vx); l.wy = toOc(inv.vy); l.wz = toOc(inv.vz);
      break;
    }
    default: break; // ambient
    }
    return l;
  };
  std::vector<std::pair<OcLight, vec3f>> all; // (light in world space, ambient radiance)
  if (lights)
    for (size_t i = 0; i < lights->size(); i++)
      if (Light *l = dynamic_cast<Light *>(lights->objectAt(i))) all.push_back({l->sh, l->ambientRadiance});
  if (instances)
    for (size_t i = 0; i < instances->size(); i++) {
      Instance *in = dynamic_cast<Instance *>(instances->objectAt(i));
      if (in && in->group && in->group->lights)
        for (size_t k = 0; k < in->group->lights->size(); k++)
          if (Light *l = dynamic_cast<Light *>(in->group->lights->objectAt(k)))
            all.push_back({in->identity ? l->sh : inWorldSpace(l->sh, in->xfm), l->ambientRadiance});
    }
  std::vector<OcLight> pt, sv;
  uint32_t visibleOnly = 0;
  vec3f aoColor = make_vec3f(0.f);
  hasHdriLight = false;
  for (const auto &la : all) {
    const OcLight &lsh = la.first;
    pt.push_back(lsh);
    if (lsh.type == OC_LIGHT_HDRI || lsh.type == OC_LIGHT_CYLINDER) hasHdriLight = true; // (lights only the extended kernels know)
    if (lsh.type == OC_LIGHT_AMBIENT || lsh.type == OC_LIGHT_HDRI) {
      // SciVisData.cpp:17-52: ambient and hdri lights go to the visible-only front partition (no illumination from them)
      if (visibleOnly == sv.size()) sv.push_back(lsh);
      else { sv.push_back(sv[visibleOnly]); sv[visibleOnly] = lsh; }
      visibleOnly++;
      aoColor = aoColor + la.second;
    } else
      sv.push_back(lsh);
  }
  if (!pt.empty()) {
    std::vector<float> cdf(pt.size(), 1.f);
    distribution1D_create(cdf);
    lightsDev.upload(pt.data(), pt.size() * sizeof(OcLight));
    cdfDev.upload(cdf.data(), cdf.size() * sizeof(float));
    scivisLightsDev.upload(sv.data(), sv.size() * sizeof(OcLight));
    sh.lights = lightsDev.as<OcLight>();
    sh.lightsCDF = cdfDev.as<float>();
    sh.scivisLights = scivisLightsDev.as<OcLight>();
  }
  hostLights = pt;
  geoKey[0] = geoKey[1] = geoKey[2] = ~0ull;
  geoActive = false;
  if (!instances) hostInstances.clear();
  sh.numLights = (uint32_t)pt.size();
  sh.numScivisLights = (uint32_t)sv.size();
  sh.numScivisLightsVisibleOnly = visibleOnly;
  sh.aoColorPi = toOc(aoColor * (float)M_PI);
  sh.bounds[0] = bounds.lower.x; sh.bounds[1] = bounds.lower.y; sh.bounds[2] = bounds.lower.z;
  sh.bounds[3] = bounds.upper.x; sh.bounds[4] = bounds.upper.y; sh.bounds[5] = bounds.upper.z;
  dirty = true;
}

// PathTracerData::generateGeometryLights + GeometricModel_gatherEmissivePrimIDs (render/pathtracer/PathTracerData.cpp:87-160,
// geometry/GeometricModel.ispc:14-56): one proxy light per (instance, model) that has emissive primitives
void World::buildGeometryLights(const Renderer &renderer)
{
  std::vector<OcLight> geo;
  std::vector<OcInstance> inst = hostInstances;
  geoBuffers.clear();
  const std::vector<OcMaterial> &table = device.materialTableHost();
  for (size_t i = 0; instances && i < instances->size(); i++) {
    Instance *in = dynamic_cast<Instance *>(instances->objectAt(i));
    if (!in || !in->group || !in->group->geometricModels) continue;
    Data *models = in->group->geometricModels.get();
    std::vector<float> areaPDF(models->size(), 0.f);
    bool any = false;
    for (size_t g = 0; g < models->size(); g++) {
      GeometricModel *model = dynamic_cast<GeometricModel *>(models->objectAt(g));
      if (!model || !model->geom) continue;
      // hasEmissiveMaterials: any material the model can reach emits
      bool emissive = false;
      for (uint32_t hm : model->hostMaterials) {
        uint32_t slot = hm;
        if (model->sh.materialMode == OC_MAT_FROM_RENDERER) {
          if (renderer.materialSlots.empty()) continue;
          slot = renderer.materialSlots[std::min<size_t>(hm, renderer.materialSlots.size() - 1)];
        }
        if (slot < table.size() && table[slot].isEmissive) emissive = true;
      }
      if (!emissive) continue;
      if (!model->geom->supportAreaLighting()) {
        postStatusMsg(OSP_LOG_WARNING, "#osp:pt Geometry " + model->toString() + " does not implement area sampling! Cannot use importance sampling for that geometry with emissive material!");
        continue;
      }
      std::vector<int32_t> primIDs;
      for (uint32_t p = 0; p < model->geom->numPrimitives(); p++) {
        const int slot = model->materialSlotOf(p, renderer.materialSlots);
        if (slot >= 0 && (size_t)slot < table.size() && table[slot].isEmissive) primIDs.push_back((int32_t)p);
      }
      if (primIDs.empty()) continue;
      std::vector<float> distribution;
      model->geom->getAreas(primIDs, in->xfm, distribution);
      const float pdf = 1.f / distribution1D_create(distribution);
      areaPDF[g] = pdf;
      any = true;
      OcLight l{};
      l.type = OC_LIGHT_GEOMETRY;
      l.isVisible = 1;
      l.geoModel = in->group->sh.models + g;
      l.geoInstance = instDevGeo() + i;
      geoBuffers.emplace_back(new DevMem); geoBuffers.back()->upload(primIDs.data(), primIDs.size() * sizeof(int32_t));
      l.geoPrimIDs = geoBuffers.back()->as<int32_t>();
      geoBuffers.emplace_back(new DevMem); geoBuffers.back()->upload(distribution.data(), distribution.size() * sizeof(float));
      l.geoDistribution = geoBuffers.back()->as<float>();
      l.geoNumPrimitives = (uint32_t)primIDs.size();
      l.geoPdf = pdf;
      geo.push_back(l);
    }
    if (any) {
      geoBuffers.emplace_back(new DevMem); geoBuffers.back()->upload(areaPDF.data(), areaPDF.size() * sizeof(float));
      inst[i].modelAreaPDF = geoBuffers.back()->as<float>();
    }
  }
  geoActive = !geo.empty();
  if (!geoActive) return;
  std::vector<OcLight> all = geo;
  all.insert(all.end(), hostLights.begin(), hostLights.end());
  std::vector<float> cdf(all.size(), 1.f);
  distribution1D_create(cdf);
  geoLightsDev.upload(all.data(), all.size() * sizeof(OcLight));
  geoCdfDev.upload(cdf.data(), cdf.size() * sizeof(float));
  geoInstDev.upload(inst.data(), inst.size() * sizeof(OcInstance));
  geoNumLights = (uint32_t)all.size();
  geoNumGeo = (uint32_t)geo.size();
}

const OcInstance *World::instDevGeo()
{
  geoInstDev.reserve(std::max<size_t>(1, hostInstances.size()) * sizeof(OcInstance));
  return geoInstDev.as<OcInstance>();
}

// OcWorld.shadowOpaque: the host-side mirror of ptOccluderIsOpaque (pathtracer_wf.cuh) over every model of the world
uint32_t World::computeShadowOpaque(const Renderer *renderer) const
{
  if (!instances) return 2u;
  const std::vector<OcMaterial> &table = device.materialTableHost();
  uint32_t level = 2u;
  for (size_t i = 0; i < instances->size(); i++) {
    Instance *in = dynamic_cast<Instance *>(instances->objectAt(i));
    if (!in || !in->group || !in->group->geometricModels) continue;
    Data *models = in->group->geometricModels.get();
    for (size_t g = 0; g < models->size(); g++) {
      GeometricModel *m = dynamic_cast<GeometricModel *>(models->objectAt(g));
      if (!m || !m->geom) return 0u;
      if (m->sh.color.addr) return 0u;
      const uint32_t gf = m->geom->sh.flags;
      if (gf & OC_GEOMFLAG_HAS_ALPHA) {
        if (!(gf & OC_GEOMFLAG_ALPHA_ALL_ONE)) return 0u;
        level = 1u;
      }
      for (uint32_t hm : m->hostMaterials) {
        uint32_t slot = hm;
        if (m->sh.materialMode == OC_MAT_FROM_RENDERER) {
          if (!renderer) return 0u;
          if (renderer->materialSlots.empty()) continue; // no material: opaque
          slot = renderer->materialSlots[std::min<size_t>(hm, renderer->materialSlots.size() - 1)];
        }
        if (slot >= table.size()) return 0u;
        const OcMaterial &mt = table[slot];
        if (mt.cutout || mt.type == OC_MATERIAL_THINGLASS) return 0u;
        if (mt.type == OC_MATERIAL_LUMINOUS && mt.transparency != 0.f) return 0u;
        if (mt.type == OC_MATERIAL_PRINCIPLED || mt.type == OC_MATERIAL_GLASS) return 0u; // (cut-out opacity, transmission: decided per occluder)
        if (mt.type == OC_MATERIAL_OBJ && !(mt.d == 1.f && mt.Tf.x == 0.f && mt.Tf.y == 0.f && mt.Tf.z == 0.f)) return 0u;
      }
    }
  }
  return level;
}

const OcWorld *World::prepareFrame(const Renderer *renderer)
{
  // models / transfer functions committed since the last frame: the groups take their new state over (Group::refreshSnapshots)
  if (seenEpoch != commitEpoch() && instances) {
    for (size_t i = 0; i < instances->size(); i++) {
      Instance *in = dynamic_cast<Instance *>(instances->objectAt(i));
      if (in && in->group) in->group->refreshSnapshots();
    }
    seenEpoch = commitEpoch();
  }
  // geometry lights: only a pathtracer that asks for them, only when some material emits; rebuilt when the scene, a
  // material or the renderer was committed since
  const bool wantGeo = renderer && renderer->sh.type == OC_RENDERER_PATHTRACER && renderer->sh.geometryLights && device.emissiveMaterials > 0;
  if (wantGeo) {
    const uint64_t key[3] = {commitEpoch(), device.materialVersion, (uint64_t)(uintptr_t)renderer ^ (renderer->commitCount << 48)};
    if (key[0] != geoKey[0] || key[1] != geoKey[1] || key[2] != geoKey[2]) {
      buildGeometryLights(*renderer);
      geoKey[0] = key[0]; geoKey[1] = key[1]; geoKey[2] = key[2];
      dirty = true;
    }
  }
  { // (every frame: a walk over the models and their material lists, no primitive data)
    const uint32_t so = computeShadowOpaque(renderer);
    if (so != sh.shadowOpaque) { sh.shadowOpaque = so; dirty = true; }
  }
  const bool useGeo = wantGeo && geoActive;
  if (useGeo != geoInUse) { geoInUse = useGeo; dirty = true; }
  if (dirty) {
    if (useGeo) {
      sh.lights = geoLightsDev.as<OcLight>(); sh.lightsCDF = geoCdfDev.as<float>();
      sh.numLights = geoNumLights; sh.numGeoLights = geoNumGeo;
      sh.instances = geoInstDev.as<OcInstance>();
    } else {
      sh.lights = hostLights.empty() ? nullptr : lightsDev.as<OcLight>(); sh.lightsCDF = hostLights.empty() ? nullptr : cdfDev.as<float>();
      sh.numLights = (uint32_t)hostLights.size(); sh.numGeoLights = 0;
      sh.instances = hostInstances.empty() ? nullptr : instDev.as<OcInstance>();
    }
  }
  const OcMaterial *mt = device.materialTable();
  if (dirty || mt != uploadedMaterials) {
    sh.materials = mt;
    dev.upload(&sh, sizeof(sh));
    uploadedMaterials = mt;
    dirty = false;
  }
  return dev.as<OcWorld>();
}

// ---- ImageOp (fb/frame_ops/ToneMapper.cpp:14-51) ------------------------------------------------------------
void ImageOp::commit()
{
  if (type != "tonemapper") return;
  tm.exposure = getParam<float>("exposure", 1.f);
  const float a = std::max(getParam<float>("contrast", 1.6773f), 0.0001f);
  const float d = std::min(std::max(getParam<float>("shoulder", 0.9714f), 0.0001f), 1.f);
  const float m = std::min(std::max(getParam<float>("midIn", 0.18f), 0.0001f), 1.f);
  const float n = std::min(std::max(getParam<float>("midOut", 0.18f), 0.0001f), 1.f);
  const float w = std::max(getParam<float>("hdrMax", 11.0785f), 1.f);
  tm.acesColor = getParam<bool>("acesColor", true);
  tm.a = a;
  tm.d = d;
  tm.b = -((powf(m, -a * d) * (-powf(m, a) + (n * (powf(m, a * d) * n * powf(w, a) - powf(m, a) * powf(w, a * d))) / (powf(m, a * d) * n - n * powf(w, a * d)))) / n);
  tm.c = std::max((powf(m, a * d) * n * powf(w, a) - powf(m, a) * powf(w, a * d)) / (powf(m, a * d) * n - n * powf(w, a * d)), 0.f);
}

// ---- Renderer (render/Renderer.cpp:38-92,146-178; PathTracer.cpp:29-51; SciVis.cpp:27-36; AORenderer.cpp:26-35)
Renderer::Renderer(CudaDevice &d, const std::string &t) : device(d), type(t) { managedObjectType = OSP_RENDERER; }

static float filterEval(uint32_t pf, float width, float x, float y) // pf/LUTPixelFilter.ispc:17-66
{
  if (pf == OSP_PIXELFILTER_GAUSS) {
    const float sigma = (0.5f * width) / 3.0f;
    return std::exp(-0.5f * ((x * x + y * y) / (sigma * sigma)));
  }
  if (pf == OSP_PIXELFILTER_MITCHELL) {
    const float B = 1.0f / 3.0f, C = 1.0f / 3.0f, oneOverSix = 1.0f / 6.0f;
    const float tx = (x / width) * 4.0f, ty = (y / width) * 4.0f;
    const float fx = std::sqrt(tx * tx + ty * ty), fx2 = fx * fx, fx3 = fx2 * fx;
    if (fx < 1.0f) return oneOverSix * ((12.0f - 9.0f * B - 6.0f * C) * fx3 + (-18.0f + 12.0f * B + 6.0f * C) * fx2 + (6.0f - 2.0f * B));
    if (fx < 2.0f) return oneOverSix * ((-B - 6.0f * C) * fx3 + (6.0f * B + 30.0f * C) * fx2 + (-12.0f * B - 48.0f * C) * fx + (8.0f * B + 24.0f * C));
    return 0.f;
  }
  // Blackman-Harris
  const float a0 = 0.35875f, a1 = 0.48829f, a2 = 0.14128f, a3 = 0.01168f;
  const float fx = std::sqrt(x * x + y * y) + width / 2.0f;
  const float ow = 1.0f / width;
  return a0 - a1 * std::cos((2.0f * (float)M_PI * fx) * ow) + a2 * std::cos((4.0f * (float)M_PI * fx) * ow)
      - a3 * std::cos((6.0f * (float)M_PI * fx) * ow);
}

void Renderer::commit()
{
  sh = OcRenderer{};
  if (type == "pathtracer") sh.type = OC_RENDERER_PATHTRACER;
  else if (type == "scivis") sh.type = OC_RENDERER_SCIVIS;
  else if (type == "ao") sh.type = OC_RENDERER_AO;
  else if (type == "debug") {
    // DebugRenderer::commit (render/debug/DebugRenderer.cpp:18-75): "method", default eyeLight; unknown names -> testFrame
    sh.type = OC_RENDERER_DEBUG;
    static const struct { const char *name; uint32_t id; } methods[] = {{"rayDir", OC_DEBUG_RAY_DIR}, {"eyeLight", OC_DEBUG_EYE_LIGHT},
        {"Ng", OC_DEBUG_NG}, {"Ns", OC_DEBUG_NS}, {"color", OC_DEBUG_COLOR}, {"texCoord", OC_DEBUG_TEX_COORD},
        {"backfacing_Ng", OC_DEBUG_BACKFACING_NG}, {"backfacing_Ns", OC_DEBUG_BACKFACING_NS}, {"dPds", OC_DEBUG_DPDS},
        {"dPdt", OC_DEBUG_DPDT}, {"primID", OC_DEBUG_PRIM_ID}, {"geomID", OC_DEBUG_GEOM_ID}, {"instID", OC_DEBUG_INST_ID},
        {"volume", OC_DEBUG_VOLUME}};
    const std::string method = getParamString("method", "eyeLight");
    uint32_t id = OC_DEBUG_TEST_FRAME;
    for (const auto &m : methods)
      if (method == m.name) id = m.id;
    if (id == OC_DEBUG_VOLUME)
      throw std::runtime_error("cuda device: debug renderer method '" + method + "' is not supported");
    sh.debugMethod = id;
  }
  sh.spp = std::max(1, getParam<int>("pixelSamples", 1));
  sh.mipBiasFactor = std::pow(2.f, getParam<float>("mipMapBias", 0.f)); // Renderer.cpp:41,91
  sh.maxDepth = (uint32_t)std::max(0, getParam<int>("maxPathLength", 20));
  sh.minContribution = getParam<float>("minContribution", 0.001f);
  errorThreshold = getParam<float>("varianceThreshold", 0.f);
  // Renderer.cpp:61-80
  maxDepthTexture = getParamObject<Texture2D>("map_maxDepth");
  backplate = getParamObject<Texture2D>("map_backplate");
  if (maxDepthTexture && (maxDepthTexture->sh.format != OSP_TEXTURE_R32F || maxDepthTexture->sh.filter != OSP_TEXTURE_FILTER_NEAREST))
    postStatusMsg(OSP_LOG_WARNING, "maxDepthTexture provided to the renderer needs to be of type OSP_TEXTURE_R32F and have the OSP_TEXTURE_FILTER_NEAREST flag");
  sh.maxDepthTexture = maxDepthTexture ? maxDepthTexture->devPtr() : nullptr;
  sh.backplate = backplate ? backplate->devPtr() : nullptr;
  const vec3f bg3 = getParam<vec3f>("backgroundColor", make_vec3f(getParam<float>("backgroundColor", 0.f)));
  const vec4f bg = getParam<vec4f>("backgroundColor", make_vec4f(bg3, 0.f));
  sh.bgColor[0] = bg.x; sh.bgColor[1] = bg.y; sh.bgColor[2] = bg.z; sh.bgColor[3] = bg.w;

  materialData = getParamData(*this, "material", OSP_MATERIAL);
  materialSlots.clear();
  commitCount++;
  if (materialData) {
    std::vector<uint32_t> slots(materialData->size());
    for (size_t i = 0; i < slots.size(); i++) {
      Material *m = dynamic_cast<Material *>(materialData->objectAt(i));
      if (!m) throw std::runtime_error("renderer 'material' array holds a non-material object");
      slots[i] = m->slot;
    }
    materialSlots = slots;
    matList.upload(slots.data(), slots.size() * sizeof(uint32_t));
    sh.materials = matList.as<uint32_t>();
    sh.numMaterials = (uint32_t)slots.size();
  }

  // pixel filter (Renderer.cpp:146-178, pf/PixelFilter.cpp:23-69, LUTPixelFilter.ispc:68-142)
  const uint32_t pf = getParam<uint32_t>("pixelFilter", OSP_PIXELFILTER_GAUSS);
  if (pf == OSP_PIXELFILTER_POINT) { sh.pfType = OC_PF_POINT; sh.pfWidth = 0.f; }
  else if (pf == OSP_PIXELFILTER_BOX) { sh.pfType = OC_PF_BOX; sh.pfWidth = 1.f; }
  else {
    const float width = pf == OSP_PIXELFILTER_MITCHELL ? 4.f : 3.f;
    const int lut = (int)std::ceil(width) * 5; // LUTPIXELFILTER_PER_PIXEL_BINS
    std::vector<float> cdfx((size_t)lut * lut), cdfy(lut);
    for (int y = 0; y < lut; y++) {
      const float fy = (y + 0.5f) * (1.0f / (float)lut);
      std::vector<float> row(lut);
      for (int x = 0; x < lut; x++) {
        const float fx = (x + 0.5f) * (1.0f / (float)lut);
        row[x] = std::max(0.0f, filterEval(pf == OSP_PIXELFILTER_MITCHELL || pf == OSP_PIXELFILTER_BLACKMAN_HARRIS ? pf : OSP_PIXELFILTER_GAUSS,
                                    width, (fx - 0.5f) * width, (fy - 0.5f) * width));
      }
      cdfy[y] = distribution1D_create(row);
      std::copy(row.begin(), row.end(), cdfx.begin() + (size_t)y * lut);
    }
    distribution1D_create(cdfy);
    pfCdfX.upload(cdfx.data(), cdfx.size() * sizeof(float));
    pfCdfY.upload(cdfy.data(), cdfy.size() * sizeof(float));
    sh.pfType = OC_PF_LUT;
    sh.pfWidth = width;
    sh.pfLutW = sh.pfLutH = (uint32_t)lut;
    sh.pfCdfX = pfCdfX.as<float>();
    sh.pfCdfY = pfCdfY.as<float>();
  }

  if (sh.type == OC_RENDERER_PATHTRACER) {
    sh.rouletteDepth = getParam<int>("roulettePathLength", 5);
    sh.maxScatteringEvents = getParam<int>("maxScatteringEvents", 20);
    sh.maxRadiance = getParam<float>("maxContribution", INFINITY);
    sh.numLightSamples = getParam<int>("lightSamples", -1);
    sh.limitIndirectLightSamples = getParam<bool>("limitIndirectLightSamples", true);
    const vec4f scp = getParam<vec4f>("shadowCatcherPlane", make_vec4f(0.f, 0.f, 0.f, 0.f));
    if (scp.x != 0.f || scp.y != 0.f || scp.z != 0.f)
      postStatusMsg(OSP_LOG_WARNING, "cuda device: the shadow catcher plane is not supported");
    sh.geometryLights = getParam<bool>("geometryLights", true);
    sh.backgroundRefraction = getParam<bool>("backgroundRefraction", false);
  } else if (sh.type == OC_RENDERER_SCIVIS) {
    sh.shadowsEnabled = getParam<bool>("shadows", false);
    sh.visibleLights = getParam<bool>("visibleLights", false);
    sh.aoSamples = getParam<int>("aoSamples", 0);
    sh.aoRadius = getParam<float>("aoDistance", 1e20f);
    sh.volumeSamplingRate = getParam<float>("volumeSamplingRate", 1.f);
  } else {
    sh.aoSamples = getParam<int>("aoSamples", 1);
    sh.aoRadius = getParam<float>("aoDistance", getParam<float>("aoRadius", 1e20f));
    sh.aoIntensity = getParam<float>("aoIntensity", 1.f);
    sh.volumeSamplingRate = getParam<float>("volumeSamplingRate", 1.f);
  }
}

} // namespace cuda
} // namespace ospray
