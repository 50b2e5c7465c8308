// FrameBuffer + rendering future of the cuda device.
// reference: fb/FrameBuffer.cpp:31-137, fb/LocalFB.cpp:49-180 (buffers), :238-261 (task list),
// :459-478 (postProcess), :489-528 (map/unmap), fb/frame_ops/Variance.cpp:32-46,
// render/RenderingFuture.h:13-103.
#include <algorithm>
#include "CudaDevice.h"
#include "Objects.h"

namespace ospray {
namespace cuda {

FrameBuffer::FrameBuffer(CudaDevice &d, const vec2i &sz, OSPFrameBufferFormat format, uint32_t ch)
    : device(d), size(sz), colorFormat(format), channels(ch), doAccum(ch & OSP_FB_ACCUM),
      hasVariance((ch & OSP_FB_COLOR) && (ch & OSP_FB_VARIANCE) && (ch & OSP_FB_ACCUM))
{
  managedObjectType = OSP_FRAMEBUFFER;
  if (size.x <= 0 || size.y <= 0) throw std::runtime_error("framebuffer has invalid size. Dimensions must be greater than 0");
  const size_t np = (size_t)size.x * size.y;
  auto mk = [&](size_t bytes) {
    buffers.emplace_back(new DevMem);
    buffers.back()->reserve(bytes);
    checkCuda(ospcuda_memset(buffers.back()->p, 0, bytes), "framebuffer clear");
    return buffers.back().get();
  };
  sh = OcFrameBuffer{};
  sh.sizeX = size.x; sh.sizeY = size.y;
  sh.frameID = -1;
  sh.channels = channels;
  sh.colorFormat = format;
  sh.targetFrames = !doAccum;
  if (ch & OSP_FB_COLOR) { colorBuf = mk(np * 16); sh.color = colorBuf->as<OcVec4f>(); }
  if (ch & OSP_FB_DEPTH) { depthBuf = mk(np * 4); sh.depth = depthBuf->as<float>(); }
  if (hasVariance) {
    varianceBuf = mk(np * 16); sh.variance = varianceBuf->as<OcVec4f>();
    const size_t tasks = (size_t)((size.x + 7) / 8) * ((size.y + 7) / 8);
    taskErrBuf = mk(tasks * 4);
    taskErrors.assign(tasks, INFINITY);
  }
  if (ch & OSP_FB_NORMAL) { normalBuf = mk(np * 12); sh.normal = normalBuf->as<OcVec3f>(); }
  if (ch & OSP_FB_FIRST_NORMAL) { firstNormalBuf = mk(np * 12); sh.firstNormal = firstNormalBuf->as<OcVec3f>(); }
  if (ch & OSP_FB_ALBEDO) { albedoBuf = mk(np * 12); sh.albedo = albedoBuf->as<OcVec3f>(); }
  if (ch & OSP_FB_POSITION) { positionBuf = mk(np * 12); sh.position = positionBuf->as<OcVec3f>(); }
  if (ch & OSP_FB_ID_PRIMITIVE) { primBuf = mk(np * 4); sh.primID = primBuf->as<uint32_t>(); }
  if (ch & OSP_FB_ID_OBJECT) { objBuf = mk(np * 4); sh.objID = objBuf->as<uint32_t>(); }
  if (ch & OSP_FB_ID_INSTANCE) { instBuf = mk(np * 4); sh.instID = instBuf->as<uint32_t>(); }
  if ((ch & OSP_FB_COLOR) && format != OSP_FB_NONE) convBuf = mk(np * (format == OSP_FB_RGBA32F ? 16 : 4));
  counterBuf = mk(6 * sizeof(unsigned long long)); // 4 ray counters + the frame control words (cancel, progress, total)
  sh.rayCounters = counterBuf->as<unsigned long long>();
  sh.control = reinterpret_cast<uint32_t *>(counterBuf->as<unsigned long long>() + 4);
  sh.numRanks = device.numRanks;
  sh.rank = device.rank;
}

FrameBuffer::~FrameBuffer()
{
  for (auto &s : staging) ospcuda_free_host(s.second.host);
  for (void *m : extraMapped) ospcuda_free_host(m);
}

void FrameBuffer::commit()
{
  imageOpData = getParamData(*this, "imageOperation", OSP_IMAGE_OPERATION);
  sh.targetFrames = std::max(0, getParam<int>("targetFrames", 0));
  sh.projectedDepth = getParam<bool>("projectedDepth", false);
  if (!doAccum) sh.targetFrames = 1;
  toneMappers.clear();
  ppBuf = nullptr;
  if (imageOpData && colorBuf) {
    for (size_t i = 0; i < imageOpData->size(); i++) {
      ImageOp *op = dynamic_cast<ImageOp *>(imageOpData->objectAt(i));
      if (op && op->type == "tonemapper") toneMappers.emplace_back(op);
    }
    if (!toneMappers.empty()) {
      // one post-processed colour buffer for the life of the framebuffer, however often it is committed
      if (!ppStore) {
        buffers.emplace_back(new DevMem);
        buffers.back()->reserve((size_t)size.x * size.y * 16);
        ppStore = buffers.back().get();
      }
      ppBuf = ppStore;
    }
  }
}

void FrameBuffer::clear()
{
  sh.frameID = -1;
  if (hasVariance) {
    mtGen.seed(43);
    frameVariance = INFINITY;
    std::fill(taskErrors.begin(), taskErrors.end(), INFINITY);
    taskErrorsValid = false;
  }
}

void FrameBuffer::beginFrame()
{
  sh.frameID = doAccum ? sh.frameID + 1 : 0;
  sh.numRanks = device.numRanks;
  sh.rank = device.rank;
  if (hasVariance) {
    const bool evenFrame = (sh.frameID & 1) == 0;
    if (evenFrame) pickOdd = mtGen() & 1;
    sh.accumulateVariance = evenFrame != pickOdd;
  }
  checkCuda(ospcuda_memset(counterBuf->p, 0, 6 * sizeof(unsigned long long)), "counter reset");
}

void FrameBuffer::postProcess()
{
  const uint32_t np = (uint32_t)size.x * (uint32_t)size.y;
  const bool oddFrame = (sh.frameID & 1) == 1;
  if (hasVariance && oddFrame) {
    checkCuda(ospcuda_fb_variance(sh.color, sh.variance, size.x, size.y, 8, taskErrBuf->as<float>()), "variance op");
    taskErrorsPending = true;
  }
  if (!colorBuf) return;
  const OcVec4f *in = sh.color;
  // chained tonemappers: all but the last run as plain float4 passes
  for (size_t i = 0; i + 1 < toneMappers.size(); i++) {
    checkCuda(ospcuda_fb_tonemap(in, ppBuf->as<OcVec4f>(), np, &toneMappers[i]->tm), "tonemapper");
    in = ppBuf->as<OcVec4f>();
  }
  const OcToneMapper *last = toneMappers.empty() ? nullptr : &toneMappers.back()->tm;
  if (convBuf || last)
    checkCuda(ospcuda_fb_postprocess_tiles(in, last ? ppBuf->as<OcVec4f>() : nullptr, convBuf ? convBuf->p : nullptr, size.x, size.y,
                  convBuf ? (uint32_t)colorFormat : 0u, last, sh.numRanks ? sh.numRanks : 1u, sh.rank),
        "colour conversion");
}

void *FrameBuffer::devicePointer(OSPFrameBufferChannel channel) const
{
  switch (channel) {
  case OSP_FB_COLOR: return convBuf ? convBuf->p : (ppBuf ? ppBuf->p : (colorBuf ? colorBuf->p : nullptr));
  case OSP_FB_DEPTH: return depthBuf ? depthBuf->p : nullptr;
  case OSP_FB_NORMAL: return normalBuf ? normalBuf->p : nullptr;
  case OSP_FB_FIRST_NORMAL: return firstNormalBuf ? firstNormalBuf->p : nullptr;
  case OSP_FB_POSITION: return positionBuf ? positionBuf->p : nullptr;
  case OSP_FB_ALBEDO: return albedoBuf ? albedoBuf->p : nullptr;
  case OSP_FB_ID_PRIMITIVE: return primBuf ? primBuf->p : nullptr;
  case OSP_FB_ID_OBJECT: return objBuf ? objBuf->p : nullptr;
  case OSP_FB_ID_INSTANCE: return instBuf ? instBuf->p : nullptr;
  case OSP_FB_ACCUM: return colorBuf ? colorBuf->p : nullptr;       // extension: raw accumulation buffer
  case OSP_FB_VARIANCE: return varianceBuf ? varianceBuf->p : nullptr;
  default: return nullptr;
  }
}

size_t FrameBuffer::bytesPerPixel(OSPFrameBufferChannel channel) const
{
  switch (channel) {
  case OSP_FB_COLOR: return convBuf ? (colorFormat == OSP_FB_RGBA32F ? 16 : 4) : 16;
  case OSP_FB_DEPTH: case OSP_FB_ID_PRIMITIVE: case OSP_FB_ID_OBJECT: case OSP_FB_ID_INSTANCE: return 4;
  case OSP_FB_NORMAL: case OSP_FB_FIRST_NORMAL: case OSP_FB_POSITION: case OSP_FB_ALBEDO: return 12;
  case OSP_FB_ACCUM: case OSP_FB_VARIANCE: return 16;
  default: return 0;
  }
}

const void *FrameBuffer::map(OSPFrameBufferChannel channel)
{
  if (channel == OSP_FB_ACCUM || channel == OSP_FB_VARIANCE) return nullptr;
  void *src = devicePointer(channel);
  if (!src) return nullptr;
  const size_t bytes = (size_t)size.x * size.y * bytesPerPixel(channel);
  // pinned staging is cached per channel: cudaMallocHost costs more than the copy itself
  void *host = nullptr;
  auto st = staging.find((int)channel);
  if (st != staging.end() && !st->second.inUse) {
    host = st->second.host;
    st->second.inUse = true;
  } else {
    checkCuda(ospcuda_malloc_host(&host, bytes), "pinned allocation");
    if (st == staging.end()) staging[(int)channel] = Staging{host, true};
    else extraMapped.insert(host); // the channel is already mapped: one-off buffer
  }
  checkCuda(ospcuda_memcpy_d2h(host, src, bytes), "framebuffer read-back"); // orders after the frame on the render stream
  this->refInc();
  return host;
}

void FrameBuffer::unmap(const void *m)
{
  if (!m) return;
  auto ex = extraMapped.find((void *)m);
  if (ex != extraMapped.end()) {
    ospcuda_free_host(*ex);
    extraMapped.erase(ex);
  } else {
    for (auto &s : staging)
      if (s.second.host == m) s.second.inUse = false;
  }
  this->refDec();
}

float FrameBuffer::getVariance()
{
  if (!hasVariance) return frameVariance;
  if (taskErrorsPending) {
    checkCuda(ospcuda_memcpy_d2h(taskErrors.data(), taskErrBuf->p, taskErrors.size() * sizeof(float)), "task error read-back");
    taskErrorsPending = false;
    taskErrorsValid = true;
  }
  if (!taskErrorsValid) return frameVariance;
  float maxErr = 0.f, sumAct = 0.f;
  int active = 0;
  for (float e : taskErrors) {
    maxErr = std::max(maxErr, e);
    if (e >= errorThreshold) { sumAct += e; active++; }
  }
  return active ? sumAct / active : maxErr;
}

// ---- Future ------------------------------------------------------------------------------------
Future::Future(FrameBuffer *f, Renderer *r, Camera *c, World *w) : fb(f), renderer(r), camera(c), world(w)
{
  managedObjectType = OSP_FUTURE;
  checkCuda(ospcuda_event_create(&evStart), "event");
  checkCuda(ospcuda_event_create(&evRendered), "event");
  checkCuda(ospcuda_event_create(&evFinished), "event");
}

Future::~Future()
{
  // releasing a future blocks until the frame is done (RenderingFuture.h:46-53): ospBenchmark and
  // ospTestSuite synchronise this way.
  try { wait(OSP_TASK_FINISHED); } catch (...) {}
  ospcuda_event_destroy(evStart);
  ospcuda_event_destroy(evRendered);
  ospcuda_event_destroy(evFinished);
}

void Future::finalizeHost()
{
  if (finalized) return;
  finalized = true;
  checkCuda(ospcuda_memcpy_d2h(fb->rayCounts, fb->sh.rayCounters, sizeof(fb->rayCounts)), "ray counter read-back");
}

bool Future::isFinished(OSPSyncEvent ev)
{
  if (ev == OSP_NONE_FINISHED) return true;
  void *e = (ev == OSP_WORLD_RENDERED) ? evRendered : evFinished;
  if (ev == OSP_WORLD_COMMITTED) return true;
  const int q = ospcuda_event_query(e);
  if (q < 0) checkCuda(q, "event query");
  return q == 1;
}

void Future::wait(OSPSyncEvent ev)
{
  if (ev == OSP_NONE_FINISHED || ev == OSP_WORLD_COMMITTED) return;
  void *e = (ev == OSP_WORLD_RENDERED) ? evRendered : evFinished;
  checkCuda(ospcuda_event_synchronize(e), "event wait");
  if (ev != OSP_WORLD_RENDERED) finalizeHost();
}

float Future::duration()
{
  wait(OSP_TASK_FINISHED);
  float ms = 0.f;
  checkCuda(ospcuda_event_elapsed_ms(evStart, evFinished, &ms), "event elapsed");
  return ms * 1e-3f;
}

} // namespace cuda
} // namespace ospray
