// libospcuda runtime: device selection, memory, the render stream, events, error text.
// Stands in for the reference's devicert::Device (ospray/common/DeviceRT.h:140-197; ISPC
// implementation modules/cpu/common/DeviceRTImpl_ispc.cpp:215-363): one in-order queue
// == one CUDA stream; AsyncEvent == cudaEvent pair.
#include "runtime.h"
#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

static thread_local std::string g_lastError;
static std::string g_lastErrorGlobal;
static std::mutex g_errMutex;
// one render stream per CUDA device of the process; `g_stream` / `g_device` are the CURRENT ones: ospcuda_init() and
// ospcuda_make_current() switch, every other entry point works on the current device (the in-process multi-device,
// csrc/device/MultiDevice.cpp, switches before each forwarded call)
#define OC_MAX_DEVICES 32
struct DevCtx { cudaStream_t stream = nullptr, own = nullptr; bool ready = false; };
static DevCtx g_ctx[OC_MAX_DEVICES];
static cudaStream_t g_stream = nullptr;
static int g_device = -1;
static std::atomic<unsigned long long> g_launches{0};

int oc_set_error(cudaError_t e, const char *where)
{
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
  g_lastError = buf;
  {
    std::lock_guard<std::mutex> l(g_errMutex);
    g_lastErrorGlobal = buf;
  }
  cudaGetLastError(); // clear sticky-free errors
  return e == cudaSuccess ? -1 : (int)e;
}
void oc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" {

const char *ospcuda_last_error(void)
{
  if (!g_lastError.empty()) return g_lastError.c_str();
  std::lock_guard<std::mutex> l(g_errMutex);
  return g_lastErrorGlobal.c_str();
}

int ospcuda_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ospcuda_init(int deviceOrdinal)
{
  int n = 0;
  OC_TRY(cudaGetDeviceCount(&n));
  if (n == 0) return oc_set_error(cudaErrorNoDevice, "ospcuda_init: no CUDA device (the cuda OSPRay device has no CPU fallback)");
  if (deviceOrdinal < 0 || deviceOrdinal >= n || deviceOrdinal >= OC_MAX_DEVICES)
    return oc_set_error(cudaErrorInvalidDevice, "ospcuda_init: bad ordinal");
  OC_TRY(cudaSetDevice(deviceOrdinal));
  DevCtx &c = g_ctx[deviceOrdinal];
  if (!c.ready) {
    OC_TRY(cudaStreamCreateWithFlags(&c.own, cudaStreamNonBlocking));
    c.stream = c.own;
    // deep traversal + shading call chains: give local-memory stacks room
    cudaDeviceSetLimit(cudaLimitStackSize, 4096);
    c.ready = true;
  }
  g_device = deviceOrdinal;
  g_stream = c.stream;
  return 0;
}

int ospcuda_make_current(int deviceOrdinal)
{
  if (deviceOrdinal == g_device) return 0;
  if (deviceOrdinal < 0 || deviceOrdinal >= OC_MAX_DEVICES || !g_ctx[deviceOrdinal].ready) return ospcuda_init(deviceOrdinal);
  OC_TRY(cudaSetDevice(deviceOrdinal));
  g_device = deviceOrdinal;
  g_stream = g_ctx[deviceOrdinal].stream;
  return 0;
}
int ospcuda_current_device(void) { return g_device; }

int ospcuda_enable_peer(int peerOrdinal)
{
  if (peerOrdinal == g_device) return 0;
  int can = 0;
  OC_TRY(cudaDeviceCanAccessPeer(&can, g_device, peerOrdinal));
  if (!can) return oc_set_error(cudaErrorPeerAccessUnsupported, "ospcuda_enable_peer: no peer access between the two devices");
  const cudaError_t e = cudaDeviceEnablePeerAccess(peerOrdinal, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  OC_TRY(e);
  return 0;
}

void *ospcuda_stream(void) { return (void *)g_stream; }

int ospcuda_set_stream(void *stream)
{
  // adopt an external stream (e.g. torch's current stream so that NCCL work and CUDA-event timing order
  // naturally with the frame); NULL restores the library's own stream
  if (g_device < 0) return oc_set_error(cudaErrorInitializationError, "ospcuda_set_stream before ospcuda_init");
  if (g_stream) OC_TRY(cudaStreamSynchronize(g_stream));
  DevCtx &c = g_ctx[g_device];
  c.stream = stream ? (cudaStream_t)stream : c.own;
  g_stream = c.stream;
  return 0;
}
unsigned long long ospcuda_kernel_launch_count(void) { return g_launches.load(); }

int ospcuda_malloc(void **ptr, size_t bytes)
{
  *ptr = nullptr;
  if (bytes == 0) return 0;
  OC_TRY(cudaMalloc(ptr, bytes));
  return 0;
}
int ospcuda_free(void *ptr)
{
  if (ptr) OC_TRY(cudaFree(ptr));
  return 0;
}
int ospcuda_malloc_host(void **ptr, size_t bytes)
{
  *ptr = nullptr;
  if (bytes == 0) return 0;
  OC_TRY(cudaMallocHost(ptr, bytes));
  return 0;
}
int ospcuda_free_host(void *ptr)
{
  if (ptr) OC_TRY(cudaFreeHost(ptr));
  return 0;
}
int ospcuda_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
  if (bytes == 0) return 0;
  OC_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
  OC_TRY(cudaStreamSynchronize(g_stream)); // source may be pageable app memory
  return 0;
}
int ospcuda_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
  if (bytes == 0) return 0;
  OC_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
  OC_TRY(cudaStreamSynchronize(g_stream));
  return 0;
}
int ospcuda_memcpy_d2h_async(void *dst, const void *src, size_t bytes)
{
  if (bytes == 0) return 0;
  OC_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
  return 0;
}
int ospcuda_memset(void *dst, int value, size_t bytes)
{
  if (bytes == 0) return 0;
  OC_TRY(cudaMemsetAsync(dst, value, bytes, g_stream));
  return 0;
}
static cudaStream_t auxStream()
{
  static thread_local cudaStream_t s = nullptr;
  static thread_local int dev = -1;
  int cur = 0;
  cudaGetDevice(&cur);
  if (!s || dev != cur) { cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); dev = cur; }
  return s;
}
int ospcuda_aux_poke_u32(void *d_word, uint32_t value)
{
  OC_TRY(cudaMemcpyAsync(d_word, &value, sizeof(value), cudaMemcpyHostToDevice, auxStream()));
  OC_TRY(cudaStreamSynchronize(auxStream()));
  return 0;
}
int ospcuda_aux_peek_u32(const void *d_word, uint32_t *value)
{
  OC_TRY(cudaMemcpyAsync(value, d_word, sizeof(*value), cudaMemcpyDeviceToHost, auxStream()));
  OC_TRY(cudaStreamSynchronize(auxStream()));
  return 0;
}
int ospcuda_synchronize(void)
{
  OC_TRY(cudaStreamSynchronize(g_stream));
  return 0;
}
int ospcuda_event_create(void **ev)
{
  cudaEvent_t e;
  OC_TRY(cudaEventCreate(&e));
  *ev = (void *)e;
  return 0;
}
int ospcuda_event_destroy(void *ev)
{
  if (ev) OC_TRY(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
int ospcuda_event_record(void *ev)
{
  OC_TRY(cudaEventRecord((cudaEvent_t)ev, g_stream));
  return 0;
}
int ospcuda_stream_wait_event(void *ev)
{
  OC_TRY(cudaStreamWaitEvent(g_stream, (cudaEvent_t)ev, 0)); // the event may belong to another device's stream
  return 0;
}
int ospcuda_event_synchronize(void *ev)
{
  OC_TRY(cudaEventSynchronize((cudaEvent_t)ev));
  return 0;
}
int ospcuda_event_query(void *ev)
{
  cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  oc_set_error(e, "cudaEventQuery");
  return -1;
}
int ospcuda_event_elapsed_ms(void *start, void *stop, float *ms)
{
  OC_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}

} // extern "C"
