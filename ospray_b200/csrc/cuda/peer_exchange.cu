// Peer-memory tile exchange: the multi-device frame composition without a staging copy or a collective call.
//
// Reference semantics: modules/multiDevice/MultiDeviceLoadBalancer.cpp:19-60 renders every sub-device's tiles and
// fb/LocalFB.cpp:286-353 (writeTiles) composes them into the row-major image of the displayed framebuffer. Here a
// sub-device is one process on its own B200 and the box's GPUs are NVLink/NVSwitch peers: rank 0 exports its display
// image and a small synchronisation block through CUDA IPC, every other rank maps them and its pushTilesKernel stores the
// tiles it owns (tile % numRanks == rank) straight into rank 0's image with 16-byte stores over NVLink -- pack, transfer
// and unpack in one kernel. Ordering is a two-flag handshake per frame (epoch), all system-scope:
//   ready[r]  (in rank r's block, written by rank 0): rank 0 has consumed frame epoch-1, the image may be overwritten
//   arrive[r] (in rank 0's block, written by rank r): every tile of rank r for this epoch is visible in rank 0's memory
// Every spin is bounded (OC_PEER_TIMEOUT_NS); a timeout sets the block's error word instead of hanging the GPU.
#include <cstdint>
#include <cstring>
#include "runtime.h"
#include "../../../include/ospcuda.h"

namespace oc {

#define OC_PEER_TIMEOUT_NS 4000000000ull
enum { PEER_READY = 0, PEER_COUNTER = 1, PEER_ERROR = 2, PEER_ARRIVE = 8 }; // word offsets in a block of OSPCUDA_PEER_BLOCK_WORDS

__device__ __forceinline__ uint64_t globalTimerNs()
{
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ldAcquireSys(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseSys(uint32_t *p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// epochs only grow; the difference is taken modulo 2^32 so that a wrap-around after 4 G frames still compares right
__device__ __forceinline__ bool spinUntil(const uint32_t *flag, uint32_t epoch, uint32_t *errWord)
{
  const uint64_t t0 = globalTimerNs();
  while ((int32_t)(ldAcquireSys(flag) - epoch) < 0) {
    __nanosleep(64);
    if (globalTimerNs() - t0 > OC_PEER_TIMEOUT_NS) { atomicExch(errWord, 1u); return false; }
  }
  return true;
}

struct PeerBlocks { uint32_t *block[OSPCUDA_MAX_PEERS]; };

// rank 0, start of a frame: the previous frame's image has been consumed in stream order
__global__ void postReadyKernel(PeerBlocks peers, uint32_t numRanks, uint32_t epoch)
{
  const uint32_t r = threadIdx.x;
  if (r >= 1 && r < numRanks) stReleaseSys(peers.block[r] + PEER_READY, epoch);
}

// rank 0, end of a frame: all ranks' tiles of this epoch have landed
__global__ void waitArriveKernel(uint32_t *rootBlock, uint32_t numRanks, uint32_t epoch)
{
  const uint32_t r = threadIdx.x;
  if (r >= 1 && r < numRanks) spinUntil(rootBlock + PEER_ARRIVE + r, epoch, rootBlock + PEER_ERROR);
  __threadfence_system();
}

// rank != 0: one CTA per owned 64x64 tile, 16 threads per tile row, 16-byte lanes (a 64-pixel sRGBA8 row = 256 B)
__global__ void __launch_bounds__(256) pushTilesKernel(uint8_t *__restrict__ peerImage, const uint8_t *__restrict__ image, int sizeX,
    int sizeY, uint32_t bpp, uint32_t numRanks, uint32_t rank, uint32_t tilesX, uint32_t ownedTiles, uint32_t *localBlock,
    uint32_t *rootBlock, uint32_t epoch)
{
  __shared__ bool go;
  if (threadIdx.x == 0) go = spinUntil(localBlock + PEER_READY, epoch, localBlock + PEER_ERROR);
  __syncthreads();
  if (go) {
    const uint32_t tileID = rank + blockIdx.x * numRanks;
    const uint32_t tx = tileID % tilesX, ty = tileID / tilesX;
    const int x0 = tx * 64, y0 = ty * 64;
    const int w = min(64, sizeX - x0), h = min(64, sizeY - y0);
    const uint32_t rowBytes = (uint32_t)w * bpp;
    for (int r = threadIdx.x >> 4; r < h; r += blockDim.x >> 4) {
      const size_t off = ((size_t)(y0 + r) * sizeX + x0) * bpp;
      if (((off | rowBytes) & 15u) == 0) {
        for (uint32_t b = (threadIdx.x & 15) * 16; b < rowBytes; b += 256)
          *reinterpret_cast<uint4 *>(peerImage + off + b) = *reinterpret_cast<const uint4 *>(image + off + b);
      } else {
        for (uint32_t b = (threadIdx.x & 15) * 4; b < rowBytes; b += 64)
          *reinterpret_cast<uint32_t *>(peerImage + off + b) = *reinterpret_cast<const uint32_t *>(image + off + b);
      }
    }
  }
  __threadfence_system(); // this thread's peer stores are ordered before whatever follows the barrier
  __syncthreads();
  if (threadIdx.x == 0) {
    // the last CTA of the grid announces the rank's arrival (even after a timeout: rank 0 must not hang on us)
    if (atomicAdd(localBlock + PEER_COUNTER, 1u) == ownedTiles - 1u) {
      localBlock[PEER_COUNTER] = 0u;
      __threadfence_system();
      stReleaseSys(rootBlock + PEER_ARRIVE + rank, epoch);
    }
  }
}

} // namespace oc

using namespace oc;

static_assert(sizeof(cudaIpcMemHandle_t) == OSPCUDA_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int ospcuda_peer_block_alloc(uint32_t **d_block)
{
  OC_TRY(cudaMalloc((void **)d_block, sizeof(uint32_t) * OSPCUDA_PEER_BLOCK_WORDS));
  OC_TRY(cudaMemset(*d_block, 0, sizeof(uint32_t) * OSPCUDA_PEER_BLOCK_WORDS));
  return 0;
}

extern "C" int ospcuda_ipc_export(const void *d_ptr, unsigned char *handle)
{
  cudaIpcMemHandle_t h;
  OC_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle, &h, sizeof(h));
  return 0;
}

extern "C" int ospcuda_ipc_open(const unsigned char *handle, void **d_ptr)
{
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  OC_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int ospcuda_ipc_close(void *d_ptr)
{
  if (d_ptr) OC_TRY(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

extern "C" int ospcuda_peer_post_ready(uint32_t *const *peerBlocks, uint32_t numRanks, uint32_t epoch)
{
  if (numRanks > OSPCUDA_MAX_PEERS) return oc_set_error(cudaErrorInvalidValue, "ospcuda_peer_post_ready: more than OSPCUDA_MAX_PEERS ranks");
  PeerBlocks p;
  for (uint32_t r = 0; r < OSPCUDA_MAX_PEERS; r++) p.block[r] = r < numRanks ? peerBlocks[r] : nullptr;
  postReadyKernel<<<1, 32, 0, (cudaStream_t)ospcuda_stream()>>>(p, numRanks, epoch);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_peer_post_ready");
}

extern "C" int ospcuda_peer_wait_arrive(uint32_t *d_rootBlock, uint32_t numRanks, uint32_t epoch)
{
  if (numRanks > OSPCUDA_MAX_PEERS) return oc_set_error(cudaErrorInvalidValue, "ospcuda_peer_wait_arrive: more than OSPCUDA_MAX_PEERS ranks");
  waitArriveKernel<<<1, 32, 0, (cudaStream_t)ospcuda_stream()>>>(d_rootBlock, numRanks, epoch);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_peer_wait_arrive");
}

extern "C" int ospcuda_peer_push_tiles(void *peerImage, const void *d_image, int sizeX, int sizeY, uint32_t bytesPerPixel,
    uint32_t numRanks, uint32_t rank, uint32_t *d_localBlock, uint32_t *peerRootBlock, uint32_t epoch)
{
  if (bytesPerPixel % 4 != 0 || numRanks == 0 || rank == 0 || rank >= numRanks || numRanks > OSPCUDA_MAX_PEERS)
    return oc_set_error(cudaErrorInvalidValue, "ospcuda_peer_push_tiles: bad rank / bytesPerPixel");
  const uint32_t tilesX = (sizeX + 63) / 64, tilesY = (sizeY + 63) / 64, total = tilesX * tilesY;
  const uint32_t owned = rank < total ? (total - rank + numRanks - 1) / numRanks : 0;
  if (owned == 0) return oc_set_error(cudaErrorInvalidValue, "ospcuda_peer_push_tiles: rank owns no tile (more ranks than tiles)");
  pushTilesKernel<<<owned, 256, 0, (cudaStream_t)ospcuda_stream()>>>((uint8_t *)peerImage, (const uint8_t *)d_image, sizeX, sizeY,
      bytesPerPixel, numRanks, rank, tilesX, owned, d_localBlock, peerRootBlock, epoch);
  oc_count_launch();
  return oc_check(cudaGetLastError(), "ospcuda_peer_push_tiles");
}

extern "C" int ospcuda_peer_error(const uint32_t *d_block, uint32_t *timedOut)
{
  OC_TRY(cudaMemcpy(timedOut, d_block + PEER_ERROR, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 0;
}
