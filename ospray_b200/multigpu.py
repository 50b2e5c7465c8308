"""One-process-per-GPU tile partition and composition of the finished tiles on GPU 0: PeerTileExchange (stores over
NVLink peer memory, the default on a CUDA box) or TileGather (pack + one NCCL gather + unpack).

The rendering analogue of modules/multiDevice (MultiDevice.cpp:395-418: every sub-device owns the 64x64 tiles
with `tileID % numSubdevices == rank`; MultiDeviceLoadBalancer.cpp:19-60: render everywhere, then
LocalFrameBuffer::writeTiles composes the row-major image). Here a sub-device is one torchrun rank on its own
B200: the scene is replicated (every rank replays the same API calls), accumulation buffers stay sharded on
their GPU, and per displayed frame only the converted display channel (4 B/px) travels: each rank packs its
owned tiles (ospcuda_fb_pack_tiles), one NCCL gather moves the shards to rank 0 over NVLink, rank 0 scatters
them into the row-major image (ospcuda_fb_unpack_tiles). No other communication exists on the path.

`ops` abstracts the two staging kernels so that the host logic (ownership, shard sizes, gather order) is
testable on CPU with the gloo backend; the default CudaTileOps calls libospcuda and refuses CPU tensors."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import lib_path

TILE = 64


def tile_grid(width, height):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def owned_tiles(width, height, world_size, rank):
    """tile ids rendered by `rank` (MultiDevice.cpp:403-416)"""
    tx, ty = tile_grid(width, height)
    return list(range(rank, tx * ty, world_size))


def max_owned(width, height, world_size):
    tx, ty = tile_grid(width, height)
    return (tx * ty + world_size - 1) // world_size


class CudaTileOps:
    """pack / unpack through the C-ABI of libospcuda.so (device pointers only)"""

    def __init__(self, shared_stream=False):
        """shared_stream: the render stream IS torch's current stream (capi.set_render_stream), so the staging
        kernels and the NCCL gather are ordered by the stream and no host synchronisation is needed"""
        self.shared_stream = shared_stream
        L = C.CDLL(lib_path("libospcuda.so"), mode=C.RTLD_GLOBAL)
        for fn in (L.ospcuda_fb_pack_tiles, L.ospcuda_fb_unpack_tiles):
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ospcuda_last_error.restype = C.c_char_p
        L.ospcuda_synchronize.restype = C.c_int
        self.L = L

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"libospcuda: {self.L.ospcuda_last_error().decode()}")

    def pack(self, packed, image_ptr, width, height, bpp, world_size, rank):
        if not packed.is_cuda:
            raise RuntimeError("CudaTileOps.pack needs CUDA tensors: the cuda device has no CPU path")
        self._check(self.L.ospcuda_fb_pack_tiles(packed.data_ptr(), image_ptr, width, height, bpp, world_size, rank))
        if not self.shared_stream:
            self._check(self.L.ospcuda_synchronize())  # staging is consumed by NCCL on torch's stream

    def unpack(self, image_ptr, packed, width, height, bpp, world_size, src_rank):
        if not packed.is_cuda:
            raise RuntimeError("CudaTileOps.unpack needs CUDA tensors: the cuda device has no CPU path")
        self._check(self.L.ospcuda_fb_unpack_tiles(image_ptr, packed.data_ptr(), width, height, bpp, world_size, src_rank))

    def synchronize(self):
        if not self.shared_stream:
            self._check(self.L.ospcuda_synchronize())


class TileGather:
    """gathers one framebuffer channel (bpp bytes per pixel) of a width x height image to rank 0"""

    def __init__(self, width, height, bpp, device, ops=None, group=None):
        self.width, self.height, self.bpp = width, height, bpp
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ops = ops if ops is not None else CudaTileOps()
        self.shard_bytes = max_owned(width, height, self.world_size) * TILE * TILE * bpp
        self.device = device
        self.packed = torch.zeros(self.shard_bytes, dtype=torch.uint8, device=device)
        self.recv = ([torch.zeros(self.shard_bytes, dtype=torch.uint8, device=device) for _ in range(self.world_size)]
                     if self.rank == 0 and self.world_size > 1 else None)
        self.bytes_gathered = 0

    def gather(self, image_ptr):
        """image_ptr: this rank's row-major channel (device pointer, or a CPU uint8 tensor's data_ptr with
        test ops). After the call rank 0's image holds every rank's tiles."""
        ws, rk = self.world_size, self.rank
        if ws == 1:
            return
        if rk != 0:
            self.ops.pack(self.packed, image_ptr, self.width, self.height, self.bpp, ws, rk)
        dist.gather(self.packed, self.recv, dst=0, group=self.group)
        if self.packed.is_cuda and not getattr(self.ops, "shared_stream", False):
            torch.cuda.current_stream().synchronize()  # NCCL ran on torch's stream, unpack runs on the render stream
        if rk == 0:
            for src in range(1, ws):
                self.ops.unpack(image_ptr, self.recv[src], self.width, self.height, self.bpp, ws, src)
            self.bytes_gathered = (ws - 1) * self.shard_bytes
            if hasattr(self.ops, "synchronize"):
                self.ops.synchronize()


class CudaPeerOps:
    """the peer-exchange entry points of libospcuda.so (include/ospcuda.h, peer_exchange.cu)"""

    def __init__(self):
        L = C.CDLL(lib_path("libospcuda.so"), mode=C.RTLD_GLOBAL)
        L.ospcuda_last_error.restype = C.c_char_p
        L.ospcuda_peer_block_alloc.argtypes = [C.POINTER(C.c_void_p)]
        L.ospcuda_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
        L.ospcuda_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.ospcuda_ipc_close.argtypes = [C.c_void_p]
        L.ospcuda_free.argtypes = [C.c_void_p]
        L.ospcuda_peer_post_ready.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32]
        L.ospcuda_peer_wait_arrive.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.ospcuda_peer_push_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                              C.c_void_p, C.c_uint32]
        L.ospcuda_peer_error.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        self.L = L

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"libospcuda: {self.L.ospcuda_last_error().decode()}")

    def alloc_block(self):
        p = C.c_void_p()
        self._check(self.L.ospcuda_peer_block_alloc(C.byref(p)))
        return p.value

    def free_block(self, ptr):
        self._check(self.L.ospcuda_free(ptr))

    def export(self, ptr):
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self.L.ospcuda_ipc_export(ptr, buf))
        return buf.raw

    def open(self, handle):
        p = C.c_void_p()
        self._check(self.L.ospcuda_ipc_open(handle, C.byref(p)))
        return p.value

    def close(self, ptr):
        self._check(self.L.ospcuda_ipc_close(ptr))

    def post_ready(self, blocks, epoch):
        arr = (C.c_void_p * len(blocks))(*blocks)
        self._check(self.L.ospcuda_peer_post_ready(arr, len(blocks), epoch))

    def wait_arrive(self, root_block, world_size, epoch):
        self._check(self.L.ospcuda_peer_wait_arrive(root_block, world_size, epoch))

    def push(self, peer_image, image_ptr, width, height, bpp, world_size, rank, local_block, root_block, epoch):
        self._check(self.L.ospcuda_peer_push_tiles(peer_image, image_ptr, width, height, bpp, world_size, rank, local_block, root_block, epoch))

    def timed_out(self, block):
        v = C.c_uint32(0)
        self._check(self.L.ospcuda_peer_error(block, C.byref(v)))
        return bool(v.value)


IPC_HANDLE_BYTES = 64


class PeerSetupError(RuntimeError):
    """raised on EVERY rank when any rank could not export or map the peer memory"""


class PeerTileExchange:
    """Composes the display channel on rank 0 by peer stores: every rank r != 0 writes the tiles it owns straight into
    rank 0's row-major image (mapped through CUDA IPC) from one kernel; no staging buffer, no collective call per frame.
    The process group is used once, to exchange the IPC handles. Per frame:

        x.begin_frame()          # rank 0: tells the peers the previous image has been consumed (stream order)
        ... render + colour conversion on the render stream ...
        x.gather(image_ptr)      # rank r: push tiles + arrival flag;  rank 0: wait for the N-1 arrival flags

    Everything is enqueued on the render stream; nothing blocks the host. `image_ptr` on rank 0 must be the pointer
    handed to the constructor (the framebuffer's display buffer: the base of its own cudaMalloc allocation)."""

    def __init__(self, width, height, bpp, image_ptr, ops=None, group=None):
        self.width, self.height, self.bpp = width, height, bpp
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ops = ops if ops is not None else CudaPeerOps()
        self.epoch = 0
        self._open = False  # begin_frame() called, gather() not yet
        self.image_ptr = image_ptr
        self.bytes_gathered = 0
        ws, rk = self.world_size, self.rank
        if ws == 1:
            return
        tx, ty = tile_grid(width, height)
        if ws > tx * ty or ws > 16:
            raise ValueError("PeerTileExchange: more ranks than tiles (or than OSPCUDA_MAX_PEERS)")
        # Set-up is collective and fails collectively: a rank that cannot export or map (no peer access between two GPUs,
        # IPC not permitted in the container) must not leave the others waiting in a collective call, so every step is
        # followed by an exchange of its outcome and all ranks raise PeerSetupError together (bench.py then takes the NCCL
        # gather on every rank).
        self.block, mine = None, {"block": None, "image": None, "error": None}
        try:
            self.block = self.ops.alloc_block()
            mine["block"] = self.ops.export(self.block)
            mine["image"] = self.ops.export(image_ptr) if rk == 0 else None
        except Exception as e:  # noqa: BLE001 -- reported to every rank below
            mine["error"] = f"rank {rk}: {e}"
        everyone = [None] * ws
        dist.all_gather_object(everyone, mine, group=group)
        errors = [m["error"] for m in everyone if m["error"]]
        opened = []
        if not errors:
            try:
                if rk == 0:
                    self.peer_blocks = [self.block]
                    for r in range(1, ws):
                        self.peer_blocks.append(self.ops.open(everyone[r]["block"]))
                        opened.append(self.peer_blocks[-1])
                    self.root_image, self.root_block = image_ptr, self.block
                else:
                    self.peer_blocks = None
                    self.root_image = self.ops.open(everyone[0]["image"])
                    opened.append(self.root_image)
                    self.root_block = self.ops.open(everyone[0]["block"])
                    opened.append(self.root_block)
                mapped = None
            except Exception as e:  # noqa: BLE001
                mapped = f"rank {rk}: {e}"
            outcomes = [None] * ws
            dist.all_gather_object(outcomes, mapped, group=group)
            errors = [o for o in outcomes if o]
        if errors:
            for ptr in opened:
                try:
                    self.ops.close(ptr)
                except Exception:  # noqa: BLE001
                    pass
            if self.block is not None:
                self.ops.free_block(self.block)
            self.world_size = 1  # close() / gather() become no-ops
            raise PeerSetupError("; ".join(errors))
        self.shard_bytes = [len(owned_tiles(width, height, ws, r)) * TILE * TILE * bpp for r in range(ws)]
        dist.barrier(group=group)  # every mapping exists before the first frame

    def begin_frame(self):
        if self.world_size == 1:
            return
        if self._open:
            return
        self.epoch += 1
        self._open = True
        if self.rank == 0:
            self.ops.post_ready(self.peer_blocks, self.epoch)

    def gather(self, image_ptr):
        if self.world_size == 1:
            return
        self.begin_frame()  # no-op when the caller opened the frame before rendering (the overlapping form)
        self._open = False
        if self.rank == 0:
            if image_ptr != self.image_ptr:
                raise ValueError("PeerTileExchange.gather: rank 0 must pass the image it exported")
            self.ops.wait_arrive(self.root_block, self.world_size, self.epoch)
            self.bytes_gathered = sum(self.shard_bytes[1:])
        else:
            self.ops.push(self.root_image, image_ptr, self.width, self.height, self.bpp, self.world_size, self.rank, self.block,
                          self.root_block, self.epoch)

    def check(self):
        """synchronous: raises if a device-side wait of this rank timed out (a peer never showed up)"""
        if self.world_size > 1 and self.ops.timed_out(self.block):
            raise RuntimeError(f"PeerTileExchange: rank {self.rank} timed out waiting for a peer")

    def close(self):
        if self.world_size == 1:
            return
        dist.barrier(group=self.group)  # nobody unmaps while a peer may still be writing
        if self.rank == 0:
            for p in self.peer_blocks[1:]:
                self.ops.close(p)
        else:
            self.ops.close(self.root_image)
            self.ops.close(self.root_block)
        dist.barrier(group=self.group)
        self.ops.free_block(self.block)
        self.world_size = 1


def tile_checksums(img, world_size):
    """per owner rank: (sum of all bytes, sum of bytes * position) over the pixels of the tiles that rank owns, of an
    (H, W, C) uint8 image"""
    import numpy as np
    h, w = img.shape[:2]
    tx, _ = tile_grid(w, h)
    owner = ((np.arange(h)[:, None] // TILE) * tx + (np.arange(w)[None, :] // TILE)) % world_size
    flat = img.reshape(h, w, -1).astype(np.uint64).sum(axis=2)
    pos = (np.arange(h, dtype=np.uint64)[:, None] * np.uint64(w) + np.arange(w, dtype=np.uint64)[None, :]) % np.uint64(65521) + np.uint64(1)
    return [(int(flat[owner == r].sum()), int((flat * pos)[owner == r].sum())) for r in range(world_size)]


def verify_composition(local_img, group=None):
    """Every rank passes its own mapped display image (rank 0: the composed one). Checks on rank 0 that the tiles of
    every rank r in the composed image carry exactly what rank r rendered (checksums of r's owned tiles, compared
    through the process group). Returns True on rank 0 when all match (raises otherwise), None elsewhere."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rk = dist.get_rank(group) if dist.is_initialized() else 0
    sums = tile_checksums(local_img, ws)
    everyone = [None] * ws
    if ws > 1:
        dist.all_gather_object(everyone, sums[rk], group=group)
    else:
        everyone[0] = sums[0]
    if rk != 0:
        return None
    bad = [r for r in range(ws) if tuple(everyone[r]) != tuple(sums[r])]
    if bad:
        raise RuntimeError(f"tile composition mismatch: GPU 0's image does not hold the tiles rendered by rank(s) {bad}")
    return True


def init_process_group_from_env():
    """torchrun environment -> (rank, world_size, local_rank); NCCL when CUDA is present, else gloo
    (OSPRAY_B200_DIST_BACKEND=gloo forces gloo: ranks that share one GPU, where NCCL refuses to run)"""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        # NCCL writes its version / debug lines to stdout by default; bench.py's stdout must stay ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = os.environ.get("OSPRAY_B200_DIST_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world, local
