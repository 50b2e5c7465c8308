"""One-process-per-GPU tile partition and NCCL gather of the finished tiles to GPU 0.

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


def init_process_group_from_env():
    """torchrun environment -> (rank, world_size, local_rank); NCCL when CUDA is present, else gloo"""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        # NCCL writes its version / debug lines to stdout by default; bench.py's stdout must stay ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world, local
