"""N>1 host logic on CPU: two gloo ranks partition the image tiles, render their share with the oracle and
gather the display channel to rank 0 through ospray_b200.multigpu.TileGather (the same class bench.py drives
over NCCL); the staging kernels are replaced by the oracle's CPU restatement of the same layout."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleTileOps:
    """CPU stand-in for CudaTileOps (checker only)"""

    def pack(self, packed, image_ptr, width, height, bpp, world_size, rank):
        from oracle import pyoracle as O
        O.lib().oracle_fb_pack_tiles(packed.data_ptr(), image_ptr, width, height, bpp, world_size, rank, 1)

    def unpack(self, image_ptr, packed, width, height, bpp, world_size, src_rank):
        from oracle import pyoracle as O
        O.lib().oracle_fb_pack_tiles(packed.data_ptr(), image_ptr, width, height, bpp, world_size, src_rank, 0)


def _worker(rank, world, port, w, h, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    from oracle import pyoracle as O
    O.lib().oracle_fb_pack_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    import scenes
    from ospray_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, ws, _ = multigpu.init_process_group_from_env()
    assert (r, ws) == (rank, world) and dist.get_backend() == "gloo"
    fb = scenes.oracle_render("cornell_pt", w, h, num_ranks=world, rank=rank, threads=1)
    srgb = O.color_convert(fb.buf["color"], 2)  # per-rank colour conversion BEFORE the gather: 4 B/px travels
    img = torch.from_numpy(srgb.view(np.uint8).copy())
    g = multigpu.TileGather(w, h, 4, "cpu", ops=OracleTileOps())
    assert g.shard_bytes == multigpu.max_owned(w, h, world) * 64 * 64 * 4
    g.gather(img.data_ptr())
    # ray statistics are reduced like the reference reduces nothing else: a plain sum
    cnt = torch.from_numpy(fb.counters.astype(np.int64))
    dist.reduce(cnt, dst=0)
    if rank == 0:
        np.savez(out_path, img=img.numpy(), rays=cnt.numpy(), gathered=g.bytes_gathered)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_tile_gather_gloo(tmp_path, world):
    import scenes
    from oracle import pyoracle as O
    from ospray_b200 import multigpu
    w, h = 200, 136  # 4 x 3 tiles, ragged right/top border
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), w, h, out), nprocs=world, join=True)
    got = np.load(out)
    full = scenes.oracle_render("cornell_pt", w, h, threads=1)
    want = O.color_convert(full.buf["color"], 2).view(np.uint8)
    np.testing.assert_array_equal(got["img"], want)
    assert int(got["rays"].sum()) == int(full.counters.sum())
    assert int(got["gathered"]) == (world - 1) * multigpu.max_owned(w, h, world) * 64 * 64 * 4


def test_tile_ownership_is_a_partition():
    from ospray_b200 import multigpu
    for (w, h) in [(1024, 1024), (3840, 2160), (65, 1), (64, 64)]:
        tx, ty = multigpu.tile_grid(w, h)
        for n in (1, 2, 4, 8):
            sets = [multigpu.owned_tiles(w, h, n, r) for r in range(n)]
            flat = sorted(t for s in sets for t in s)
            assert flat == list(range(tx * ty))
            assert max(len(s) for s in sets) == multigpu.max_owned(w, h, n)
            assert max(len(s) for s in sets) - min(len(s) for s in sets) <= 1


def test_cuda_tile_ops_refuse_cpu_tensors():
    from ospray_b200 import multigpu
    ops = multigpu.CudaTileOps()
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.pack(torch.zeros(16, dtype=torch.uint8), 0, 64, 64, 4, 2, 1)


# ---- PeerTileExchange: the handle exchange, epochs and the ready/arrive handshake, on CPU ------------------------------
class ShmPeerOps:
    """stands in for CudaPeerOps: "device memory" = POSIX shared memory, an IPC handle = the segment's name, the push
    kernel = a numpy tile copy into the mapped image. Same call sequence and flag protocol as peer_exchange.cu."""
    READY, ARRIVE = 0, 8

    def __init__(self):
        from multiprocessing import shared_memory
        self.shm_mod, self.segs = shared_memory, {}

    def _new(self, nbytes):
        seg = self.shm_mod.SharedMemory(create=True, size=nbytes)
        self.segs[seg.name] = seg
        np.ndarray(nbytes, np.uint8, seg.buf)[:] = 0
        return seg.name

    def _arr(self, name, dtype):
        seg = self.segs[name]
        return np.ndarray(seg.size // np.dtype(dtype).itemsize, dtype, seg.buf)

    def alloc_block(self): return self._new(256)
    def alloc_image(self, nbytes): return self._new(nbytes)
    def export(self, ptr): return ptr.encode()

    def open(self, handle):
        name = handle.decode()
        if name not in self.segs:
            self.segs[name] = self.shm_mod.SharedMemory(name=name)
        return name

    def close(self, ptr): pass

    def free_block(self, ptr): pass

    def _spin(self, get, epoch):
        import time
        t0 = time.time()
        while get() < epoch:
            time.sleep(0.001)
            assert time.time() - t0 < 60, "peer never showed up"

    def post_ready(self, blocks, epoch):
        for b in blocks[1:]:
            self._arr(b, np.uint32)[self.READY] = epoch

    def wait_arrive(self, root_block, world_size, epoch):
        a = self._arr(root_block, np.uint32)
        for r in range(1, world_size):
            self._spin(lambda: a[self.ARRIVE + r], epoch)

    def push(self, peer_image, image, width, height, bpp, world_size, rank, local_block, root_block, epoch):
        from ospray_b200 import multigpu
        self._spin(lambda: self._arr(local_block, np.uint32)[self.READY], epoch)
        dst = self._arr(peer_image, np.uint8)[: width * height * bpp].reshape(height, width, bpp)
        tx, _ = multigpu.tile_grid(width, height)
        for t in multigpu.owned_tiles(width, height, world_size, rank):
            x0, y0 = (t % tx) * 64, (t // tx) * 64
            dst[y0:y0 + 64, x0:x0 + 64] = image[y0:y0 + 64, x0:x0 + 64]
        self._arr(root_block, np.uint32)[self.ARRIVE + rank] = epoch

    def timed_out(self, block): return False

    def unlink(self):
        for seg in self.segs.values():
            seg.close()


def _peer_worker(rank, world, port, w, h, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import pyoracle as O
    import scenes
    from ospray_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    multigpu.init_process_group_from_env()
    ops = ShmPeerOps()
    root_image = ops.alloc_image(w * h * 4) if rank == 0 else None
    x = multigpu.PeerTileExchange(w, h, 4, root_image, ops=ops)
    frames = []
    for frame in range(2):  # two epochs: the second frame may only be pushed after rank 0 has consumed the first
        x.begin_frame()
        fb = scenes.oracle_render("cornell_pt", w, h, num_ranks=world, rank=rank, threads=1, frames=frame + 1)
        srgb = O.color_convert(fb.buf["color"], 2).view(np.uint8).reshape(h, w, 4).copy()
        if rank == 0:
            img = ops._arr(root_image, np.uint8).reshape(h, w, 4)
            tx, _ = multigpu.tile_grid(w, h)
            for t in multigpu.owned_tiles(w, h, world, 0):
                x0, y0 = (t % tx) * 64, (t // tx) * 64
                img[y0:y0 + 64, x0:x0 + 64] = srgb[y0:y0 + 64, x0:x0 + 64]
            x.gather(root_image)
            frames.append(img.copy())  # "ospMapFrameBuffer": consumed before the next begin_frame
        else:
            x.gather(srgb)
    assert x.epoch == 2
    x.check()
    assert multigpu.verify_composition(frames[-1] if rank == 0 else srgb) is (True if rank == 0 else None)
    if rank == 0:
        np.savez(out_path, f0=frames[0], f1=frames[1], gathered=x.bytes_gathered)
    x.close()
    ops.unlink()
    if rank == 0:
        for name in (root_image,):
            ops.segs[name].unlink()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_tile_exchange_protocol_gloo(tmp_path, world):
    """PeerTileExchange with shared memory standing in for NVLink peer memory: IPC handle exchange through the process
    group, per-frame epochs, ready/arrive flags; the composed frames equal the single-process frames bit for bit"""
    import scenes
    from oracle import pyoracle as O
    from ospray_b200 import multigpu
    w, h = 200, 136
    out = str(tmp_path / "peer.npz")
    mp.spawn(_peer_worker, args=(world, _free_port(), w, h, out), nprocs=world, join=True)
    got = np.load(out)
    for k, frames in (("f0", 1), ("f1", 2)):
        full = scenes.oracle_render("cornell_pt", w, h, threads=1, frames=frames)
        want = O.color_convert(full.buf["color"], 2).view(np.uint8).reshape(h, w, 4)
        np.testing.assert_array_equal(got[k], want)
    tiles = multigpu.tile_grid(w, h)[0] * multigpu.tile_grid(w, h)[1]
    assert int(got["gathered"]) == (tiles - len(multigpu.owned_tiles(w, h, world, 0))) * 64 * 64 * 4


def test_tile_checksums_see_a_missing_tile():
    from ospray_b200 import multigpu
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (136, 200, 4), dtype=np.uint8)
    a = multigpu.tile_checksums(img, 3)
    broken = img.copy()
    broken[64:128, 64:128] = 0          # tile 5 of a 4 x 3 grid: owned by rank 5 % 3 == 2
    b = multigpu.tile_checksums(broken, 3)
    assert a[0] == b[0] and a[1] == b[1] and a[2] != b[2]
    swapped = img.copy()
    swapped[0:64, 0:64], swapped[64:128, 128:192] = img[64:128, 128:192], img[0:64, 0:64]   # tiles 0 and 6, both rank 0
    assert multigpu.tile_checksums(swapped, 3)[0] != a[0]                                   # position-weighted sum


class _FailingOpenOps(ShmPeerOps):
    """rank 1 cannot map rank 0's memory (no peer access / IPC not permitted)"""

    def __init__(self, rank):
        super().__init__()
        self.rank = rank

    def open(self, handle):
        if self.rank == 1:
            raise RuntimeError("peer access is not supported between these two devices")
        return super().open(handle)


def _peer_failure_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from ospray_b200 import multigpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    multigpu.init_process_group_from_env()
    ops = _FailingOpenOps(rank)
    img = ops.alloc_image(128 * 128 * 4) if rank == 0 else None
    msg = "no error"
    try:
        multigpu.PeerTileExchange(128, 128, 4, img, ops=ops)
    except multigpu.PeerSetupError as e:
        msg = str(e)
    dist.barrier()  # every rank got here: nobody is stuck in a collective of the failed set-up
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write(msg)
    ops.unlink()
    if rank == 0:
        ops.segs[img].unlink()
    dist.destroy_process_group()


def test_peer_setup_failure_is_raised_on_every_rank(tmp_path):
    """one rank failing to map the peer memory makes ALL ranks raise PeerSetupError (so that all of them fall back to the NCCL
    gather together) instead of leaving the others waiting in a collective call"""
    mp.spawn(_peer_failure_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    msgs = [open(tmp_path / f"rank{r}.txt").read() for r in range(3)]
    assert all("rank 1: peer access is not supported" in m for m in msgs), msgs
