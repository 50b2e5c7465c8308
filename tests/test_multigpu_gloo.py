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
