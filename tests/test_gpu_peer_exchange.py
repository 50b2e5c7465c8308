"""PeerTileExchange on real CUDA IPC memory: two processes (sub-devices 0 and 1 of a 2-way tile partition) share cuda:0,
rank 1 pushes its tiles into rank 0's display buffer with ospcuda_peer_push_tiles, the composed sRGBA8 frames equal the
single-device frames bit for bit (MultiDevice.cpp:395-418 partition + LocalFB.cpp:286-353 composition). On the 8-GPU box
the same code runs with one GPU per rank (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAMES = 3


def _render_setup(osp, w, h):
    from ospray_b200 import testing as T
    world, keep = T.to_ospray(T.cornell_box())
    cam, pose = T.make_camera(world, w, h)
    ren = T.make_renderer("pathtracer", pixelSamples=1)
    fb = osp.FrameBuffer(w, h, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM)
    return world, keep, cam, ren, fb


def _worker(rank, world_size, port, w, h, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK="0",
                      OSPRAY_B200_DIST_BACKEND="gloo")
    import torch.distributed as dist
    from ospray_b200 import capi as osp, multigpu
    multigpu.init_process_group_from_env()
    osp.init(cudaDevice=0, numSubdevices=world_size, subdeviceRank=rank)
    world, keep, cam, ren, fb = _render_setup(osp, w, h)
    ptr, bpp = fb.device_pointer(osp.OSP_FB_COLOR)
    x = multigpu.PeerTileExchange(w, h, bpp, ptr)
    frames = []
    for _ in range(FRAMES):
        x.begin_frame()
        f = fb.render_frame(ren, cam, world)
        x.gather(ptr)
        f.wait()
        if rank == 0:
            frames.append(fb.map(osp.OSP_FB_COLOR).copy())
    x.check()
    if rank == 0:
        np.savez(out_path, frames=np.stack(frames), gathered=x.bytes_gathered)
    x.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world_size", [2, 3])
def test_peer_push_composes_the_single_device_frame(osp, tmp_path, world_size):
    w, h = 328, 200  # 6 x 4 tiles, ragged right/top border
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "peer_gpu.npz")
    mp.spawn(_worker, args=(world_size, port, w, h, out), nprocs=world_size, join=True)
    got = np.load(out)
    world, keep, cam, ren, fb = _render_setup(osp, w, h)
    for i in range(FRAMES):
        fb.render_frame(ren, cam, world).wait()
        np.testing.assert_array_equal(got["frames"][i], fb.map(osp.OSP_FB_COLOR))
    from ospray_b200 import multigpu
    tiles = multigpu.tile_grid(w, h)[0] * multigpu.tile_grid(w, h)[1]
    assert int(got["gathered"]) == (tiles - len(multigpu.owned_tiles(w, h, world_size, 0))) * 64 * 64 * 4
