#!/usr/bin/env python3
"""Per-rank, per-step device and wall times of the bench.py multi-GPU step under variations (diagnostic).
torchrun --nproc-per-node N tools/debug_multigpu_step.py"""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ospray_b200 import capi as osp, multigpu, testing as T

rank, world, local = multigpu.init_process_group_from_env()
torch.cuda.set_device(local)
w = h = 1024
osp.init(cudaDevice=local, numSubdevices=world, subdeviceRank=rank)
stream = torch.cuda.Stream()
osp.set_render_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    world_obj, keep = T.to_ospray(T.cornell_box())
    cam, pose = T.make_camera(world_obj, w, h)
    ren = T.make_renderer("pathtracer", pixelSamples=1)
    fb = osp.FrameBuffer(w, h, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH)
    gather = multigpu.TileGather(w, h, 4, torch.device("cuda", local), ops=multigpu.CudaTileOps(shared_stream=True)) if world > 1 else None
    conv_ptr, bpp = fb.device_pointer(osp.OSP_FB_COLOR)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run(tag, do_flush, do_counts, do_gather, smi, barrier_each):
        p = None
        if smi and rank == 0:
            p = subprocess.Popen(["nvidia-smi", "-i", str(local), "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "20"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            time.sleep(0.3)
        for _ in range(3):
            f = fb.render_frame(ren, cam, world_obj)
            if gather is not None and do_gather:
                gather.gather(conv_ptr)
            f.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dev, wall = [], []
        for i in range(10):
            if barrier_each and world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            f = fb.render_frame(ren, cam, world_obj)
            if gather is not None and do_gather:
                gather.gather(conv_ptr)
            e1.record(stream)
            f.wait()
            if do_counts:
                fb.ray_counts()
            if do_flush:
                flush.zero_()
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            dev.append(e0.elapsed_time(e1))
        if p:
            p.terminate()
        print(f"[{tag}] rank {rank}: dev ms {np.round(dev, 2).tolist()} wall ms {np.round(wall, 2).tolist()}", flush=True)
        if world > 1:
            dist.barrier()

    run("plain", False, False, True, False, False)
    run("flush", True, False, True, False, False)
    run("flush+counts", True, True, True, False, False)
    run("flush+counts+smi", True, True, True, True, False)
    run("flush+counts+barrier", True, True, True, False, True)
    run("nogather", True, True, False, False, False)
if world > 1:
    dist.destroy_process_group()
