#!/usr/bin/env python3
"""Bare host->device copy roof: cudaMemcpyAsync of one bench capture (402 MB) from pinned memory, nothing else on the
GPU.  Run alone (N=1) or under torchrun (one process per GPU) to see what the box gives N ranks at once; the bench's
e2e legs are judged against this number."""
import os
import sys
import time

import torch

rank = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
nbytes = int(float(os.environ.get("ROOF_MB", "402")) * 1e6)
reps = 20
for label, nchunk, nstream in (("one copy", 1, 1), ("16 chunks, 1 stream", 16, 1), ("16 chunks, 2 streams", 16, 2), ("2 buffers, 2 streams", 2, 2)):
    total = nbytes * (2 if label.startswith("2 buffers") else 1)
    h = torch.empty(total, dtype=torch.uint8).pin_memory()
    d = torch.empty(total, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(nstream)]
    step = total // nchunk
    def once():
        for c in range(nchunk):
            with torch.cuda.stream(streams[c % nstream]):
                d[c * step:(c + 1) * step].copy_(h[c * step:(c + 1) * step], non_blocking=True)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = total * reps / dt / 1e9
    if world > 1:
        t = torch.tensor([gbs], device="cuda")
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(t)
        if rank == 0:
            print("h2d roof, %d ranks, %s: aggregate %.1f GB/s, slowest rank %.1f GB/s" % (world, label, float(t), float(lo)), flush=True)
    else:
        print("h2d roof, 1 rank, %s: %.1f GB/s" % (label, gbs), flush=True)
    del h, d
# device -> host of one TS (19.6 MB)
h = torch.empty(19_643_744, dtype=torch.uint8).pin_memory()
d = torch.empty(19_643_744, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
if rank == 0:
    print("d2h 19.6 MB: %.1f GB/s" % (h.numel() * reps / (time.perf_counter() - t0) / 1e9), flush=True)
if world > 1:
    dist.destroy_process_group()
