#!/usr/bin/env python
"""Concurrent device-to-host bandwidth of N ranks (pinned 100 MB buffers), to tell a platform ceiling
from a library problem:  python -m torch.distributed.run --nproc-per-node N ... tools/d2h_bandwidth.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
src = torch.empty(100_000_000, dtype=torch.uint8, device=dev)
dst = torch.empty(100_000_000, dtype=torch.uint8, pin_memory=True)


def measure(active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    return 20 * 0.1 / dt if active else 0.0


measure(True)
alone = measure(rank == 0)
together = measure(True)
t = torch.tensor([together], device=dev)
if world > 1:
    dist.all_reduce(t)
if rank == 0:
    print("rank 0 alone: %.1f GB/s; all %d ranks together: %.1f GB/s per rank (rank 0), %.1f GB/s aggregate"
          % (alone, world, together, float(t.item())))
if world > 1:
    dist.destroy_process_group()
