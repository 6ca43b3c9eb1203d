#!/usr/bin/env python
"""Where is the wall of the end-to-end detection pass: the PCIe/IOMMU path of the device-to-host copy,
or the host's own memory system?  Measured with N concurrent ranks (one per GPU), 10^8 flags per rank:

  bytes   : D2H of 100 MB of uint8 flags into pinned memory (what round 1 shipped)
  bits    : D2H of 12.5 MB of packed flags into pinned memory
  unpack  : rvmc_unpack_bits_host of those 12.5 MB into a 100 MB pinned bool array, T threads per rank
  both    : bits + unpack back to back (the cost of the packed route per pass)
  hybridXX: XX % of the flags as bytes by DMA WHILE the CPU widens the packed rest (are the two walls independent?)

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tools/d2h_bandwidth.py

Prints one JSON line (rank 0): per-rank GB/s of RESULT bytes (100 MB per pass) for each route, alone
(rank 0 only) and with all ranks active."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rvmc_b200 import _abi  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lib = _abi.lib()
N = 100_000_000
NW = N // 32
src_bytes = torch.randint(0, 2, (N,), dtype=torch.uint8, device=dev)
src_bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (NW,), dtype=torch.int32, device=dev)
dst_bytes = torch.empty(N, dtype=torch.uint8, pin_memory=True)
dst_bits = torch.empty(NW, dtype=torch.int32, pin_memory=True)
REPS = 10


def run(route, threads):
    if route.startswith("hybrid"):
        # a share of the flags as bytes by DMA, the rest as bits widened by the CPU -- both at the same time:
        # are the two walls (PCIe/IOMMU path of the DMA, memory system under CPU stores) independent?
        share = int(route[6:]) / 100.0
        nb = int(N * share) // 4096 * 4096          # flags that travel as bytes
        nw0 = nb // 32
        for _ in range(REPS):
            dst_bits[nw0:].copy_(src_bits[nw0:], non_blocking=True)        # the packed words first (small) ...
            e_bits = torch.cuda.Event()
            e_bits.record()
            dst_bytes[:nb].copy_(src_bytes[:nb], non_blocking=True)        # ... then the byte block by DMA
            e_bits.synchronize()
            assert lib.rvmc_unpack_bits_host(dst_bits.data_ptr() + 4 * nw0, NW - nw0, 1, N - nb,
                                             dst_bytes.data_ptr() + nb, N - nb, threads) == 0   # while the DMA runs
            torch.cuda.synchronize()
        return
    for _ in range(REPS):
        if route == "bytes":
            dst_bytes.copy_(src_bytes, non_blocking=True)
            torch.cuda.synchronize()
        if route in ("bits", "both"):
            dst_bits.copy_(src_bits, non_blocking=True)
            torch.cuda.synchronize()
        if route in ("unpack", "both"):
            assert lib.rvmc_unpack_bits_host(dst_bits.data_ptr(), NW, 1, N, dst_bytes.data_ptr(), N, threads) == 0


def measure(route, threads, active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if active:
        run(route, threads)
    dt = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    return REPS * N / dt / 1e9 if active else 0.0


def agg(x):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    return float(t.item())


cores = len(os.sched_getaffinity(0))
out = {"world": world, "host_cores": cores, "flags_per_pass": N, "unit": "GB/s of result bytes per rank"}
run("both", 2)          # warm-up: pins, pool threads, page faults
for route, threads in (("bytes", 0), ("bits", 0), ("unpack", 1), ("unpack", 2), ("unpack", 4), ("unpack", 8),
                       ("both", 2), ("both", 4), ("both", 8), ("hybrid30", 2), ("hybrid30", 4), ("hybrid40", 2),
                       ("hybrid40", 4), ("hybrid50", 2), ("hybrid50", 4)):
    if threads and threads * world > 2 * cores:
        continue
    alone = measure(route, threads, rank == 0)
    together = measure(route, threads, True)
    total = agg(together)
    if rank == 0:
        out["%s%s" % (route, "_t%d" % threads if threads else "")] = {
            "alone": round(alone, 2), "together_rank0": round(together, 2), "aggregate": round(total, 2)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
