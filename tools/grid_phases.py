#!/usr/bin/env python
"""Where does a sharded grid_search spend its wall time?  One rank per GPU under torchrun; every rank
prints the totals of its host-side spans (RVMC_TRACE=sync: each span is bracketed by device syncs, so its
time is the time of the GPU work it issued) and the untraced wall time of a whole call.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/grid_phases.py [--indep]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rvmc_b200 import _trace, grid  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
axes = dict(period_pi=np.linspace(-0.95, 0.5, 64), mass_ratio_kappa=np.linspace(-0.95, 1.0, 64), fbin=np.linspace(0.0625, 1, 16))
kw = dict(number_of_samples=10 ** 5, seed=2024, observed_sigma=25.0, device=dev, share_fbin_draws="--indep" not in sys.argv)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


grid.grid_search({k: v[:4] for k, v in axes.items()}, number_of_samples=1000, seed=1, device=dev)
grid.grid_search(axes, **kw)
walls = []
for rep in range(3):
    barrier()
    t0 = time.perf_counter()
    grid.grid_search(axes, **kw)
    t_local = time.perf_counter() - t0
    barrier()
    walls.append((time.perf_counter() - t0, t_local))
_trace.enable("sync")
_trace.totals(reset=True)
for rep in range(2):
    barrier()
    grid.grid_search(axes, **kw)
barrier()
tot = {k: round(v[1] / 2 * 1e3, 3) for k, v in _trace.totals().items()}
for r in range(world):
    if r == rank:
        print(json.dumps({"rank": rank, "world": world, "wall_ms_incl_barrier": [round(w[0] * 1e3, 2) for w in walls],
                          "wall_ms_this_rank": [round(w[1] * 1e3, 2) for w in walls], "span_ms_synced": tot}), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
