#!/usr/bin/env python
"""BASELINE config 5: the 64 x 64 x 16 (pi, kappa, fbin) grid x 10^5 realizations of 12 stars,
grid points dealt round-robin to the ranks, per-point statistics all-gathered (strong scaling).
    python tools/bench_grid.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/bench_grid.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rvmc_b200 import grid  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = [int(x) for x in os.environ.get("RVMC_GRID_SHAPE", "64,64,16").split(",")]
    axes = dict(period_pi=np.linspace(-0.95, 0.5, n[0]), mass_ratio_kappa=np.linspace(-0.95, 1.0, n[1]),
                fbin=np.linspace(0.0625, 1, n[2]))
    n_real = int(os.environ.get("RVMC_GRID_NREAL", "100000"))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    grid.grid_search({k: v[:4] for k, v in axes.items()}, number_of_samples=1000, seed=1)      # warm-up
    for share in (True, False):
        times = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            r = grid.grid_search(axes, number_of_samples=n_real, seed=2024, observed_sigma=25.0, share_fbin_draws=share)
            barrier()
            times.append(time.perf_counter() - t0)
        cnt = torch.from_numpy(r["counters"].astype(np.int64)).to(dev)
        if world > 1:
            dist.all_reduce(cnt)
        if rank == 0:
            npts = int(np.prod(r["shape"]))
            dt = min(times)
            print(json.dumps({"case": "C5 grid %dx%dx%d x %d realizations x 12 stars" % (n[0], n[1], n[2], n_real),
                              "fbin_axis": "one pass, draws shared along f_bin (rvmc_grid_run_fbins)" if share else
                              "independent draws per cell (rvmc_grid_run)",
                              "n_gpus": world, "points": npts, "seconds_best_of_3": dt, "seconds_all": times,
                              "points_per_s": npts / dt, "binary_rvs_solved_per_s": float(cnt[0]) / dt,
                              "scaling": "strong", "best_index": r["best_index"], "best_params": r["best_params"],
                              "median_range": [float(r["stats"]["median"].min()), float(r["stats"]["median"].max())]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
