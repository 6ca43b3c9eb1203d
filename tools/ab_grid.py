#!/usr/bin/env python
"""A/B timing of the grid kernels for an experimental build (RVMC_B200_LIB=build/variants/...): the one-pass C5
grid (best of 5), the independent-cells grid (best of 2) and one C3-shaped launch of the fused sigma kernel.
    RVMC_B200_LIB=$PWD/build/variants/librvmc_b200_v1.so python tools/ab_grid.py v1"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rvmc_b200 import _abi, grid  # noqa: E402
from rvmc_b200 import _device as dv  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
axes = dict(period_pi=np.linspace(-0.95, 0.5, 64), mass_ratio_kappa=np.linspace(-0.95, 1.0, 64),
            fbin=np.linspace(0.0625, 1, 16))
grid.grid_search({k: v[:4] for k, v in axes.items()}, number_of_samples=1000, seed=1, device=dev)


def best(reps, **kw):
    grid.grid_search(axes, number_of_samples=10 ** 5, seed=2023, observed_sigma=25.0, device=dev, **kw)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = grid.grid_search(axes, number_of_samples=10 ** 5, seed=2024, observed_sigma=25.0, device=dev, **kw)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


t_fused, r = best(5)
t_indep, ri = best(2, share_fbin_draws=False) if os.environ.get("AB_INDEP") else (float("nan"), r)
lib = _abi.lib()
m17 = dv.to_device(np.array([0.8, 2.3, 1.0, 1.8, 0.8, 1.0, 1.4, 2.5, 2.0, 0.4, 1.5, 0.6]), "float32", dev)
n_real = int(np.ceil(1e9 / (12 * 0.7)))
sig = dv.empty((n_real,), "float32", dev)
pop = _abi.PopParams.defaults(min_period=1.4, fbin=0.7)
st = dv.stream_ptr(dev)


def c3():
    _abi.check(lib.rvmc_sigma1d_fused(C.byref(pop), 1, dv.ptr(m17), 12, 0, 1, 0, n_real, dv.ptr(sig), None, st))


c3()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    c3()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"variant": sys.argv[1] if len(sys.argv) > 1 else "?", "lib": os.path.basename(_abi.LIB_PATH),
                  "grid_fused_s": t_fused, "grid_indep_s": t_indep, "c3_ms": e0.elapsed_time(e1) / 3,
                  "best_fused": list(r["best_index"]), "best_indep": list(ri["best_index"]),
                  "median_sum": float(r["stats"]["median"].sum()), "mode_sum": float(np.nansum(r["stats"]["mode"])),
                  "c3_median": float(torch.median(sig).item())}))
