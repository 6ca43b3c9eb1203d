#!/usr/bin/env python
"""Host-side cost of grid_search on the C5 axes (GPU work made negligible by a tiny realization count)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rvmc_b200 import grid  # noqa: E402

torch.cuda.set_device(0)
axes = dict(period_pi=np.linspace(-0.95, 0.5, 64), mass_ratio_kappa=np.linspace(-0.95, 1.0, 64), fbin=np.linspace(0.0625, 1, 16))
for _ in range(2):
    grid.grid_search(axes, number_of_samples=64, seed=1, observed_sigma=25.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    grid.grid_search(axes, number_of_samples=64, seed=1, observed_sigma=25.0)
torch.cuda.synchronize()
print("grid_search with 64 realizations per point: %.2f ms per call" % ((time.perf_counter() - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    grid.grid_search(axes, number_of_samples=64, seed=1, observed_sigma=25.0)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
