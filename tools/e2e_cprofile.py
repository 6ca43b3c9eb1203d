#!/usr/bin/env python
"""cProfile of the host side of BiasCorr.calculate_radial_velocities (time to first launch matters)."""
import cProfile
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rvmc_b200.bias_corr import BiasCorr  # noqa: E402

DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0
dev = torch.device("cuda", 0)
t_obs = [DAYS6.copy() for _ in range(100)]
objs = [BiasCorr(t_obs, number_of_samples=10 ** 6, rv_errors=2.0, seed=k, device=dev) for k in range(22)]
for b in objs[:2]:
    b.calculate_radial_velocities()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for b in objs[2:]:
    b.calculate_radial_velocities()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
