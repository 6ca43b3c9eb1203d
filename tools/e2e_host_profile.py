#!/usr/bin/env python
"""cProfile of BiasCorr(...).calculate_radial_velocities() in the steady state of bench.py's e2e leg
(the GPU is idle when each call starts, results are fetched and the object dropped after each step)."""
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
pr = cProfile.Profile()
for k in range(24):
    bc = BiasCorr(t_obs, number_of_samples=10 ** 6, rv_errors=2.0, seed=k, device=dev)
    if k >= 4:
        pr.enable()
    bc.calculate_radial_velocities()
    pr.disable()
    d = bc.get_detected_binaries()
    del bc, d
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
