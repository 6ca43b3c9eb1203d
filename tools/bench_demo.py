#!/usr/bin/env python
"""The reference's own timed demo (`RVMC.py:368-396`, its only timing print) on the drop-in class:
RVMC(number_of_stars=1e6); calculate_binary_velocities(); subsample(fbin=f, number_of_samples=1e6)
for 11 binary fractions (finite-pool bootstrap, sigma_1d returned as a host array each time)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rvmc_b200 import RVMC  # noqa: E402

dev = torch.device("cuda", 0)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rv = RVMC(number_of_stars=10 ** 6, seed=rep, device=dev)
    rv.calculate_binary_velocities()
    _ = rv.rv_binary          # the reference materialises it; read it on the host as a script would
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    fbins = np.linspace(0, 1, 11)
    dists = [rv.subsample(fbin=fbin, number_of_samples=10 ** 6) for fbin in fbins]
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    med = [float(np.median(d)) for d in dists]
    print(json.dumps({"case": "RVMC.py __main__ demo", "rep": rep, "generate_s": t1 - t0, "subsample_11_s": t2 - t1,
                      "star_draws_per_s": 11 * 12 * 10 ** 6 / (t2 - t1), "median_sigma_fbin0": med[0],
                      "median_sigma_fbin1": med[-1]}))
