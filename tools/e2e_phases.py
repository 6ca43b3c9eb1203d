#!/usr/bin/env python
"""Host-side timeline of the end-to-end multi-epoch call (bench.py's e2e leg), phase by phase."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rvmc_b200.bias_corr import BiasCorr  # noqa: E402

DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0
dev = torch.device("cuda", 0)
nstars, ns = 100, 10 ** 6
t_obs = [DAYS6.copy() for _ in range(nstars)]
for chunks in (8, 12, 16, 24):
    BiasCorr._d2h_chunks = chunks
    acc = np.zeros(4)
    n = 12
    for k in range(n + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bc = BiasCorr(t_obs, number_of_samples=ns, rv_errors=2.0, seed=k, device=dev)
        t1 = time.perf_counter()
        bc.calculate_radial_velocities()
        t2 = time.perf_counter()
        d = bc.get_detected_binaries()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if k >= 2:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t0]
    acc *= 1e3 / n
    print("chunks %2d: ctor %.3f ms, calculate (launches) %.3f ms, get_detected (wait+D2H) %.3f ms, total %.3f ms -> %.3e RV/s"
          % (chunks, acc[0], acc[1], acc[2], acc[3], nstars * ns * 6 / (acc[3] * 1e-3)))

# a single star with many samples: the pass is split along the sample axis
for chunks in (1, 12):
    BiasCorr._d2h_chunks = chunks
    acc, n = 0.0, 8
    for k in range(n + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bc = BiasCorr([DAYS6], number_of_samples=10 ** 8, rv_errors=2.0, seed=k, device=dev)
        bc.calculate_radial_velocities()
        d = bc.get_detected_binaries()
        torch.cuda.synchronize()
        if k >= 2:
            acc += time.perf_counter() - t0
    print("1 star x 1e8 samples, %2d range(s): %.3f ms -> %.3e RV/s" % (chunks, acc / n * 1e3, 6e8 / (acc / n)))
