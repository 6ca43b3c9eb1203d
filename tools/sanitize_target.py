#!/usr/bin/env python
"""Tiny launches of every kernel for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rvmc_b200 import _abi, grid  # noqa: E402
from rvmc_b200 import _device as dv  # noqa: E402
from rvmc_b200.bias_corr import BiasCorr  # noqa: E402
from rvmc_b200.rvmc import RVMC  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
M17 = grid.M17_ERRORS
days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0

c = RVMC(number_of_stars=3000, seed=1)
c.calculate_binary_velocities(materialize=False)
for S, w in ((12, False), (5, True), (100, False), (333, True), (3000, False)):
    c.subsample(sample_size=S, measured_errors=np.resize(M17, S), number_of_samples=257, weighted=w)
    c.subsample_fused(sample_size=S, measured_errors=np.resize(M17, S), number_of_samples=257, weighted=w)
c.subsample_fused(sample_size=6000, measured_errors=np.resize(M17, 6000), number_of_samples=3)
c.calculate_eccentric_anomaly()
c.binary_radial_velocity(c.max_orbital_velocity(c.calculate_semi_major_axis()))
_ = c.sample_time(), c.sample_eccentricity(), c.sample_powerlaw(1.0, 2.0, -0.5, size=(7, 3))

rng = np.random.RandomState(0)
for t_obs, errs in (([days, days[:4]], [2.0, [1.0, 2.0, 0.5, 1.5]]),
                    ([np.sort(rng.uniform(0, 3e7, 40)), np.sort(rng.uniform(0, 3e7, 11))], [1.0, 2.0]),
                    ([days[:3]] * 5, 2.0)):
    bc = BiasCorr(t_obs, number_of_samples=1000, rv_errors=errs, seed=3)
    bc.calculate_radial_velocities()
    bc.get_detected_binaries()
    bc.get_detected_binaries(delta_RV=10, n_sigma_RV=3)
    _ = bc.radial_velocities
    bc.generate_single_delta_rv()
    _ = bc.period                       # -> replay mode
    bc.calculate_radial_velocities()
    bc.get_detected_binaries()
bc = BiasCorr([days], number_of_samples=777, rv_errors=1.0, seed=3, physical_phase=True, masses=[30 * 2e30],
              mass_errors=[3 * 2e30])
bc.calculate_radial_velocities()

pops = grid.points_array(40, fbin=np.linspace(0, 1, 40))
grid.run_points(pops, grid.point_seeds(1, 40), M17, 12, 3001, keep_hist=True, distributed=False)
grid.run_points(pops[:3], grid.point_seeds(1, 3), M17, 12, 2, keep_sigma=True, distributed=False)
ops, ms = C.c_double(), C.c_double()
for kind in range(5):
    _abi.check(_abi.lib().rvmc_pipe_peak(kind, 2, C.byref(ops), C.byref(ms)))
torch.cuda.synchronize()
print("sanitize target finished")
