"""How fast is the oracle port relative to the UNMODIFIED reference on the same CPU?  (build container
only: reads /root/reference).  The bench's cpu_baseline / --impl reference legs time the port because
the Python reference cannot travel to the GPU box; this script documents that the two are equivalent.

    python oracle/compare_speed.py
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shims"))
warnings.simplefilter("ignore")
import RVMC_bias_corr as ref_bc  # noqa: E402

import bench  # noqa: E402

DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0


def time_reference(nstars, ns, reps=3):
    best = 1e9
    for r in range(reps):
        np.random.seed(r)
        t0 = time.perf_counter()
        bc = ref_bc.BiasCorr([DAYS6] * nstars, number_of_samples=ns, rv_errors=2.0)
        bc.calculate_radial_velocities()
        bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
        best = min(best, time.perf_counter() - t0)
    return nstars * ns * 6 / best


def time_port(nstars, ns, reps=3):
    best = 1e9
    for r in range(reps):
        t0 = time.perf_counter()
        bench._cpu_chunk((r, nstars, ns))
        best = min(best, time.perf_counter() - t0)
    return nstars * ns * 6 / best


if __name__ == "__main__":
    a, b = time_reference(2, 50000), time_port(2, 50000)
    print("reference BiasCorr (unmodified, 1 core): %.3g binary-epoch RV/s" % a)
    print("oracle port        (bench._cpu_chunk, 1 core): %.3g binary-epoch RV/s" % b)
    print("port / reference = %.2f" % (b / a))
