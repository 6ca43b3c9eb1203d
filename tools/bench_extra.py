#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): throughput of the fused sigma_1D kernel on
the BASELINE config shapes C1/C3/C4/C5, CUDA-event timed.  python tools/bench_extra.py [--quick]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rvmc_b200 import _abi, grid  # noqa: E402
from rvmc_b200 import _device as dv  # noqa: E402

M17 = grid.M17_ERRORS


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fused_case(name, pop, err, S, n_real, weighted=0, ddof=1):
    dev = torch.device("cuda", 0)
    lib = _abi.lib()
    errd = dv.to_device(np.asarray(err, dtype=np.float32), "float32", dev)
    out = dv.empty((n_real,), "float32", dev)
    cnt = dv.zeros((4,), "int64", dev)
    st = dv.stream_ptr(dev)

    def run():
        _abi.check(lib.rvmc_sigma1d_fused(C.byref(pop), 7, dv.ptr(errd), S, weighted, ddof, 0, n_real, dv.ptr(out),
                                          dv.ptr(cnt), st))
    cnt.zero_()
    run()
    torch.cuda.synchronize()
    nbin = int(dv.to_host(cnt)[0])
    ms = timed(run)
    print(json.dumps({"case": name, "S": S, "n_real": n_real, "fbin": pop.fbin, "ms": ms,
                      "binary_rvs_per_s": nbin / (ms * 1e-3), "star_slots_per_s": n_real * S / (ms * 1e-3),
                      "median_sigma": float(np.median(dv.to_host(out)))}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    scale = 10 if a.quick else 1
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    P = _abi.PopParams.defaults
    # C3: period / fbin variants, S = 12, M17 errors
    for name, kw in (("C3 fbin0.70 Pmin1.4", dict(fbin=0.7, min_period=1.4, max_period=3500.)),
                     ("C3 fbin0.28 Pmin1.4", dict(fbin=0.28, min_period=1.4, max_period=3500.)),
                     ("C3 fbin0.12 Pmin1.4", dict(fbin=0.12, min_period=1.4, max_period=3500.)),
                     ("C3 fbin0.70 Pmin30", dict(fbin=0.7, min_period=30., max_period=3500.)),
                     ("C3 fbin0.70 Pmin8yr", dict(fbin=0.7, min_period=2922., max_period=3500.))):
        fused_case(name, P(**kw), M17, 12, 40_000_000 // scale)
    # C4: clusters of data_RT20 (N = 5 ... 332), weighted
    for n in (5, 12, 44, 332):
        fused_case("C4 N=%d weighted" % n, P(min_mass=15., max_mass=60.), np.resize(M17, n), n,
                   max(1000, 40_000_000 // n // scale), weighted=1, ddof=0)
    # C1: RVMC_example shape, infinite-pool
    fused_case("C1 S=100 1e4 realizations", P(), np.full(100, 1.2), 100, 10 ** 4)
    fused_case("C1 S=100 1e6 realizations", P(), np.full(100, 1.2), 100, 10 ** 6 // scale)
    # C5: grid points per second (sigma + on-device statistics), 1e5 realizations per point
    npts = 8192 // scale
    pi = np.linspace(-0.95, 0.5, 8)
    idx = np.arange(npts)
    pops = grid.points_array(npts, period_pi=pi[idx % 8], mass_ratio_kappa=-0.95 + 1.95 * ((idx // 8) % 8) / 7.0,
                             fbin=0.0625 + 0.9375 * ((idx // 64) % 16) / 15.0)
    seeds = grid.point_seeds(3, npts)

    def run_grid():
        return grid.run_points(pops, seeds, M17, 12, 10 ** 5, distributed=False)
    r = run_grid()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    r = run_grid()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"case": "C5 grid", "points": npts, "realizations": 10 ** 5, "S": 12, "seconds": dt,
                      "points_per_s": npts / dt, "binary_rvs_per_s": float(r["counters"][0]) / dt,
                      "full_grid_65536_points_s": 65536 / (npts / dt)}))
    # C1 finite pool through the class
    from rvmc_b200.rvmc import RVMC
    t0 = time.perf_counter()
    c = RVMC(number_of_stars=10 ** 5, seed=1)
    c.calculate_binary_velocities(materialize=False)
    s = c.subsample(sample_size=100, number_of_samples=10 ** 4)
    dt = time.perf_counter() - t0
    print(json.dumps({"case": "C1 class end-to-end (1e5 pool + RV + 1e4x100 bootstrap, host result)", "seconds": dt,
                      "median_sigma": float(np.median(s))}))


if __name__ == "__main__":
    main()
