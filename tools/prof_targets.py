#!/usr/bin/env python
"""Small fixed launches of the three hot kernels for ncu (run plain first, then under ncu):
   python tools/prof_targets.py [biascorr|sigma|grid|all]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rvmc_b200 import _abi, grid  # noqa: E402
from rvmc_b200 import _device as dv  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _abi.lib()
st = dv.stream_ptr(dev)
p = _abi.PopParams.defaults()
DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0

if what in ("biascorr", "all"):
    nstars, ns = 100, 200000
    d_nep = dv.to_device(np.full(nstars, 6, np.int32), "int32", dev)
    d_tobs = dv.to_device(np.tile(DAYS6, (nstars, 1)), "float64", dev)
    d_err = dv.to_device(np.full((nstars, 6), 2.0, np.float32), "float32", dev)
    drv = dv.empty((nstars, ns), "float32", dev)
    det = dv.empty((nstars, ns), "uint8", dev)
    cnt = dv.zeros((4,), "int64", dev)
    for _ in range(4):
        _abi.check(lib.rvmc_biascorr_fused(C.byref(p), 1, 2, 0, nstars, 6, dv.ptr(d_nep), dv.ptr(d_tobs), dv.ptr(d_err), 0,
                                           None, None, 1, 20.0, 4.0, 0, ns, ns, dv.ptr(drv), dv.ptr(det), None,
                                           dv.ptr(cnt), None, st))
    torch.cuda.synchronize()
    print("biascorr", dv.to_host(cnt))

if what in ("sigma", "all"):
    errd = dv.to_device(grid.M17_ERRORS.astype(np.float32), "float32", dev)
    n_real = 4_000_000
    out = dv.empty((n_real,), "float32", dev)
    cnt = dv.zeros((4,), "int64", dev)
    for _ in range(4):
        _abi.check(lib.rvmc_sigma1d_fused(C.byref(p), 7, dv.ptr(errd), 12, 0, 1, 0, n_real, dv.ptr(out), dv.ptr(cnt), st))
    torch.cuda.synchronize()
    print("sigma", dv.to_host(cnt))

if what in ("grid", "all"):
    npts = 256
    pops = [_abi.PopParams.defaults(period_pi=-0.95 + 1.45 * (i % 8) / 7.0, mass_ratio_kappa=-0.95 + 1.95 * ((i // 8) % 8) / 7.0,
                                    fbin=0.0625 + 0.9375 * ((i // 64) % 4) / 3.0) for i in range(npts)]
    for _ in range(3):
        r = grid.run_points(pops, grid.point_seeds(3, npts), grid.M17_ERRORS, 12, 10 ** 5, distributed=False)
    torch.cuda.synchronize()
    print("grid", r["counters"], float(np.median(r["stats"]["median"])))
