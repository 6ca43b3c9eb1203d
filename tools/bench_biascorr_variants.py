#!/usr/bin/env python
"""Throughput of rvmc_biascorr_fused away from the headline configuration: physical phase, two
Halley steps, per-epoch errors (pairwise test), masses with errors, 8/16/24 epochs, RVs written out."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rvmc_b200 import _abi  # noqa: E402
from rvmc_b200 import _device as dv  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _abi.lib()
st = dv.stream_ptr(dev)
DAYS = np.array([0, 1, 7, 30, 180, 365.0, 400, 500, 730, 800, 900, 1000, 1100, 1200, 1300, 1400, 1500, 1600, 1700, 1800,
                 1900, 2000, 2100, 2200]) * 86400.0


def run(name, nstars=100, ns=10 ** 6, ne=6, quirk=1, iters=1, uniform_err=True, mass_mode=0, rv_out=False, steps=10,
        no_errors=False, **pop):
    p = _abi.PopParams.defaults(kepler_iters=iters, **pop)
    d_nep = dv.to_device(np.full(nstars, ne, np.int32), "int32", dev)
    d_tobs = dv.to_device(np.tile(DAYS[:ne], (nstars, 1)), "float64", dev)
    err = np.full((nstars, ne), 2.0, np.float32)
    if not uniform_err:
        err = err * np.linspace(0.5, 1.5, ne, dtype=np.float32)[None, :]
    d_err = None if no_errors else dv.to_device(err, "float32", dev)
    masses = dv.to_device(np.full(nstars, 20 * 2e30), "float64", dev) if mass_mode else None
    merr = dv.to_device(np.full(nstars, 2 * 2e30), "float64", dev) if mass_mode == 1 else None
    drv = dv.empty((nstars, ns), "float32", dev)
    det = dv.empty((nstars, ns), "uint8", dev)
    rvo = dv.empty((nstars, ne, ns), "float32", dev) if rv_out else None
    cnt = dv.zeros((4,), "int64", dev)

    def step():
        _abi.check(lib.rvmc_biascorr_fused(C.byref(p), 1, 2, 0, nstars, ne, dv.ptr(d_nep), dv.ptr(d_tobs), dv.ptr(d_err),
                                           mass_mode, dv.ptr(masses), dv.ptr(merr), quirk, 20.0, 4.0, 0, ns, ns,
                                           dv.ptr(drv), dv.ptr(det), dv.ptr(rvo), dv.ptr(cnt), None, st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"case": name, "epochs": ne, "ms": ms, "binary_epoch_rvs_per_s": nstars * ns * ne / (ms * 1e-3)}))


run("headline (quirk phase, 1 Halley step, equal errors)")
run("no measurement errors (rv_errors=None, the docstring example: Delta-RV only)", no_errors=True)
run("physical phase", quirk=0)
run("per-epoch errors (15 pair tests)", uniform_err=False)
run("two Halley steps (general instantiation)", iters=2)
run("masses with errors (mass_mode 1)", mass_mode=1)
run("per-epoch RVs written to HBM (2.4 GB)", rv_out=True, ns=10 ** 6)
run("period_pi = -0.3 (general period-law exponent)", period_pi=-0.3)
run("period_pi = 0 (flat in log P)", period_pi=0.0)
run("kappa = -0.5, e_power = 0.3 (general q and e exponents)", mass_ratio_kappa=-0.5, e_power=0.3)
run("8 epochs", ne=8, ns=750000)
run("16 epochs", ne=16, ns=375000)
run("24 epochs (run-time epoch loop)", ne=24, ns=250000)
