#!/usr/bin/env python
"""How much of the 1e-5 parity bound do the fp32 kernels use?  Largest |RV - oracle| / max(|RV|, K) of the fused
single-epoch path (rvmc_fused_trace, 1.2e6 slots, two populations) and of the multi-epoch kernel (6 epochs x
4e5 binaries, both phase conventions), oracle = oracle/rvmc_oracle.py on the same Philox draws.  Test
infrastructure (imports oracle/ and tests/): run on a GPU box, prints one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rvmc_oracle as orc  # noqa: E402
from rvmc_b200 import _abi  # noqa: E402
from rvmc_b200 import _device as dv  # noqa: E402
from tests import gpu_util as g  # noqa: E402
from tests.oracle_from_draws import (biascorr_item, fused_slots_from_draws, near_regime_edge,  # noqa: E402
                                     orbits_from_draws)

M17 = np.array([0.8, 2.3, 1.0, 1.8, 0.8, 1.0, 1.4, 2.5, 2.0, 0.4, 1.5, 0.6])
DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
out = {"lib": os.path.basename(_abi.LIB_PATH)}
for name, kw in (("default", {}), ("variant", dict(min_mass=15., max_mass=60., min_period=1.4, max_period=3500.,
                                                    period_pi=0.2, mass_ratio_kappa=-0.3, mass_power=-1.9, min_q=0.2,
                                                    max_q=0.95, e_power=0.5))):
    p = _abi.PopParams.defaults(**kw)
    S, seed, first, n = 12, 2024, 12 * 100000 + 7, 1200000
    tens = {k: g.empty((n,), "float32") for k in ("K", "rv")}
    tr = _abi.TraceSoA()
    for k, v in tens.items():
        setattr(tr, k, v.data_ptr())
    g.ok(g.lib().rvmc_fused_trace(C.byref(p), seed, dv.ptr(g.up(M17, "float32")), S, first, n, tr, g.stream()))
    want = fused_slots_from_draws(p, seed, M17, S, first, n)
    keep = ~near_regime_edge(want["orbits"])
    err = g.rel_err(g.host(tens["rv"]), want["rv"], want["K"])[keep]
    out["fused_" + name] = {"max": float(err.max()), "p999": float(np.percentile(err, 99.9)), "median": float(np.median(err))}
p = _abi.PopParams.defaults()
ns = 400000
d_nep = g.up(np.full(1, 6, np.int32), "int32")
d_tobs = g.up(DAYS6[None, :], "float64")
o = orbits_from_draws(p, 77, biascorr_item(0, np.arange(ns)))
K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"], o["inclination"])
keep = ~near_regime_edge(o, 1e-9)
for quirk in (1, 0):
    rv = g.empty((1, 6, ns), "float32")
    g.ok(g.lib().rvmc_biascorr_fused(C.byref(p), 77, 0, 0, 1, 6, dv.ptr(d_nep), dv.ptr(d_tobs), None, 0, None, None, quirk,
                                     20.0, 4.0, 0, ns, ns, None, None, dv.ptr(rv), None, None, g.stream()))
    if quirk:
        want = orc.multi_epoch_rv(DAYS6, o["time"], o["period"], o["eccentricity"], o["mass_ratio"], o["primary_mass"],
                                  o["inclination"], o["orbit_rotation"])
    else:
        tt = DAYS6[:, None] + o["time"]
        E = orc.kepler_exact(2 * np.pi * ((tt / o["period"]) % 1.0), o["eccentricity"])
        a = orc.semi_major_axis(o["mass_ratio"], o["primary_mass"], o["period"])
        vm = orc.max_orbital_velocity(o["mass_ratio"], o["primary_mass"], o["eccentricity"], a)
        want = orc.binary_radial_velocity(vm, o["eccentricity"], o["inclination"], o["orbit_rotation"], E)
    err = g.rel_err(g.host(rv)[0], want, K)[:, keep]
    out["multi_epoch_quirk%d" % quirk] = {"max": float(err.max()), "p999": float(np.percentile(err, 99.9)),
                                          "median": float(np.median(err))}
print(json.dumps(out))
