"""Test helper: regenerate on the CPU, from (seed, item), the Philox draws a kernel consumed and
push them through the oracle -- the 'same exported uniform draws' leg of the parity protocol."""
import numpy as np

from oracle import philox_ref as ph
from oracle import rvmc_oracle as orc


def pop_kwargs(p):
    return {n: getattr(p, n) for n, _ in p._fields_}


def orbits_from_draws(p, seed, items):
    """Oracle orbit arrays (reference units) for global items under population `p` (PopParams)."""
    wa = ph.philox_block(seed, items, ph.STREAM_ORBIT_A)
    wb = ph.philox_block(seed, items, ph.STREAM_ORBIT_B)
    ua = ph.u01(wa)
    ub = ph.u01(wb)
    assert (p.e_min, p.e_max_mid, p.e_max_long, p.p_circ_days, p.p_mid_days) == (1e-5, 0.5, 0.9, 4.0, 6.0)
    o = {}
    o["primary_mass"] = orc.masses_from_uniform(ua[:, 0], p.min_mass, p.max_mass, p.mass_power)
    o["period"] = orc.period_from_uniform(ua[:, 1], p.min_period, p.max_period, p.period_pi)
    o["time"] = orc.time_from_uniform(ua[:, 2], o["period"])
    o["mass_ratio"] = orc.mass_ratio_from_uniform(ua[:, 3], p.min_q, p.max_q, p.mass_ratio_kappa)
    o["eccentricity"] = orc.eccentricity_from_uniform_slotted(o["period"], ub[:, 0], p.e_power)
    o["inclination"] = orc.inclination_from_uniform(ub[:, 1])
    o["orbit_rotation"] = orc.orbit_rotation_from_uniform(ub[:, 2])
    o["u"] = np.concatenate([ua, ub], axis=1)
    return o


def rv_from_orbits(o):
    """Reference E (SciPy-style secant) and orbital RV / K for oracle orbit arrays."""
    E = orc.eccentric_anomaly_single_epoch(o["time"], o["period"], o["eccentricity"])
    rv = orc.orbital_rv(o["primary_mass"], o["period"], o["mass_ratio"], o["eccentricity"],
                        o["inclination"], o["orbit_rotation"], E)
    K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"],
                           o["inclination"])
    return E, rv, K


def near_regime_edge(o, tol=2e-6):
    """Orbits whose log10 P sits within `tol` of the 4 d / 6 d eccentricity switches: fp32 and
    fp64 may legitimately land on different sides (documented exclusion)."""
    x = np.log10(o["period"] / 86400.0)
    return (np.abs(x - np.log10(4.0)) < tol) | (np.abs(x - np.log10(6.0)) < tol)
