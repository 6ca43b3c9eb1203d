"""The kernels' per-orbit source (philox/rvmc_math/orbit .cuh) compiled for the CPU and checked
against the oracle on identical Philox draws.  This is NOT the product path (that is CUDA-only,
tests/test_gpu_*.py); it lets logic errors surface without a GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import philox_ref as ph
from oracle import rvmc_oracle as orc
from rvmc_b200._abi import PopParams
from tests.hostcheck_lib import load_hostcheck
from tests.oracle_from_draws import near_regime_edge, orbits_from_draws, rv_from_orbits

VARIANTS = {
    "sana_defaults": {},
    "variant": dict(min_mass=15., max_mass=60., min_period=1.4, max_period=3500., period_pi=0.2,
                    mass_ratio_kappa=-0.3, mass_power=-1.9, min_q=0.2, max_q=0.95, e_power=0.5),
    "steep": dict(period_pi=-0.95, mass_ratio_kappa=1.0, e_power=3.0),
}


def _f32(n):
    return np.empty(n, dtype=np.float32)


def _run_f32(p, seed, first, n):
    lib = load_hostcheck()
    names = ["log2m", "x", "uphase", "q", "e", "cosi", "wturn", "K", "rv"]
    arrs = {k: _f32(n) for k in names}
    guard = np.empty(n, dtype=np.int32)
    bad = np.empty(n, dtype=np.int32)
    rc = lib.hc_orbits_f32(C.byref(p), C.c_uint64(seed), C.c_uint64(first), C.c_int64(n),
                           *[arrs[k].ctypes.data_as(C.c_void_p) for k in names],
                           guard.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(C.c_void_p))
    assert rc == 0
    arrs["guard"], arrs["bad"] = guard, bad
    return arrs


def _run_f64(p, seed, first, n):
    lib = load_hostcheck()
    names = ["primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination",
             "orbit_rotation", "E", "rv", "K"]
    arrs = {k: np.empty(n) for k in names}
    rc = lib.hc_orbits_f64(C.byref(p), C.c_uint64(seed), C.c_uint64(first), C.c_int64(n),
                           *[arrs[k].ctypes.data_as(C.c_void_p) for k in names])
    assert rc == 0
    return arrs


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_fp64_path_matches_oracle(variant):
    p = PopParams.defaults(**VARIANTS[variant])
    seed, first, n = 0xABCDEF12345, 2 ** 33 + 5, 20000
    items = np.arange(first, first + n, dtype=np.uint64)
    o = orbits_from_draws(p, seed, items)
    got = _run_f64(p, seed, first, n)
    keep = ~near_regime_edge(o, 1e-12)
    for k in ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination", "orbit_rotation"):
        np.testing.assert_allclose(got[k][keep], o[k][keep], rtol=2e-14, atol=1e-300, err_msg=k)
    E, rv, K = rv_from_orbits(o)
    assert np.max(np.abs(got["E"] - E)[keep]) < 1e-9          # the reference's own root accuracy
    # north_star: 1e-9 relative on the fp64 path (measured against max(|RV|, K))
    err = np.abs(got["rv"] - rv) / np.maximum(np.abs(rv), K)
    assert err[keep].max() < 1e-9
    np.testing.assert_allclose(got["K"][keep], K[keep], rtol=1e-13)


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_fp32_path_matches_oracle_within_1e5(variant):
    p = PopParams.defaults(**VARIANTS[variant])
    seed, first, n = 77, 123456789, 200000
    items = np.arange(first, first + n, dtype=np.uint64)
    o = orbits_from_draws(p, seed, items)
    got = _run_f32(p, seed, first, n)
    keep = ~near_regime_edge(o)
    assert (~keep).sum() <= 5
    x = np.log10(o["period"] / 86400.0)
    # fp32 rounding of the draw is amplified by the inverse-CDF exponent 1/(pi+1) (20 for pi=-0.95)
    amp = max(1.0, abs(1.0 / (p.period_pi + 1.0)) / 2.0)
    assert np.max(np.abs(got["x"] - x)[keep]) < 2e-6 * amp
    assert np.max(np.abs(2.0 ** got["log2m"].astype(np.float64) * 2e30 / o["primary_mass"] - 1)) < 2e-6
    assert np.max(np.abs(got["q"] / o["mass_ratio"] - 1)) < 1e-6
    assert np.max(np.abs(got["e"] - o["eccentricity"])[keep]) < 2e-6
    np.testing.assert_array_equal(got["uphase"].astype(np.float64), o["u"][:, 2])      # exact
    np.testing.assert_array_equal(got["cosi"].astype(np.float64), o["u"][:, 5])
    np.testing.assert_array_equal(got["wturn"].astype(np.float64), o["u"][:, 6])
    E, rv, K = rv_from_orbits(o)
    err = np.abs(got["rv"].astype(np.float64) - rv) / np.maximum(np.abs(rv), K)
    assert err[keep].max() < 1e-5, err[keep].max()
    assert got["bad"].sum() == 0
    assert got["guard"].sum() == 0          # e <= 0.9 never needs the fp64 guard


def test_fp64_guard_engages_for_extreme_eccentricity():
    lib = load_hostcheck()
    rng = np.random.RandomState(4)
    n = 200000
    e = rng.uniform(0.9, 0.995, n).astype(np.float32)
    phase = ((rng.randint(0, 2 ** 24, n) / 2.0 ** 24) ** 2 * 0.5 * rng.choice([-1, 1], n)).astype(np.float32)
    wt = (rng.randint(0, 2 ** 24, n) / 2.0 ** 24).astype(np.float32)
    shape, E, guard = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.int32)
    lib.hc_kepler_shape_f32(C.c_int64(n), phase.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                            wt.ctypes.data_as(C.c_void_p), 1, C.c_float(64.0),
                            shape.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p),
                            guard.ctypes.data_as(C.c_void_p))
    e64, M = e.astype(np.float64), 2 * np.pi * phase.astype(np.float64)
    Ex = np.sign(M) * orc.kepler_exact(np.abs(M), e64)
    om = 2 * np.pi * wt.astype(np.float64)
    d = 1 - e64 * np.cos(Ex)
    want = (np.cos(Ex) - e64) / d * np.cos(om) - np.sqrt(1 - e64 ** 2) * np.sin(Ex) / d * np.sin(om) + e64 * np.cos(om)
    err = np.abs(shape - want)
    assert guard.sum() > 100                      # the guard path is exercised
    assert err.max() < 1e-5
    amp = np.sqrt(1 - e64 ** 2) / d ** 2
    assert np.all(guard[amp > 70] == 1) and np.all(guard[amp < 58] == 0)


def test_kepler_f64_any_argument():
    lib = load_hostcheck()
    rng = np.random.RandomState(8)
    n = 50000
    M = rng.uniform(-50, 2000, n)
    e = rng.uniform(0, 0.9, n)
    E = np.empty(n)
    lib.hc_kepler_f64(C.c_int64(n), M.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p),
                      E.ctypes.data_as(C.c_void_p))
    assert np.max(np.abs(E - e * np.sin(E) - M)) < 1e-9 * 2000
    # unique root: compare against the reference-style secant from x0 = M on [0, 2 pi)
    M2 = rng.uniform(0, 2 * np.pi, 4000)
    e2 = rng.uniform(0, 0.9, 4000)
    E2 = np.empty(4000)
    lib.hc_kepler_f64(C.c_int64(4000), M2.ctypes.data_as(C.c_void_p), e2.ctypes.data_as(C.c_void_p),
                      E2.ctypes.data_as(C.c_void_p))
    root, conv, _ = orc.secant_array(orc.anomaly_function, M2, (e2, M2), 1e-6)
    assert conv.all() and np.max(np.abs(root - E2)) < 1e-9


def test_quirk_phase_matches_reference_expression():
    lib = load_hostcheck()
    rng = np.random.RandomState(2)
    n = 20000
    P = 86400 * 10 ** rng.uniform(0.15, 3.5, n)
    t = rng.uniform(0, np.pi * 1e7, n) + rng.uniform(0, 1, n) * P
    for quirk in (1, 0):
        M = np.empty(n)
        lib.hc_mean_anomaly_epoch(C.c_int64(n), t.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p),
                                  quirk, M.ctypes.data_as(C.c_void_p))
        want = orc.mean_anomaly_multi_epoch(t, P) if quirk else 2 * np.pi * ((t / P) % 1.0)
        assert np.max(np.abs(M - want)) < 1e-9


def test_box_muller_and_sincos():
    lib = load_hostcheck()
    rng = np.random.RandomState(0)
    n = 100000
    w1 = rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    w2 = rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    z0, z1 = np.empty(n, np.float32), np.empty(n, np.float32)
    d0, d1 = np.empty(n), np.empty(n)
    lib.hc_box_muller(C.c_int64(n), w1.ctypes.data_as(C.c_void_p), w2.ctypes.data_as(C.c_void_p),
                      z0.ctypes.data_as(C.c_void_p), z1.ctypes.data_as(C.c_void_p),
                      d0.ctypes.data_as(C.c_void_p), d1.ctypes.data_as(C.c_void_p))
    r0, r1 = ph.box_muller(w1, w2)
    assert np.max(np.abs(d0 - r0)) < 1e-12 and np.max(np.abs(d1 - r1)) < 1e-12
    assert np.max(np.abs(z0 - r0)) < 5e-6 and np.max(np.abs(z1 - r1)) < 5e-6
    assert abs(d0.mean()) < 0.02 and abs(d0.std() - 1) < 0.02
    t = rng.uniform(-1, 2, n).astype(np.float32)
    s, c = np.empty(n, np.float32), np.empty(n, np.float32)
    lib.hc_sincos(C.c_int64(n), t.ctypes.data_as(C.c_void_p), 1, s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    assert np.max(np.abs(s - np.sin(2 * np.pi * t.astype(np.float64)))) < 3e-7
    assert np.max(np.abs(c - np.cos(2 * np.pi * t.astype(np.float64)))) < 3e-7
    x = rng.uniform(-1, 7, n).astype(np.float32)
    lib.hc_sincos(C.c_int64(n), x.ctypes.data_as(C.c_void_p), 0, s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) < 2e-7
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) < 2e-7
    # the eccentric anomaly's own path: [0, pi] with the starter's slack past both ends; sin keeps its
    # RELATIVE accuracy towards 0 and pi (periastron/apastron), cos its absolute accuracy
    E = np.concatenate([rng.uniform(-1e-3, np.pi + 1e-3, n), rng.uniform(0, 1e-2, 2000), np.pi - rng.uniform(0, 1e-2, 2000),
                        [0.0, np.pi / 2, np.float32(np.pi)]]).astype(np.float32)
    s, c = np.empty(E.size, np.float32), np.empty(E.size, np.float32)
    lib.hc_sincos(C.c_int64(E.size), E.ctypes.data_as(C.c_void_p), 2, s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    E64 = E.astype(np.float64)
    assert np.max(np.abs(c - np.cos(E64))) < 1.5e-7
    assert np.max(np.abs(s - np.sin(E64))) < 1.5e-7
    inner = (np.abs(E64) > 1e-6) & (np.abs(E64 - np.pi) > 1e-6)
    assert np.max(np.abs(s[inner] / np.sin(E64[inner]) - 1)) < 2.5e-7
    assert s[E == 0][0] == 0.0 and c[E == 0][0] == 1.0


def test_noise6_six_normals_per_block():
    """noise6_*: the device source (compiled for the host) against the NumPy model of the bit fields,
    plus the distribution the fields are meant to give."""
    lib = load_hostcheck()
    rng = np.random.RandomState(3)
    n = 200000
    w = rng.randint(0, 2 ** 32, (n, 4), dtype=np.uint64).astype(np.uint32)
    w[0] = 0                                   # smallest radius uniform, angle 0
    w[1] = 0xFFFFFFFF                          # u1 = 1 exactly -> r = 0
    z32, z64 = np.empty((n, 6), np.float32), np.empty((n, 6))
    lib.hc_noise6(C.c_int64(n), w.ctypes.data_as(C.c_void_p), z32.ctypes.data_as(C.c_void_p),
                  z64.ctypes.data_as(C.c_void_p))
    want = ph.noise6(w).T
    assert np.max(np.abs(z64 - want)) < 1e-12
    assert np.max(np.abs(z32 - want)) < 5e-6
    assert abs(want[0, 0] - np.sqrt(-2 * np.log(2.0 ** -22))) < 1e-12 and np.all(want[1] == 0)
    assert np.abs(want).max() <= 5.5226 and abs(want.mean()) < 5e-3 and abs(want.std() - 1) < 5e-3
    # the six normals of a block are uncorrelated (disjoint bit fields)
    c = np.corrcoef(want[2:].T)
    assert np.max(np.abs(c - np.eye(6))) < 0.01
    from scipy import stats as sps
    assert sps.kstest(want[2:, 3], "norm").pvalue > 1e-3


def test_exp10_fast_f64():
    lib = load_hostcheck()
    y = np.concatenate([np.linspace(-12, 6, 200001), [0.0, -4.9365137424788932 - 0.15, -4.9365137424788932 - 3.5]])
    out = np.empty_like(y)
    lib.hc_exp10_fast(C.c_int64(y.size), y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.max(np.abs(out / 10.0 ** y - 1)) < 5e-14


def test_sincos_f64_bounded():
    lib = load_hostcheck()
    rng = np.random.RandomState(11)
    x = np.concatenate([rng.uniform(-10, 10, 100000), rng.uniform(-300, 300, 20000),
                        np.arange(-8, 9) * np.pi / 4, [0.0]])
    s, c = np.empty_like(x), np.empty_like(x)
    lib.hc_sincos_f64(C.c_int64(x.size), x.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                      c.ctypes.data_as(C.c_void_p))
    assert np.max(np.abs(s - np.sin(x))) < 5e-16
    assert np.max(np.abs(c - np.cos(x))) < 5e-16


def test_index_minus_one_is_zero_division():
    lib = load_hostcheck()
    p = PopParams.defaults(period_pi=-1.0)
    a = _f32(1)
    i = np.empty(1, np.int32)
    rc = lib.hc_orbits_f32(C.byref(p), C.c_uint64(1), C.c_uint64(0), C.c_int64(1),
                           *[a.ctypes.data_as(C.c_void_p)] * 9, i.ctypes.data_as(C.c_void_p),
                           i.ctypes.data_as(C.c_void_p))
    assert rc == -2          # RVMC_ERR_ZERODIV, the reference raises at RVMC.py:112
