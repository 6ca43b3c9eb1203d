"""GPU parity, part 2: pool mixing, finite-pool bootstrap, and the fused sigma_1D kernel (K1-K4)."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats as sps

from oracle import rvmc_oracle as orc
from rvmc_b200 import _abi
from rvmc_b200 import _device as dv
from tests import gpu_util as g
from tests.oracle_from_draws import fused_slots_from_draws, near_regime_edge, pool_bootstrap_from_draws

pytestmark = pytest.mark.gpu

M17 = np.array([0.8, 2.3, 1.0, 1.8, 0.8, 1.0, 1.4, 2.5, 2.0, 0.4, 1.5, 0.6])


# ---- a15: mix_pool ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,fbin", [(1000, 0.7), (4097, 0.28), (17, 1.0), (50, 0.0), (1, 0.5)])
def test_mix_pool_semantics(n, fbin):
    """RVMC.py:228-236: first int(n*fbin) = distinct binaries, rest = singles with replacement."""
    rvb = np.arange(n, dtype=np.float64) + 1000.0          # every binary value identifies its index
    cl = -(np.arange(n, dtype=np.float64) + 1.0)
    out = g.empty((n,), "float64")
    g.ok(g.lib().rvmc_mix_pool(11, 0, n, fbin, dv.ptr(g.up(rvb, "float64")), dv.ptr(g.up(cl, "float64")),
                               dv.ptr(out), g.stream()))
    out = g.host(out)
    nb = int(n * fbin)
    b, s = out[:nb], out[nb:]
    assert np.all(b >= 1000) and len(np.unique(b)) == nb           # without replacement
    assert np.all(s < 0) and np.all(np.isin(s, cl))
    if nb > 200:
        # a selection, not a prefix, and different under another seed
        assert not np.array_equal(np.sort(b), rvb[:nb])
        out2 = g.empty((n,), "float64")
        g.ok(g.lib().rvmc_mix_pool(12, 0, n, fbin, dv.ptr(g.up(rvb, "float64")), dv.ptr(g.up(cl, "float64")),
                                   dv.ptr(out2), g.stream()))
        assert not np.array_equal(g.host(out2), out)
        # uniformity of the selected indices (KS against U[0,n))
        assert sps.kstest((b - 1000.0 + 0.5) / n, "uniform").pvalue > 1e-3
    if n - nb > 200:
        assert sps.kstest((-s - 0.5) / n, "uniform").pvalue > 1e-3


# ---- a16: sigma1d_pool -----------------------------------------------------------------------
@pytest.mark.parametrize("S,weighted,ddof,n_real,first", [(12, 0, 1, 1000, 0), (12, 1, 1, 777, 5), (20, 0, 0, 513, 3),
                                                          (5, 0, 1, 2000, 1), (332, 0, 1, 64, 0), (1, 0, 0, 100, 0),
                                                          (33, 1, 0, 100, 7), (2500, 0, 1, 3, 0)])
def test_sigma1d_pool_matches_oracle_on_same_draws(S, weighted, ddof, n_real, first):
    rng = np.random.RandomState(S)
    pool = rng.normal(0, 30, 5000)
    err = np.resize(M17, S) if S != 12 else M17
    out = g.empty((n_real,), "float64")
    seed = 31337
    g.ok(g.lib().rvmc_sigma1d_pool(seed, 2, dv.ptr(g.up(pool, "float64")), len(pool), dv.ptr(g.up(err, "float64")),
                                   S, weighted, ddof, first, n_real, dv.ptr(out), g.stream()))
    rvs = pool_bootstrap_from_draws(seed, 2, pool, err, S, first, n_real)
    want = orc.sigma1d_weighted(rvs, err) * np.sqrt(S / (S - ddof)) if weighted else \
        orc.sigma1d_unweighted(rvs, ddof=ddof)          # weighted + ddof=1: the commented variant of RVMC.py:267
    np.testing.assert_allclose(g.host(out), want, rtol=1e-11, atol=1e-12)


def test_sigma1d_pool_golden_arithmetic():
    """The reference's own sigma_1D values from its recorded choice/normal draws: checks the
    dispersion arithmetic (ddof=1 and weighted) end to end against RVMC.py:262-271."""
    z, d = g.golden("rvmc_default_n2048")
    pool2 = orc.mix_radial_velocities(z["rv_binary"], z["cluster_velocities"], 0.7, d[12], d[13])
    np.testing.assert_allclose(orc.sigma1d_unweighted(pool2[d[14]] + d[15]), z["sigma_1d_unweighted"], rtol=1e-12)
    # feed the SAME per-star values through the kernel: pool = the realised values, indices forced
    # by a pool of exactly one value per slot is not possible, so check single-realization identity
    vals = (pool2[d[14]] + d[15])[0]
    out = g.empty((1,), "float64")
    one = np.array([vals.mean()])       # degenerate pool: all picks equal -> sigma = noise only
    g.ok(g.lib().rvmc_sigma1d_pool(1, 0, dv.ptr(g.up(one, "float64")), 1, dv.ptr(g.up(np.zeros(len(vals)), "float64")),
                                   len(vals), 0, 1, 0, 1, dv.ptr(out), g.stream()))
    assert g.host(out)[0] == 0.0


def test_sigma1d_pool_errors():
    pool = g.up(np.zeros(4), "float64")
    err = g.up(np.ones(5000), "float64")
    out = g.empty((4,), "float64")
    assert g.lib().rvmc_sigma1d_pool(1, 0, dv.ptr(pool), 4, dv.ptr(err), 5000, 0, 1, 0, 4, dv.ptr(out),
                                     g.stream()) == _abi.RVMC_ERR_UNSUPPORTED
    assert g.lib().rvmc_sigma1d_pool(1, 0, dv.ptr(pool), 4, dv.ptr(err), 0, 0, 1, 0, 4, dv.ptr(out),
                                     g.stream()) == _abi.RVMC_ERR_INVALID
    g.ok(g.lib().rvmc_sigma1d_pool(1, 0, dv.ptr(pool), 4, dv.ptr(err), 3, 0, 1, 0, 0, dv.ptr(out), g.stream()))


# ---- K1-K4 fused ---------------------------------------------------------------------------------
def _trace(p, seed, err, S, first_slot, n):
    names = ["log2m", "x", "uphase", "q", "e", "cosi", "wturn", "K", "rv", "noise"]
    tens = {k: g.empty((n,), "float32") for k in names}
    tens["words"] = g.empty((n, 12), "int32")
    tens["is_binary"] = g.empty((n,), "uint8")
    tr = _abi.TraceSoA()
    for k, v in tens.items():
        setattr(tr, k, v.data_ptr())
    g.ok(g.lib().rvmc_fused_trace(C.byref(p), seed, dv.ptr(g.up(err, "float32")), S, first_slot, n, tr, g.stream()))
    return {k: g.host(v) for k, v in tens.items()}


@pytest.mark.parametrize("variant", [{}, dict(min_mass=15., max_mass=60., min_period=1.4, max_period=3500.,
                                              period_pi=0.2, mass_ratio_kappa=-0.3, mass_power=-1.9, min_q=0.2,
                                              max_q=0.95, e_power=0.5, sigma_dyn=3.5, fbin=0.28)])
def test_fused_per_orbit_rv_matches_oracle_within_1e5(variant):
    """north_star: per-orbit RVs within 1e-5 relative of the reference's NumPy path on the same
    uniform draws (relative to max(|RV|, K))."""
    p = _abi.PopParams.defaults(**variant)
    S, seed, first, n = 12, 2024, 12 * 100000 + 7, 120000
    tr = _trace(p, seed, M17, S, first, n)
    want = fused_slots_from_draws(p, seed, M17, S, first, n)
    keep = ~near_regime_edge(want["orbits"])
    assert (~keep).sum() <= 5
    assert g.rel_err(tr["rv"], want["rv"], want["K"])[keep].max() < 1e-5
    assert np.max(np.abs(tr["K"] / want["K"] - 1)[keep]) < 5e-6
    np.testing.assert_array_equal(tr["is_binary"].astype(bool), want["is_binary"])
    assert np.max(np.abs(tr["noise"] - want["noise"])) < 2e-5 * max(1.0, np.abs(want["noise"]).max())
    frac = tr["is_binary"].mean()
    assert abs(frac - p.fbin) < 4 * np.sqrt(p.fbin * (1 - p.fbin) / n)


@pytest.mark.parametrize("S,weighted,ddof,fbin,n_real,first", [(12, 0, 1, 0.7, 3000, 0), (12, 1, 1, 0.7, 1000, 11),
                                                              (12, 0, 0, 0.12, 2000, 3), (5, 0, 1, 1.0, 2001, 1),
                                                              (100, 0, 1, 0.7, 300, 0), (332, 1, 0, 0.7, 50, 2),
                                                              (7, 0, 1, 0.0, 500, 9), (1, 0, 0, 0.5, 64, 0),
                                                              (4096, 0, 1, 0.7, 3, 0), (5000, 0, 1, 0.3, 2, 1)])
def test_fused_sigma_equals_dispersion_of_its_own_slots(S, weighted, ddof, fbin, n_real, first):
    """sigma_out must be the dispersion of exactly the per-slot values the trace entry reports
    (queue compaction, tiling and reduction lose or duplicate nothing), and match the oracle's
    fp64 dispersion of the oracle's per-slot values to fp32 accuracy."""
    p = _abi.PopParams.defaults(fbin=fbin)
    err = np.resize(M17, S)
    seed = 555
    out = g.empty((n_real,), "float32")
    cnt = g.zeros((4,), "int64")
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), seed, dv.ptr(g.up(err, "float32")), S, weighted, ddof, first, n_real,
                                    dv.ptr(out), dv.ptr(cnt), g.stream()))
    got = g.host(out).astype(np.float64)
    tr = _trace(p, seed, err, S, first * S, n_real * S)
    v = (tr["noise"].astype(np.float64) + np.where(tr["is_binary"] > 0, tr["rv"].astype(np.float64), 0.0))
    v = v.reshape(n_real, S)
    with np.errstate(invalid="ignore", divide="ignore"):
        bessel = np.sqrt(S / (S - ddof)) if (weighted and S > ddof) else 1.0
        want = orc.sigma1d_weighted(v, err) * bessel if weighted else orc.sigma1d_unweighted(v, ddof=ddof)
    if S - ddof == 0 and not weighted:
        assert np.all(np.isnan(got))
    else:
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    c = g.host(cnt)
    assert c[0] == int(tr["is_binary"].sum()) and c[2] == 0
    if S <= 332 and S - ddof > 0:
        o = fused_slots_from_draws(p, seed, err, S, first * S, n_real * S)
        vo = o["value"].reshape(n_real, S)
        wo = orc.sigma1d_weighted(vo, err) * bessel if weighted else orc.sigma1d_unweighted(vo, ddof=ddof)
        keep = ~near_regime_edge(o["orbits"]).reshape(n_real, S).any(axis=1)
        np.testing.assert_allclose(got[keep], wo[keep], rtol=5e-5, atol=5e-4)


def test_fused_guard_path_with_extreme_eccentricities():
    """With eccentricity bounds beyond the reference's 0.9 the fp64 guard of the Kepler solve becomes
    reachable: the guard-enabled kernel variant is selected, engages, and the per-orbit RVs keep the
    1e-5 bound (checked against an exact fp64 solve of the same draws)."""
    p = _abi.PopParams.defaults(e_max_long=0.985, e_power=3.0)       # e piles up near e_max
    S, n_real, seed = 12, 4000, 99
    out = g.empty((n_real,), "float32")
    cnt = g.zeros((4,), "int64")
    errd = g.up(M17, "float32")
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), seed, dv.ptr(errd), S, 0, 1, 0, n_real, dv.ptr(out), dv.ptr(cnt),
                                    g.stream()))
    c = g.host(cnt)
    assert c[1] > 0 and c[2] == 0                                    # guard lanes, no non-converged lanes
    tr = _trace(p, seed, M17, S, 0, n_real * S)
    v = (tr["noise"].astype(np.float64) + np.where(tr["is_binary"] > 0, tr["rv"].astype(np.float64), 0.0))
    np.testing.assert_allclose(g.host(out), orc.sigma1d_unweighted(v.reshape(n_real, S), ddof=1), rtol=2e-5, atol=2e-5)
    # exact reference values for the traced orbits
    e = tr["e"].astype(np.float64)
    M = 2 * np.pi * tr["uphase"].astype(np.float64)
    E = orc.kepler_exact(M, e)
    om = 2 * np.pi * tr["wturn"].astype(np.float64)
    d = 1 - e * np.cos(E)
    shape = (np.cos(E) - e) / d * np.cos(om) - np.sqrt(1 - e * e) * np.sin(E) / d * np.sin(om) + e * np.cos(om)
    K = tr["K"].astype(np.float64)
    assert e.max() > 0.97
    assert np.max(np.abs(tr["rv"] - K * shape) / np.maximum(K, 1e-30)) < 1e-5


def test_fused_sigma_is_partition_invariant():
    """Counter = global slot index: any split of the realizations gives bit-identical results."""
    p = _abi.PopParams.defaults()
    err = g.up(M17, "float32")
    n = 5000
    full = g.empty((n,), "float32")
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), 9, dv.ptr(err), 12, 0, 1, 100, n, dv.ptr(full), None, g.stream()))
    parts = []
    for a, b in ((0, 1), (1, 342), (342, 1707), (1707, 5000)):
        t = g.empty((b - a,), "float32")
        g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), 9, dv.ptr(err), 12, 0, 1, 100 + a, b - a, dv.ptr(t), None,
                                        g.stream()))
        parts.append(g.host(t))
    np.testing.assert_array_equal(np.concatenate(parts), g.host(full))
    again = g.empty((n,), "float32")
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), 9, dv.ptr(err), 12, 0, 1, 100, n, dv.ptr(again), None, g.stream()))
    np.testing.assert_array_equal(g.host(again), g.host(full))


def _oracle_infinite_pool_sigma(rng, p, err, S, n_real, ddof=1):
    """Independent-RNG oracle of the infinite-pool model: the reference's samplers + SciPy-style
    secant + RV formula (oracle/rvmc_oracle.py), NumPy MT19937 draws."""
    n = n_real * S
    m1 = orc.masses_from_uniform(rng.random_sample(n), p.min_mass, p.max_mass, p.mass_power)
    per = orc.period_from_uniform(rng.random_sample(n), p.min_period, p.max_period, p.period_pi)
    t = orc.time_from_uniform(rng.random_sample(n), per)
    q = orc.mass_ratio_from_uniform(rng.random_sample(n), p.min_q, p.max_q, p.mass_ratio_kappa)
    e = orc.eccentricity_from_uniform_slotted(per, rng.random_sample(n), p.e_power)
    inc = orc.inclination_from_uniform(rng.random_sample(n))
    om = orc.orbit_rotation_from_uniform(rng.random_sample(n))
    E = orc.eccentric_anomaly_single_epoch(t, per, e)
    rv = orc.orbital_rv(m1, per, q, e, inc, om, E)
    is_bin = rng.random_sample(n) < p.fbin
    v = np.where(is_bin, rv, 0.0) + rng.normal(0, p.sigma_dyn, n) + rng.normal(0, np.tile(err, n_real), n)
    return orc.sigma1d_unweighted(v.reshape(n_real, S), ddof=ddof)


@pytest.mark.parametrize("variant", [dict(fbin=0.7), dict(fbin=0.28, min_period=30.0, max_period=3500.0)])
def test_fused_sigma_distribution_ks_vs_oracle_independent_rng(variant):
    """north_star: dispersion distributions agree with the reference under a two-sample KS test
    at p > 0.01 with independent RNG."""
    p = _abi.PopParams.defaults(**variant)
    n_real = 20000
    out = g.empty((n_real,), "float32")
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), 20260101, dv.ptr(g.up(M17, "float32")), 12, 0, 1, 0, n_real,
                                    dv.ptr(out), None, g.stream()))
    want = _oracle_infinite_pool_sigma(np.random.RandomState(7), p, M17, 12, n_real)
    ks = sps.ks_2samp(g.host(out).astype(np.float64), want)
    assert ks.pvalue > 0.01, ks
    assert abs(np.median(g.host(out)) / np.median(want) - 1) < 0.03


def test_fused_errors_and_edges():
    err = g.up(np.ones(9000), "float32")
    out = g.empty((8,), "float32")
    p = _abi.PopParams.defaults()
    L = g.lib()
    assert L.rvmc_sigma1d_fused(C.byref(p), 1, dv.ptr(err), 9000, 0, 1, 0, 8, dv.ptr(out), None,
                                g.stream()) == _abi.RVMC_ERR_UNSUPPORTED
    assert L.rvmc_sigma1d_fused(C.byref(p), 1, dv.ptr(err), 0, 0, 1, 0, 8, dv.ptr(out), None,
                                g.stream()) == _abi.RVMC_ERR_INVALID
    assert L.rvmc_sigma1d_fused(C.byref(_abi.PopParams.defaults(e_power=-1.0)), 1, dv.ptr(err), 12, 0, 1, 0, 8,
                                dv.ptr(out), None, g.stream()) == _abi.RVMC_ERR_ZERODIV
    assert L.rvmc_sigma1d_fused(C.byref(_abi.PopParams.defaults(fbin=1.5)), 1, dv.ptr(err), 12, 0, 1, 0, 8,
                                dv.ptr(out), None, g.stream()) == _abi.RVMC_ERR_INVALID
    assert L.rvmc_sigma1d_fused(C.byref(_abi.PopParams.defaults(min_period=0.9)), 1, dv.ptr(err), 12, 0, 1, 0, 8,
                                dv.ptr(out), None, g.stream()) == _abi.RVMC_ERR_INVALID
    g.ok(L.rvmc_sigma1d_fused(C.byref(p), 1, dv.ptr(err), 12, 0, 1, 0, 0, dv.ptr(out), None, g.stream()))
    # no binaries and no noise: every dispersion is exactly zero
    p0 = _abi.PopParams.defaults(fbin=0.0, sigma_dyn=0.0)
    g.ok(L.rvmc_sigma1d_fused(C.byref(p0), 1, dv.ptr(g.up(np.zeros(12), "float32")), 12, 0, 1, 0, 8, dv.ptr(out), None,
                              g.stream()))
    assert np.all(g.host(out) == 0)
