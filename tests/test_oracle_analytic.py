"""Analytic known-answer tests of the oracle: closed-form facts the reference's
formulas must satisfy, independent of any recorded run."""
import numpy as np
from scipy import stats as sps

from oracle import rvmc_oracle as orc


def test_powerlaw_inverse_cdf_matches_closed_form_cdf():
    rng = np.random.RandomState(1)
    for lo, hi, power in ((6.0, 20.0, -2.35), (0.1, 1.0, 0.0), (0.15, 3.5, -0.5), (1e-5, 0.9, -0.5), (2.0, 8.0, 1.5)):
        x = orc.powerlaw_from_uniform(rng.random_sample(40000), lo, hi, power)
        assert x.min() >= min(lo, hi) * (1 - 1e-12) and x.max() <= max(lo, hi) * (1 + 1e-12)
        k = power + 1
        cdf = (x ** k - lo ** k) / (hi ** k - lo ** k)          # pdf ~ x^power on [lo, hi]
        assert sps.kstest(cdf, "uniform").pvalue > 1e-3, (lo, hi, power)
    # u = 0 maps to max, u -> 1 maps to min (np.random.uniform(pmax, 1))
    assert orc.powerlaw_from_uniform(0.0, 6.0, 20.0, -2.35) == np.float64(20.0) or \
        abs(orc.powerlaw_from_uniform(0.0, 6.0, 20.0, -2.35) - 20.0) < 1e-12
    assert abs(orc.powerlaw_from_uniform(1.0, 6.0, 20.0, -2.35) - 6.0) < 1e-12


def test_eccentricity_regimes_at_4_and_6_days():
    P = np.array([3.9, 4.0, 4.1, 5.9, 6.0, 6.1, 1000.0]) * 86400.0
    u = np.full(P.size, 0.0)                    # u = 0 -> the upper bound of each regime
    e = orc.eccentricity_from_uniform_slotted(P, u, -0.5)
    np.testing.assert_allclose(e, [0, 0, 0.5, 0.5, 0.5, 0.9, 0.9], rtol=1e-12)
    e = orc.eccentricity_from_uniform_slotted(P, np.full(P.size, 1.0 - 2.0 ** -53), -0.5)
    np.testing.assert_allclose(e[2:], 1e-5, rtol=1e-6)


def test_circular_orbit_has_E_equal_M_and_sinusoidal_rv():
    rng = np.random.RandomState(2)
    n = 2000
    P = np.full(n, 3.0 * 86400.0)
    t = rng.random_sample(n) * P
    E = orc.eccentric_anomaly_single_epoch(t, P, np.zeros(n))
    np.testing.assert_allclose(E, 2 * np.pi * t / P, rtol=0, atol=1e-12)
    m1, q, inc, om = np.full(n, 20 * 2e30), np.full(n, 0.5), np.full(n, np.pi / 2), rng.random_sample(n) * 2 * np.pi
    rv = orc.orbital_rv(m1, P, q, np.zeros(n), inc, om, E)
    K = orc.semi_amplitude(q, m1, P, np.zeros(n), inc)
    np.testing.assert_allclose(rv, K * np.cos(E + om), rtol=1e-12, atol=1e-12)
    # closed-form K = q/(1+q) (2 pi G M1 (1+q) / P)^(1/3) sin i / sqrt(1-e^2) * 1e-3
    Kc = q / (1 + q) * (2 * np.pi * orc.G_VALUE * m1 * (1 + q) / P) ** (1 / 3.) * 1e-3
    np.testing.assert_allclose(K, Kc, rtol=1e-13)
    # in reference units the constant is 213.32 km/s for M1 = 2e30 kg, P = 1 d (SURVEY 7.2 quotes 213.34)
    assert abs((2 * np.pi * orc.G_VALUE * 2e30 / 86400.0) ** (1 / 3.) * 1e-3 - 213.32) < 0.01


def test_rv_at_periastron_and_kepler_residual():
    rng = np.random.RandomState(3)
    n = 5000
    e = rng.uniform(0, 0.9, n)
    om = rng.uniform(0, 2 * np.pi, n)
    m1, P, q, inc = np.full(n, 15 * 2e30), np.full(n, 30 * 86400.0), np.full(n, 0.7), np.full(n, 1.0)
    K = orc.semi_amplitude(q, m1, P, e, inc)
    rv0 = orc.orbital_rv(m1, P, q, e, inc, om, np.zeros(n))            # E = 0: periastron
    np.testing.assert_allclose(rv0, K * (1 + e) * np.cos(om), rtol=1e-12, atol=1e-12)
    M = rng.uniform(0, 2 * np.pi, n)
    E, conv, nev = orc.secant_array(orc.anomaly_function, M, (e, M), 1e-6)
    assert conv.all() and nev <= 50
    assert np.max(np.abs(E - e * np.sin(E) - M)) < 1e-9               # all lanes iterate until the worst meets tol
    assert np.max(np.abs(E - orc.kepler_exact(M, e))) < 1e-9
    # closed-form true anomaly == the reference's 2 atan(sqrt((1+e)/(1-e)) tan(E/2)) form
    nu = 2 * np.arctan(np.sqrt((1 + e) / (1 - e)) * np.tan(0.5 * E))
    d = 1 - e * np.cos(E)
    np.testing.assert_allclose(np.cos(nu), (np.cos(E) - e) / d, atol=1e-9)
    np.testing.assert_allclose(np.sin(nu), np.sqrt(1 - e * e) * np.sin(E) / d, atol=1e-9)


def test_isotropic_inclination_and_dispersion_formulas():
    u = (np.arange(200000) + 0.5) / 200000
    i = orc.inclination_from_uniform(u)
    assert abs(np.mean(np.sin(i)) - np.pi / 4) < 1e-5                  # <sin i> = pi/4
    rng = np.random.RandomState(4)
    rvs = rng.normal(0, 10, (300, 12))
    err = rng.uniform(0.5, 3, 12)
    np.testing.assert_allclose(orc.sigma1d_unweighted(rvs, 1), np.std(rvs, axis=1, ddof=1))
    w = 1 / err ** 2
    mu = (rvs * w).sum(axis=1) / w.sum()
    np.testing.assert_allclose(orc.sigma1d_weighted(rvs, err), np.sqrt((w * (rvs - mu[:, None]) ** 2).sum(axis=1) / w.sum()))
    assert orc.needed_singles(10 ** 5, 0.7) == 42857


def test_detection_rule_is_strict_and_pairwise():
    rvs = np.array([[0.0, 0.0, 0.0], [20.0, 20.0001, 30.0], [0.0, 0.0, 0.0]])
    det = orc.detected_binaries(rvs, 1.0, delta_RV=20, n_sigma_RV=4)
    assert det.tolist() == [False, True, True]                         # strict > on delta_RV
    det = orc.detected_binaries(rvs, 10.0, delta_RV=20, n_sigma_RV=4)  # 30/sqrt(200) = 2.1 sigma
    assert det.tolist() == [False, False, False]
    det = orc.detected_binaries(rvs, [1.0, 10.0, 1.0], delta_RV=20, n_sigma_RV=2.5)
    assert det.tolist() == [False, False, True]                        # 30 / sqrt(101) = 2.99
