"""Pins oracle/rvmc_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/make_golden.py): every logged draw is replayed through the restatement and the
result must equal what the reference computed from the same draw."""
import os

import numpy as np
import pytest

from oracle import rvmc_oracle as orc

RTOL = 1e-12


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    draws = [z[k] for k in sorted(k for k in z.files if k.startswith("draw_"))]
    return z, draws


def _kw(z):
    defaults = dict(min_mass=6., max_mass=20., min_period=10. ** 0.15, max_period=10. ** 3.5,
                    period_pi=-0.5, mass_ratio_kappa=0, mass_power=-2.35, min_q=0.1, max_q=1.0,
                    e_power=-0.5, sigma_dyn=2.0, fbin=0.7)
    for k in z.files:
        if k.startswith("kw_") and k != "kw_names":
            defaults[k[3:]] = float(z[k])
    return defaults


@pytest.mark.parametrize("name", ["rvmc_default_n2048", "rvmc_variant_n1536"])
def test_rvmc_ctor_and_rv_match_reference(golden_dir, name):
    z, d = _load(golden_dir, name)
    kw = _kw(z)
    m1 = orc.masses_from_uniform(d[0], kw["min_mass"], kw["max_mass"], kw["mass_power"])
    period = orc.period_from_uniform(d[1], kw["min_period"], kw["max_period"], kw["period_pi"])
    time = orc.time_from_uniform(d[2], period)
    q = orc.mass_ratio_from_uniform(d[3], kw["min_q"], kw["max_q"], kw["mass_ratio_kappa"])
    ecc = orc.eccentricity_from_uniform_compact(period, d[4], d[5], kw["e_power"])
    E = orc.eccentric_anomaly_single_epoch(time, period, ecc)
    inc = orc.inclination_from_uniform(d[6])
    om = orc.orbit_rotation_from_uniform(d[7])
    for got, key in ((m1, "primary_mass"), (period, "period"), (time, "time"), (q, "mass_ratio"),
                     (ecc, "eccentricity"), (inc, "inclination"), (om, "orbit_rotation")):
        np.testing.assert_array_equal(got, z[key], err_msg=key)      # same NumPy expression
    np.testing.assert_array_equal(d[8], z["cluster_velocities"])
    # restated secant vs scipy's: identical iteration => identical iterates
    np.testing.assert_allclose(E, z["eccentric_anomaly"], rtol=0, atol=1e-14)
    # the secant root is the Kepler root to <1e-9 (all lanes iterate until the worst meets tol=1e-6)
    M = 2 * np.pi * time / period
    assert np.max(np.abs(orc.kepler_exact(M, ecc) - z["eccentric_anomaly"])) < 1e-9
    a = orc.semi_major_axis(q, m1, period)
    np.testing.assert_allclose(a, z["semi_major_axis"], rtol=RTOL)
    vmax = orc.max_orbital_velocity(q, m1, ecc, a)
    np.testing.assert_allclose(vmax, z["v_max"], rtol=RTOL)
    rv = orc.binary_radial_velocity(vmax, ecc, inc, om, z["eccentric_anomaly"])
    np.testing.assert_allclose(rv, z["rv_orbit"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(rv + d[9], z["rv_binary"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("name", ["rvmc_default_n2048", "rvmc_variant_n1536"])
def test_rvmc_mix_and_subsample_match_reference(golden_dir, name):
    z, d = _load(golden_dir, name)
    kw = _kw(z)
    pool = orc.mix_radial_velocities(z["rv_binary"], z["cluster_velocities"], kw["fbin"], d[10], d[11])
    np.testing.assert_array_equal(pool, z["radial_velocities"])
    # subsample() re-mixes first (RVMC.py:259): draws 12,13 ; then choice 14 + normal 15
    pool2 = orc.mix_radial_velocities(z["rv_binary"], z["cluster_velocities"], kw["fbin"], d[12], d[13])
    sig = orc.sigma1d_unweighted(pool2[d[14]] + d[15], ddof=1)
    np.testing.assert_allclose(sig, z["sigma_1d_unweighted"], rtol=RTOL)
    pool3 = orc.mix_radial_velocities(z["rv_binary"], z["cluster_velocities"], kw["fbin"], d[16], d[17])
    sigw = orc.sigma1d_weighted(pool3[d[18]] + d[19], z["m17_errors"])
    np.testing.assert_allclose(sigw, z["sigma_1d_weighted"], rtol=RTOL)


def _biascorr_common(z, d, first_e_draw, nstars, kw):
    period = z["period"]
    ecc = np.zeros(period.shape)
    for i in range(nstars):
        ecc[i] = orc.eccentricity_from_uniform_compact(period[i], d[first_e_draw + 2 * i],
                                                       d[first_e_draw + 2 * i + 1], kw["e_power"])
    np.testing.assert_array_equal(ecc, z["eccentricity"])
    k = first_e_draw + 2 * nstars
    np.testing.assert_array_equal(orc.inclination_from_uniform(d[k]), z["inclination"])
    np.testing.assert_array_equal(orc.orbit_rotation_from_uniform(d[k + 1]), z["orbit_rotation"])
    return k + 2


@pytest.mark.parametrize("name,mode", [("biascorr_masses_ragged", "masses"), ("biascorr_imf_6epochs", "imf")])
def test_biascorr_matches_reference(golden_dir, name, mode):
    z, d = _load(golden_dir, name)
    nstars, nsamples = int(z["nstars"]), int(z["nsamples"])
    kw = dict(period_pi=-0.5, mass_ratio_kappa=0, e_power=-0.5)
    for k in kw:
        if "kw_" + k in z.files:
            kw[k] = float(z["kw_" + k])
    if mode == "masses":
        np.testing.assert_array_equal(d[0], z["primary_mass"])       # normal(masses, errors)
    else:
        np.testing.assert_array_equal(orc.masses_from_uniform(d[0], 6, 20, -2.35), z["primary_mass"])
    period = orc.period_from_uniform(d[1], 10 ** 0.15, 10 ** 3.5, kw["period_pi"])
    np.testing.assert_array_equal(period, z["period"])
    np.testing.assert_array_equal(orc.time_from_uniform(d[2], period), z["time"])
    np.testing.assert_array_equal(orc.mass_ratio_from_uniform(d[3], 0.1, 1.0, kw["mass_ratio_kappa"]),
                                  z["mass_ratio"])
    nxt = _biascorr_common(z, d, 4, nstars, kw)
    assert nxt == int(z["ctor_draws"])
    det = np.zeros((nstars, nsamples), dtype=bool)
    for i in range(nstars):
        rv0 = orc.multi_epoch_rv(z["t_obs_%d" % i], z["time"][i], z["period"][i], z["eccentricity"][i],
                                 z["mass_ratio"][i], z["primary_mass"][i], z["inclination"][i],
                                 z["orbit_rotation"][i])
        rv = rv0 + d[nxt + i]
        np.testing.assert_allclose(rv, z["radial_velocities_%d" % i], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(orc.delta_rv(rv), z["delta_rv"][i], rtol=1e-11, atol=1e-11)
        err = z["rv_errors_%d" % i] if mode == "masses" else float(z["rv_errors_scalar"])
        det[i] = orc.detected_binaries(z["radial_velocities_%d" % i], err, 20, 4)
        single = d[nxt + nstars + i]
        np.testing.assert_array_equal(single.max(axis=0) - single.min(axis=0), z["delta_rv_single"][i])
    np.testing.assert_array_equal(det, z["detected"])
    assert z["delta_rv_single"].shape[1] == orc.needed_singles(nsamples, float(z["fbin"]))
    np.testing.assert_array_equal(
        np.concatenate([z["delta_rv"], z["delta_rv_single"]], axis=1).ravel(), z["all_delta_rv"])


def test_quirk_q1_mean_anomaly_range(golden_dir):
    """RVMC_bias_corr.py:357 yields M in [0,1) 'radians' and E in [0, ~1.86]."""
    z, _ = _load(golden_dir, "biascorr_imf_6epochs")
    t = z["t_obs_0"].reshape(-1, 1) + z["time"][0]
    M = orc.mean_anomaly_multi_epoch(t, z["period"][0])
    assert M.min() >= 0 and M.max() < 1
    E = orc.eccentric_anomaly_multi_epoch(t, z["period"][0], z["eccentricity"][0])
    assert np.max(np.abs(E - orc.kepler_exact(M, z["eccentricity"][0]))) < 1e-9


def test_distribution_stats_rules():
    """findBestDistributions.py:47-60 semantics (that file cannot run -- Python 2; parity of this
    piece is pinned to NumPy's own histogram/percentile only)."""
    rng = np.random.RandomState(3)
    sig = rng.gamma(4.0, 5.0, 20001)
    st, values, bins = orc.distribution_stats(sig, 0.5)
    s = np.sort(sig)
    for q, key in ((5, "p05"), (16, "p16"), (50, "median"), (84, "p84"), (95, "p95")):
        assert st[key] == pytest.approx(orc.percentile_linear(s, q), rel=1e-15)
        assert orc.percentile_linear(s, q) == np.percentile(sig, q)
    assert bins[0] == 0 and np.allclose(np.diff(bins), 0.5)
    assert st["mode"] == bins[np.argmax(values)] + 0.25


def test_best_fit_indices_first_strict_minimum():
    med = [30.0, 20.0, 10.0, 10.0, 5.0]
    flat = [200.0] * 5                       # never closer than the initial 100 -> index stays 0
    idx = orc.best_fit_indices(10.0, med, med, flat, med, med, med)
    assert idx == (2, 0, 2, 2, 2)


def test_oracle_distributions_vs_five_independent_reference_runs(golden_dir):
    """The oracle as a MODEL (not only as arithmetic on recorded draws): driven by NumPy's own generator it
    must produce the sigma_1D and Delta-RV DISTRIBUTIONS the unmodified reference produces
    (tests/golden/reference_multiseed.npz: five independent reference runs, 10^7-star pools -- the infinite-pool
    limit the fused GPU kernels implement).  Two-sample KS, Bonferroni over the 15 comparisons (SURVEY H2 iii)."""
    from scipy import stats as sps
    z = np.load(os.path.join(golden_dir, "reference_multiseed.npz"), allow_pickle=False)
    m17 = z["m17_errors"]
    days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
    pvals = []
    for tag, fbin in (("fbin070_pmin14", 0.7), ("fbin028_pmin14", 0.28)):
        rng = np.random.RandomState(12345)
        n_real, S = 20000, 12
        n = n_real * S
        m1 = orc.masses_from_uniform(rng.random_sample(n), 6., 20., -2.35)
        per = orc.period_from_uniform(rng.random_sample(n), 1.4, 3500., -0.5)
        t = orc.time_from_uniform(rng.random_sample(n), per)
        q = orc.mass_ratio_from_uniform(rng.random_sample(n), 0.1, 1.0, 0)
        e = orc.eccentricity_from_uniform_slotted(per, rng.random_sample(n), -0.5)
        inc = orc.inclination_from_uniform(rng.random_sample(n))
        om = orc.orbit_rotation_from_uniform(rng.random_sample(n))
        E = orc.eccentric_anomaly_single_epoch(t, per, e)
        rv = orc.orbital_rv(m1, per, q, e, inc, om, E)
        is_bin = rng.random_sample(n) < fbin
        v = np.where(is_bin, rv, 0.0) + rng.normal(0, 2.0, n) + rng.normal(0, np.tile(m17, n_real), n)
        ours = orc.sigma1d_unweighted(v.reshape(n_real, S), ddof=1)
        for k in range(z["sigma_" + tag].shape[0]):
            pvals.append(sps.ks_2samp(ours, z["sigma_" + tag][k].astype(np.float64)).pvalue)
    rng = np.random.RandomState(777)
    ns = 20000
    m1 = orc.masses_from_uniform(rng.random_sample(ns), 6., 20., -2.35)
    per = orc.period_from_uniform(rng.random_sample(ns), 10 ** 0.15, 10 ** 3.5, -0.5)
    t = orc.time_from_uniform(rng.random_sample(ns), per)
    q = orc.mass_ratio_from_uniform(rng.random_sample(ns), 0.1, 1.0, 0)
    e = orc.eccentricity_from_uniform_slotted(per, rng.random_sample(ns), -0.5)
    inc = orc.inclination_from_uniform(rng.random_sample(ns))
    om = orc.orbit_rotation_from_uniform(rng.random_sample(ns))
    rvs = orc.multi_epoch_rv(days, t, per, e, q, m1, inc, om) + rng.normal(0, 2.0, (6, ns))
    drv = orc.delta_rv(rvs)
    frac = orc.detected_binaries(rvs, 2.0, 20, 4).mean()
    for k in range(z["biascorr_delta_rv"].shape[0]):
        pvals.append(sps.ks_2samp(drv, z["biascorr_delta_rv"][k].astype(np.float64)).pvalue)
    ps = np.array(pvals)
    assert len(ps) == 15 and ps.min() > 0.01 / len(ps), np.sort(ps)[:3]
    assert np.median(ps) > 0.1, np.sort(ps)
    p_ref = float(np.mean(z["biascorr_detected_fraction"]))
    se = np.sqrt(p_ref * (1 - p_ref) * (1 / 50000 + 1 / ns))
    assert abs(frac - p_ref) < 4 * se
