"""GPU parity, part 4: the grid search (K6) and the drop-in Python classes."""
import ctypes as C
import os
import pickle
import warnings

import numpy as np
import pytest
from scipy import stats as sps

from oracle import rvmc_oracle as orc
from rvmc_b200 import _abi, grid
from rvmc_b200 import _device as dv
from rvmc_b200.bias_corr import BiasCorr
from rvmc_b200.rvmc import RVMC
from tests import gpu_util as g

pytestmark = pytest.mark.gpu
M17 = grid.M17_ERRORS
DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400


# ---- K6 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_real,S", [(10 ** 4, 12), (10007, 5), (1000, 44), (2, 12), (1, 12)])
def test_grid_stats_match_numpy_rules_on_the_same_sigmas(n_real, S):
    """findBestDistributions.py:47-60: mode / mean / median / percentiles computed on the device
    must equal NumPy's on the very sigma values the kernel produced."""
    pops = [_abi.PopParams.defaults(fbin=f, min_period=p) for f in (0.0, 0.3, 1.0) for p in (1.4, 30.0, 2000.0)]
    seeds = grid.point_seeds(5, len(pops))
    res = grid.run_points(pops, seeds, np.resize(M17, S), S, n_real, bin=0.5, ddof=0, keep_sigma=True, keep_hist=True,
                          hist_bins=512, distributed=False)
    for i, st in enumerate(res["stats"]):
        sig = res["sigma"][i].astype(np.float64)
        want, values, bins = orc.distribution_stats(sig, 0.5)
        for k in ("median", "p05", "p16", "p84", "p95"):
            assert st[k] == want[k], (i, k, st[k], want[k])
        assert st["mean"] == pytest.approx(want["mean"], rel=1e-12)
        assert st["mode"] == want["mode"]
        assert st["sigma_max"] == sig.max() and st["n_hist_bins"] == len(bins) - 1
        cnt, _ = np.histogram(sig, bins=bins)
        np.testing.assert_array_equal(res["hist"][i][:len(cnt)], cnt)
        assert res["hist"][i][len(cnt):].sum() == 0
    # the same points one by one through the single-population entry give the same sigmas
    out = g.empty((n_real,), "float32")
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(pops[4]), int(seeds[4]), dv.ptr(g.up(np.resize(M17, S), "float32")), S, 0, 0, 0,
                                    n_real, dv.ptr(out), None, g.stream()))
    np.testing.assert_array_equal(g.host(out), res["sigma"][4])


def test_grid_scratch_path_equals_sigma_path_and_batches():
    """Without sigma_out the points go through the scratch buffer in batches of
    rvmc_grid_scratch_slots(); statistics must not depend on that."""
    nb = g.lib().rvmc_grid_scratch_slots(0)
    pops = [_abi.PopParams.defaults(fbin=f) for f in np.linspace(0.05, 1, nb + 37)]
    seeds = grid.point_seeds(77, len(pops))
    a = grid.run_points(pops, seeds, M17, 12, 2000, keep_sigma=True, distributed=False)
    b = grid.run_points(pops, seeds, M17, 12, 2000, keep_sigma=False, distributed=False)
    for k in ("mode", "mean", "median", "p05", "p16", "p84", "p95", "sigma_max"):
        np.testing.assert_array_equal(a["stats"][k], b["stats"][k])
    np.testing.assert_array_equal(a["counters"], b["counters"])
    # median rises with the binary fraction
    assert a["stats"]["median"][-1] > 3 * a["stats"]["median"][0]
    # duplicates everywhere: fbin = 0, no noise -> all sigmas are 0 and every statistic is 0
    z = grid.run_points([_abi.PopParams.defaults(fbin=0.0, sigma_dyn=0.0)], [1], np.zeros(12), 12, 5000,
                        distributed=False)
    st = z["stats"][0]
    assert st["median"] == 0 and st["p95"] == 0 and st["mean"] == 0 and st["n_hist_bins"] == 0 and np.isnan(st["mode"])


@pytest.mark.parametrize("S,weighted,ddof,n_real", [(12, 0, 0, 5000), (5, 0, 1, 3001), (44, 1, 0, 700), (100, 0, 0, 300),
                                                   (2048, 0, 1, 5)])
def test_fbin_axis_in_one_pass_is_bit_identical_to_pointwise(S, weighted, ddof, n_real):
    """rvmc_grid_run_fbins: every binary fraction of a point from ONE evaluation of its star slots,
    bit-identical to rvmc_grid_run on the expanded point list with the seed shared along f_bin."""
    pops = grid.points_array(5, period_pi=np.linspace(-0.9, 0.3, 5), mass_ratio_kappa=np.linspace(0.8, -0.8, 5),
                             min_period=[1.4, 1.4, 30.0, 1.4, 3.0], sigma_dyn=[2.0, 0.0, 2.0, 5.0, 2.0])
    seeds = grid.point_seeds(2027, 5)
    fbins = np.array([0.7, 0.0, 1.0, 0.12, 0.999, 0.5])                       # unsorted on purpose
    err = np.resize(M17, S)
    a = grid.run_points_fbins(pops, seeds, fbins, err, S, n_real, ddof=ddof, weighted=bool(weighted), keep_sigma=True,
                              distributed=False)
    exp = np.repeat(pops, len(fbins))
    exp["fbin"] = np.tile(fbins, 5)
    b = grid.run_points(exp, np.repeat(seeds, len(fbins)), err, S, n_real, ddof=ddof, weighted=bool(weighted),
                        keep_sigma=True, distributed=False)
    np.testing.assert_array_equal(a["sigma"].reshape(-1, n_real), b["sigma"])
    for k in grid.STATS_DTYPE.names:
        if k not in ("n_guard", "n_nonconverged", "reserved"):
            np.testing.assert_array_equal(a["stats"][k].ravel(), b["stats"][k], err_msg=k)
    # far fewer orbits are solved: once per slot under the largest fraction (1.0 here)
    assert a["counters"][0] == 5 * n_real * S and b["counters"][0] > 3 * a["counters"][0]
    # more than 32 fractions go out in chunks of the same draws
    many = np.linspace(0, 1, 40)
    c = grid.run_points_fbins(pops[:2], seeds[:2], many, err, S, min(n_real, 500), ddof=ddof, weighted=bool(weighted),
                              distributed=False)
    assert c["stats"].shape == (2, 40) and np.all(np.diff(c["stats"]["mean"], axis=1) >= -0.35)


@pytest.mark.gpu
@pytest.mark.parametrize("S,nf,n_real", [(12, 16, 1000), (12, 11, 700), (7, 24, 513), (3, 32, 2000), (16, 8, 300)])
def test_fbin_axis_fraction_groups_and_tile_sizes(S, nf, n_real):
    """Small clusters reduce one (realization, group of eight fractions) item per thread and the host
    sizes the tile from (S, number of groups): every combination -- a partial last group, partial
    last tiles, two to four groups -- must stay bit-identical to point-by-point evaluation."""
    pops = grid.points_array(3, period_pi=[-0.5, 0.2, -0.9], mass_ratio_kappa=[0.0, 0.5, -0.5])
    seeds = grid.point_seeds(99, 3)
    fbins = np.linspace(0.03, 1.0, nf)
    err = np.resize(M17, S)
    a = grid.run_points_fbins(pops, seeds, fbins, err, S, n_real, keep_sigma=True, distributed=False)
    exp = np.repeat(pops, nf)
    exp["fbin"] = np.tile(fbins, 3)
    b = grid.run_points(exp, np.repeat(seeds, nf), err, S, n_real, keep_sigma=True, distributed=False)
    np.testing.assert_array_equal(a["sigma"].reshape(-1, n_real), b["sigma"])
    for k in ("median", "mode", "mean", "p05", "p95"):
        np.testing.assert_array_equal(a["stats"][k].ravel(), b["stats"][k], err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("S,n_real", [(12, 1500), (40, 257)])
def test_shared_seed_with_substreams_equals_per_point_keys(S, n_real):
    """Grid points under one seed, told apart by their substream (the Philox counter word), run kernels
    that read one launch-uniform key schedule.  Appending a point with a different seed forces the
    per-point key schedule for the same points: identical bits, in both grid entry points; the
    single-population entry with the same (seed, substream) agrees too, and substreams differ."""
    pops = grid.with_streams(grid.points_array(4, period_pi=[-0.5, 0.1, -0.8, -0.5], fbin=[0.7, 0.3, 1.0, 0.7]), 4711)
    err = np.resize(M17, S)
    keyed = grid.run_points(pops, None, err, S, n_real, keep_sigma=True, distributed=False)
    extra = np.concatenate([pops, grid.with_streams(grid.points_array(1), 1234)])
    unkeyed = grid.run_points(extra, None, err, S, n_real, keep_sigma=True, distributed=False)
    np.testing.assert_array_equal(keyed["sigma"], unkeyed["sigma"][:4])
    fb = np.array([0.2, 0.7, 1.0])
    kf = grid.run_points_fbins(pops, None, fb, err, S, n_real, keep_sigma=True, distributed=False)
    uf = grid.run_points_fbins(extra, None, fb, err, S, n_real, keep_sigma=True, distributed=False)
    np.testing.assert_array_equal(kf["sigma"], uf["sigma"][:4])
    np.testing.assert_array_equal(kf["sigma"][0, 1], keyed["sigma"][0])            # fbin 0.7 of point 0
    # points 0 and 3 have the same population and seed but different substreams: different draws, same law
    assert not np.array_equal(keyed["sigma"][0], keyed["sigma"][3])
    assert abs(np.median(keyed["sigma"][0]) / np.median(keyed["sigma"][3]) - 1) < 0.1
    one = g.empty((n_real,), "float32")
    p = _abi.PopParams.defaults(period_pi=0.1, fbin=0.3, substream=1)
    g.ok(g.lib().rvmc_sigma1d_fused(C.byref(p), 4711, dv.ptr(g.up(err, "float32")), S, 0, 0, 0, n_real, dv.ptr(one),
                                    None, g.stream()))
    np.testing.assert_array_equal(g.host(one), keyed["sigma"][1])


def test_reference_entry_points_and_best_fit(tmp_path):
    t = grid.get_Pdistr_stats(measured_errors=M17, sample_size=12, pmin=1.4, pmax=3500, Npoints=12, bin=0.5, fbin=0.7,
                              min_mass=6, max_mass=20, number_of_samples=20000, seed=3)
    allSig, mode, mean, median, values, bins, p05, p16, p84, p95, periods = t
    assert len(allSig) == 12 and allSig[0].shape == (20000,)
    np.testing.assert_allclose(periods, np.logspace(np.log10(1.4), np.log10(3500), 12))
    assert median[0] > median[-1]                               # dispersion falls as P_min grows
    for i in range(12):
        want, v, b = orc.distribution_stats(allSig[i], 0.5)
        assert median[i] == want["median"] and mode[i] == want["mode"]
        np.testing.assert_allclose(values[i], v, rtol=1e-12)
        np.testing.assert_array_equal(bins[i], b)
    obs = float(median[5])
    row = grid.find_best_Pmin(obs, 6, 20, 12, mode, median, p05, p16, p84, p95, bins, values, periods,
                              outdir=str(tmp_path))
    want_idx = orc.best_fit_indices(obs, mode, median, p05, p16, p84, p95)
    assert row == [periods[i] for i in want_idx] and row[0] == periods[5]
    txt = open(tmp_path / ("Pmin_ObsSig_%s_Nstars_12_Mass_6_20.dat" % obs)).read().split("\n")
    assert txt[0].split() == ['best_Pmin', '0.05perc_Pmin', '0.16perc_Pmin', '0.84perc_Pmin', '0.95perc_Pmin']
    assert [float(x) for x in txt[1].split()] == row
    t = grid.get_fbinDist_stats(measured_errors=M17, sample_size=12, Npoints=11, number_of_samples=20000, seed=4,
                                keep_sigma=False)
    _, mode, mean, median, values, bins, p05, p16, p84, p95, fbins = t
    assert grid.find_best_fbin(float(median[7]), 6, 20, 12, mode, median, p05, p16, p84, p95, bins, values, fbins,
                               outdir=str(tmp_path)) == float(median[7])
    assert grid.find_best_fbin.last[0] == fbins[7]
    sig = grid.create_sig1D_distribution(measured_errors=M17, number_of_samples=20000, seed=8)
    sig_pool = grid.create_sig1D_distribution(measured_errors=M17, number_of_samples=20000, seed=9, pool=True,
                                              number_of_stars=10 ** 6)
    assert sps.ks_2samp(sig, sig_pool).pvalue > 1e-3            # infinite pool ~ a 10^6-star pool
    l_res = grid.run_bigGrid(number_stars=8, min_mass=15, max_mass=60, n_fbin=5, n_pmin=6, number_of_samples=5000,
                             seed=1, outdir=str(tmp_path))
    back = pickle.load(open(tmp_path / "grid_Pmin_fbin_Nstars_8_Mass_15_60.pkl", "rb"))
    assert len(back) == 5 and sorted(back[0]) == sorted(['values', 'bins', 'period', 'fbin', 'mean', 'median', 'mode',
                                                         'p05', 'p16', 'p84', 'p95'])
    assert back[4]['median'][0] > back[0]['median'][0] and len(back[0]['period']) == 6


def test_grid_search_recovers_known_parameters():
    """BASELINE config 5 in miniature: the generalised (pi, kappa, fbin) grid; the observed sigma is
    the median of one interior point, the best fit must land on it (common random numbers)."""
    axes = dict(period_pi=np.linspace(-0.9, 0.4, 4), mass_ratio_kappa=np.linspace(-0.9, 0.9, 3),
                fbin=np.linspace(0.25, 1.0, 4))
    r = grid.grid_search(axes, number_of_samples=20000, seed=11, common_random_numbers=True)
    assert r["shape"] == (4, 3, 4)
    target = (1, 2, 2)
    r2 = grid.grid_search(axes, number_of_samples=20000, seed=11, common_random_numbers=True,
                          observed_sigma=float(r["stats"]["median"][target]))
    assert r2["best_index"] == target and r2["best_distance"] == 0.0
    # shallower period law (more long periods) lowers the dispersion; more binaries raise it
    med = r["stats"]["median"]
    assert np.all(np.diff(med[:, 1, 3]) < 0) and np.all(np.diff(med[2, 1, :]) > 0)


@pytest.mark.parametrize("bin", [0.5, 0.3, 2.0])
def test_fast_statistics_kernel_equals_the_general_one_and_numpy(bin, monkeypatch):
    """point_stats_fast_kernel (16-bit prefix histogram + candidate lists, two reads) against the general
    three-pass kernel (RVMC_STATS_GENERAL=1) and NumPy on the same sigma arrays: identical order statistics,
    mode and histogram; means equal to rounding.  Rows the fast form cannot take (all zeros: the ranks fall
    outside its window) go through the general kernel inside the same call."""
    pops = [_abi.PopParams.defaults(fbin=f, min_period=p) for f in (0.1, 0.7, 1.0) for p in (1.4, 300.0)]
    pops.append(_abi.PopParams.defaults(fbin=0.0, sigma_dyn=0.0))               # all-zero sigma row
    pops.append(_abi.PopParams.defaults(fbin=0.02, sigma_dyn=0.0))              # mostly zeros: low ranks outside the window
    seeds = list(range(40, 40 + len(pops)))
    errs = np.zeros(12)
    errs_noisy = M17
    for err, n_real in ((errs_noisy, 100003), (errs, 20000), (errs_noisy, 7)):
        use = pops if err is errs else pops[:6]
        monkeypatch.delenv("RVMC_STATS_GENERAL", raising=False)
        fast = grid.run_points(use, seeds[:len(use)], err, 12, n_real, bin=bin, keep_sigma=True, keep_hist=True,
                               hist_bins=1024, distributed=False)
        monkeypatch.setenv("RVMC_STATS_GENERAL", "1")
        gen = grid.run_points(use, seeds[:len(use)], err, 12, n_real, bin=bin, keep_sigma=True, keep_hist=True,
                              hist_bins=1024, distributed=False)
        monkeypatch.delenv("RVMC_STATS_GENERAL", raising=False)
        np.testing.assert_array_equal(fast["sigma"], gen["sigma"])
        for k in grid.STATS_DTYPE.names:
            if k == "mean":
                np.testing.assert_allclose(fast["stats"][k], gen["stats"][k], rtol=1e-13, atol=0)
            else:
                np.testing.assert_array_equal(fast["stats"][k], gen["stats"][k], err_msg=k)
        np.testing.assert_array_equal(fast["hist"], gen["hist"])
        for i in range(len(use)):
            sig = fast["sigma"][i].astype(np.float64)
            for q, k in ((5, "p05"), (16, "p16"), (50, "median"), (84, "p84"), (95, "p95")):
                assert fast["stats"][k][i] == np.percentile(sig, q), (i, k)
            if sig.max() > 0:
                edges = np.arange(0, sig.max() + bin, bin)
                cnt, _ = np.histogram(sig, bins=edges)
                nb = len(edges) - 1
                assert fast["stats"]["n_hist_bins"][i] == nb
                np.testing.assert_array_equal(fast["hist"][i][:min(nb, 1024)], cnt[:1024])
                assert fast["stats"]["mode"][i] == edges[int(np.argmax(cnt))] + bin / 2


def test_fast_statistics_kernel_hands_over_rows_with_crowded_target_bins(monkeypatch):
    """With 1.5e6 realizations per point the 16-bit-prefix bins around the percentiles hold more values than
    the fast kernel's candidate list (16 384): it must hand the row to the general kernel, and the call must
    return the same statistics as the general kernel alone and as NumPy."""
    pops = [_abi.PopParams.defaults(fbin=0.7), _abi.PopParams.defaults(fbin=0.3, min_period=30.0)]
    n_real = 1_500_000
    monkeypatch.delenv("RVMC_STATS_GENERAL", raising=False)
    fast = grid.run_points(pops, [5, 6], M17, 12, n_real, keep_sigma=True, distributed=False)
    monkeypatch.setenv("RVMC_STATS_GENERAL", "1")
    gen = grid.run_points(pops, [5, 6], M17, 12, n_real, keep_sigma=True, distributed=False)
    monkeypatch.delenv("RVMC_STATS_GENERAL", raising=False)
    for k in grid.STATS_DTYPE.names:
        if k == "mean":
            np.testing.assert_allclose(fast["stats"][k], gen["stats"][k], rtol=1e-13, atol=0)
        else:
            np.testing.assert_array_equal(fast["stats"][k], gen["stats"][k], err_msg=k)
    for i in range(2):
        sig = fast["sigma"][i].astype(np.float64)
        assert fast["stats"]["median"][i] == np.median(sig) and fast["stats"]["p95"][i] == np.percentile(sig, 95)
        assert abs(fast["stats"]["mean"][i] / sig.mean() - 1) < 1e-12


def test_grid_per_point_failure_counters_and_best_fit_ignores_nan():
    """rvmc_point_stats.n_guard / n_nonconverged are filled per point by the producer kernels (what
    scipy.optimize.newton reports per call, RVMC.py:177-180): with eccentricities up to 0.999 the fp64
    guard engages, more often the more binaries a point has; the per-point counts add up to the call's
    counters, along the fused f_bin axis too.  An all-zero sigma_1D point (mode NaN) never wins the best
    fit (np.argmin returned the first NaN)."""
    pops = [_abi.PopParams.defaults(e_max_long=0.999, fbin=f) for f in (0.2, 0.9)] + [_abi.PopParams.defaults()]
    res = grid.run_points(pops, [3, 4, 5], M17, 12, 20000, distributed=False)
    st = res["stats"]
    assert st["n_guard"][0] > 0 and st["n_guard"][1] > 2 * st["n_guard"][0] and st["n_guard"][2] == 0
    assert int(st["n_guard"].sum()) == int(res["counters"][1])
    assert int(st["n_nonconverged"].sum()) == int(res["counters"][2]) == 0
    fb = np.array([0.2, 1.0, 0.6])
    rf = grid.run_points_fbins(pops[:1], [3], fb, M17, 12, 20000, distributed=False)
    ng = rf["stats"]["n_guard"][0]
    assert ng[1] == int(rf["counters"][1]) and ng[0] < ng[2] < ng[1] and ng[0] > 0
    # bit-identity with the pointwise evaluation extends to the counters
    exp = [_abi.PopParams.defaults(e_max_long=0.999, fbin=float(f)) for f in fb]
    rp = grid.run_points(exp, [3, 3, 3], M17, 12, 20000, distributed=False)
    np.testing.assert_array_equal(rp["stats"]["n_guard"], ng)
    np.testing.assert_array_equal(rp["stats"]["median"], rf["stats"]["median"][0])
    # NaN statistics are skipped by the best-fit rule (first strict minimum, findBestDistributions.py:138-158)
    axes = dict(sigma_dyn=np.array([0.0, 2.0]), fbin=np.array([0.0, 0.5]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = grid.grid_search(axes, measured_errors=np.zeros(12), number_of_samples=2000, seed=3, observed_sigma=3.0,
                             usemode=True, share_fbin_draws=False)
    assert np.isnan(r["stats"]["mode"][0, 0])                 # sigma_dyn = 0, no binaries, no errors: all zeros
    assert r["best_index"] != (0, 0) and np.isfinite(r["best_distance"])


# ---- drop-in classes --------------------------------------------------------------------------------
def test_rvmc_class_walkthrough_like_the_example_script():
    """RVMC_example.py:6-37 on the GPU."""
    c = RVMC(number_of_stars=10 ** 5, seed=1234)
    assert c.period.shape == (10 ** 5,) and c.period.dtype == np.float64
    assert c.period.min() >= 10 ** 0.15 * 86400 * (1 - 1e-12) and c.period.max() <= 10 ** 3.5 * 86400 * (1 + 1e-12)
    assert np.all(c.eccentricity[c.period <= 4 * 86400] == 0) and c.eccentricity.max() <= 0.9
    assert abs(np.mean(np.sin(c.inclination)) - np.pi / 4) < 0.005
    assert c.rv_binary is None and c.sigma_1d is None and np.all(c.radial_velocities == 0)
    # reference formulas on the class's own arrays
    E = orc.eccentric_anomaly_single_epoch(c.time, c.period, c.eccentricity)
    assert np.max(np.abs(c.eccentric_anomaly - E)) < 1e-9
    a = c.calculate_semi_major_axis()
    vm = c.max_orbital_velocity(a)
    np.testing.assert_allclose(a, orc.semi_major_axis(c.mass_ratio, c.primary_mass, c.period), rtol=1e-13)
    rv0 = c.binary_radial_velocity(vm)
    want = orc.orbital_rv(c.primary_mass, c.period, c.mass_ratio, c.eccentricity, c.inclination, c.orbit_rotation,
                          c.eccentric_anomaly)
    K = orc.semi_amplitude(c.mass_ratio, c.primary_mass, c.period, c.eccentricity, c.inclination)
    assert g.rel_err(rv0, want, K).max() < 1e-9
    # "the distributions can be changed if you wish" (RVMC_example.py:11-14)
    c.time = np.zeros(10 ** 5)
    c.eccentric_anomaly = c.calculate_eccentric_anomaly()
    assert np.all(c.eccentric_anomaly == 0)
    rvb = c.calculate_binary_velocities()
    peri = orc.orbital_rv(c.primary_mass, c.period, c.mass_ratio, c.eccentricity, c.inclination, c.orbit_rotation,
                          np.zeros(10 ** 5))
    resid = rvb - peri                                   # = the N(0, sigma_dyn) term of RVMC.py:224
    assert abs(resid.std() - 2.0) < 0.03 and abs(resid.mean()) < 0.03
    c.calculate_radial_velocity()
    pool = c.radial_velocities
    nb = int(10 ** 5 * 0.7)
    assert np.all(np.isin(pool[:nb], rvb)) and len(np.unique(pool[:nb])) == nb
    assert np.all(np.isin(pool[nb:], c.cluster_velocities))
    s10 = c.subsample(sample_size=10, number_of_samples=10 ** 5)
    s100 = c.subsample(sample_size=100, number_of_samples=10 ** 4)
    assert s10.shape == (10 ** 5,) and c.sigma_1d is s100 or np.array_equal(c.sigma_1d, s100)
    assert s100.std() < s10.std()
    # fbin argument is persistent
    c.subsample(sample_size=12, number_of_samples=1000, fbin=0.3)
    assert c.fbin == 0.3


def test_rvmc_subsample_statistics_vs_reference_model():
    """Finite-pool subsample() vs an independent NumPy bootstrap of the SAME pool (KS p > 0.01), and
    the fused infinite-pool limit vs a 10^6 pool."""
    c = RVMC(number_of_stars=10 ** 6, seed=42)
    c.calculate_binary_velocities(materialize=False)
    got = c.subsample(sample_size=12, measured_errors=1.2, number_of_samples=30000)
    pool = c.radial_velocities
    rng = np.random.RandomState(1)
    want = np.std(rng.choice(pool, (30000, 12)) + rng.normal(0, 1.2, (30000, 12)), axis=1, ddof=1)
    assert sps.ks_2samp(got, want).pvalue > 0.01
    gw = c.subsample(sample_size=12, measured_errors=M17, number_of_samples=30000, weighted=True)
    rvs = rng.choice(c.radial_velocities, (30000, 12)) + rng.normal(0, M17, (30000, 12))
    assert sps.ks_2samp(gw, orc.sigma1d_weighted(rvs, M17)).pvalue > 0.01
    # weighted=True with a SCALAR error: documented divergence from RVMC.py:263-268, whose
    # np.sum(scalar weight) makes rv_mean the SUM of the velocities -- here equal weights = the ddof 0 dispersion
    gs = c.subsample(sample_size=12, measured_errors=1.2, number_of_samples=30000, weighted=True)
    want0 = np.std(rng.choice(pool, (30000, 12)) + rng.normal(0, 1.2, (30000, 12)), axis=1, ddof=0)
    assert sps.ks_2samp(gs, want0).pvalue > 0.01
    ref_quirk = np.sqrt(((rvs.T - rvs.sum(axis=1)).T ** 2).sum(axis=1))      # what the reference computes there
    assert np.median(ref_quirk) > 3 * np.median(gs)                           # ... is not a dispersion at all
    fused, cnt = c.subsample_fused(sample_size=12, measured_errors=1.2, number_of_samples=30000, return_counters=True)
    assert sps.ks_2samp(fused.astype(np.float64), want).pvalue > 1e-3
    assert abs(cnt["binary_rvs"] / (30000 * 12) - c.fbin) < 0.01 and cnt["nonconverged"] == 0


def test_rvmc_errors_and_custom_distributions():
    with pytest.raises(ZeroDivisionError):
        RVMC(number_of_stars=100, mass_power=-1)
    c = RVMC(number_of_stars=1000, seed=1)
    with pytest.raises(TypeError):
        c.subsample(measured_errors=1)                     # int error: len() of an int
    with pytest.raises(ValueError):
        c.calculate_radial_velocity()                      # before calculate_binary_velocities
    c.calculate_binary_velocities()
    with pytest.raises(ValueError):
        c.subsample(sample_size=12, measured_errors=np.ones(5))
    with pytest.raises(NotImplementedError):
        c.subsample(sample_size=5000, number_of_samples=10)
    per = np.full(1000, 5.0 * 86400)
    m = np.full(1000, 30 * 2e30)
    c2 = RVMC(number_of_stars=1000, mass_dist=m, period_dist=per, seed=2, mass_power=-1)   # IMF law unused
    assert np.array_equal(c2.period, per) and np.array_equal(c2.primary_mass, m)
    assert np.all((c2.eccentricity >= 1e-5) & (c2.eccentricity <= 0.5)) and np.all(c2.time < per)
    with pytest.raises(ValueError):
        RVMC(number_of_stars=1000, period_dist=per[:10])
    # fresh draws from the sample_* methods; same seed -> same population
    assert not np.array_equal(c.sample_period(), c.sample_period())
    assert np.array_equal(RVMC(number_of_stars=500, seed=5).period, RVMC(number_of_stars=500, seed=5).period)
    pl = c.sample_powerlaw(2.0, 8.0, -2.0, size=20000)
    assert pl.min() >= 2.0 and pl.max() <= 8.0
    cdf = (1 / 2.0 - 1 / np.sort(pl)) / (1 / 2.0 - 1 / 8.0)
    assert sps.kstest(cdf, "uniform").pvalue > 1e-3


def test_rvmc_physics_knobs():
    """the eccentricity regime constants (RVMC.py:155-159) as keyword arguments."""
    c = RVMC(number_of_stars=50000, seed=3, e_max_mid=0.3, e_max_long=0.6, p_circ_days=10.0, p_mid_days=100.0)
    P = c.period / 86400.0
    e = c.eccentricity
    assert np.all(e[P <= 10.0] == 0)
    mid = (P > 10.0) & (P <= 100.0)
    assert e[mid].max() <= 0.3 and e[mid].min() >= 1e-5 and e[P > 100.0].max() <= 0.6 and e[P > 100.0].max() > 0.5
    s = c.subsample_fused(sample_size=12, number_of_samples=2000)
    assert np.isfinite(s).all()
    with pytest.raises(TypeError):
        RVMC(number_of_stars=10, e_maximum=0.5)


def test_biascorr_class_fused_and_replay_modes():
    t_obs = [DAYS6, DAYS6[:4], DAYS6]
    bc = BiasCorr(t_obs, number_of_samples=20000, rv_errors=[2.0, [1.0, 2.0, 1.5, 0.5], 3.0], seed=7)
    assert not bc.get_detected_binaries().any()            # before any RV pass (radial_velocities == [[]]*n)
    bc.calculate_radial_velocities()
    det = bc.get_detected_binaries()
    assert det.shape == (3, 20000) and det.dtype == bool and 0.2 < det.mean() < 0.9
    per_star, counters = bc.detection_counts()
    np.testing.assert_array_equal(per_star, det.sum(axis=1))
    assert counters["binary_epoch_rvs"] == 20000 * 16 and counters["nonconverged"] == 0
    drv = bc.delta_rv
    rvs = bc.radial_velocities                                # regenerated from the same seeds
    assert rvs[1].shape == (4, 20000)
    for i in range(3):
        np.testing.assert_allclose(drv[i], rvs[i].max(axis=0) - rvs[i].min(axis=0), rtol=0, atol=1e-4)
    # other thresholds re-run the fused pass; stricter thresholds detect a subset
    det2 = bc.get_detected_binaries(delta_RV=50, n_sigma_RV=6)
    assert det2.sum() < det.sum() and not (det2 & ~det).any()
    # oracle detection on the class's own noisy RVs
    for i, e in enumerate([2.0, [1.0, 2.0, 1.5, 0.5], 3.0]):
        assert (orc.detected_binaries(rvs[i], e, 20, 4) != det[i]).sum() <= 3
    bc.generate_single_delta_rv()
    assert bc.delta_rv_single.shape == (3, orc.needed_singles(20000, 0.7))
    assert bc.get_all_delta_rv().shape == (3 * (20000 + orc.needed_singles(20000, 0.7)),)
    # reading an attribute switches to the fp64 replay path, which must agree with the fused pass
    P = bc.period
    assert P.shape == (3, 20000)
    bc2 = BiasCorr(t_obs, number_of_samples=20000, rv_errors=[2.0, [1.0, 2.0, 1.5, 0.5], 3.0], seed=7)
    _ = bc2.eccentricity
    bc2.calculate_radial_velocities()
    assert bc2._rv_cache["mode"] == "replay" and bc2.semi_major_axis is not None
    rv2 = bc2.radial_velocities
    K = orc.semi_amplitude(bc2.mass_ratio, bc2.primary_mass, bc2.period, bc2.eccentricity, bc2.inclination)
    for i in range(3):
        assert g.rel_err(rvs[i], rv2[i], np.maximum(K[i], 1.0)).max() < 2e-5
    assert (bc2.get_detected_binaries() != det).sum() <= 6
    # user edits are honoured: zero eccentricity + zero inclination -> no orbital motion -> noise only
    bc2.inclination = np.zeros((3, 20000))
    bc2.calculate_radial_velocities()
    assert bc2.get_detected_binaries().mean() < 1e-3


@pytest.mark.gpu
def test_biascorr_large_pass_in_star_ranges_equals_one_launch():
    """Above 4e6 binaries the class launches the fused kernel over tapered star ranges on two
    alternating streams and copies each range's flags to the host as it completes; the result must be
    the same bits as one launch over all stars."""
    days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0
    t_obs = [days + 3600.0 * i for i in range(9)]
    errs = [1.0 + 0.25 * i for i in range(9)]
    res = []
    saved = BiasCorr._d2h_chunks
    try:
        for chunks in (1, 12, 4):
            BiasCorr._d2h_chunks = chunks
            bc = BiasCorr(t_obs, number_of_samples=500_000, rv_errors=errs, seed=11)
            bc.calculate_radial_velocities()
            det = bc.get_detected_binaries()
            res.append((np.asarray(bc.delta_rv).copy(), det.copy(), bc.detection_counts()))
    finally:
        BiasCorr._d2h_chunks = saved
    for drv, det, (per_star, counters) in res[1:]:
        np.testing.assert_array_equal(drv, res[0][0])
        np.testing.assert_array_equal(det, res[0][1])
        np.testing.assert_array_equal(per_star, res[0][2][0])
        assert counters == res[0][2][1]
    assert res[0][1].shape == (9, 500_000) and res[0][2][1]["binary_epoch_rvs"] == 9 * 500_000 * 6
    np.testing.assert_array_equal(res[0][2][0], res[0][1].sum(axis=1))


@pytest.mark.gpu
def test_biascorr_few_stars_go_out_in_sample_ranges():
    """With only a few stars the large pass is split along the sample axis instead (column blocks of
    the result arrays, Philox counter = sample index): same bits as one launch, ragged epochs included."""
    days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0
    t_obs = [days, days[:4] + 777.0]
    res = []
    saved = BiasCorr._d2h_chunks
    try:
        for chunks in (1, 12, 5):
            BiasCorr._d2h_chunks = chunks
            bc = BiasCorr(t_obs, number_of_samples=2_300_001, rv_errors=[2.0, np.array([1.0, 2.0, 3.0, 0.5])], seed=5)
            bc.calculate_radial_velocities()
            det = bc.get_detected_binaries()
            res.append((np.asarray(bc.delta_rv).copy(), det.copy(), bc.detection_counts()))
    finally:
        BiasCorr._d2h_chunks = saved
    for drv, det, (per_star, counters) in res[1:]:
        np.testing.assert_array_equal(drv, res[0][0])
        np.testing.assert_array_equal(det, res[0][1])
        np.testing.assert_array_equal(per_star, res[0][2][0])
        assert counters == res[0][2][1]
    assert res[0][2][1]["binary_epoch_rvs"] == 2_300_001 * 10
    np.testing.assert_array_equal(res[0][2][0], res[0][1].sum(axis=1))


@pytest.mark.gpu
def test_biascorr_more_epochs_than_the_fused_kernel_holds():
    """A monitoring campaign with more than 64 epochs of one star goes through the materialised fp64
    path (the reference has no limit): Delta-RV and the detection flags must follow from the class's own
    attribute arrays by the reference's formulas."""
    rng = np.random.RandomState(8)
    t_obs = [np.sort(rng.uniform(0, 3e7, 90)), np.array([0.0, 86400.0, 5e6])]
    bc = BiasCorr(t_obs, number_of_samples=3000, rv_errors=[1.5, 2.5], seed=21)
    bc.calculate_radial_velocities()
    det = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
    rvs = bc.radial_velocities
    assert rvs[0].shape == (90, 3000) and rvs[1].shape == (3, 3000) and det.shape == (2, 3000)
    for i, err in enumerate([1.5, 2.5]):
        np.testing.assert_allclose(np.asarray(bc.delta_rv)[i], rvs[i].max(axis=0) - rvs[i].min(axis=0), rtol=1e-12)
        np.testing.assert_array_equal(det[i], orc.detected_binaries(rvs[i], err, 20, 4))
    clean = BiasCorr(t_obs, number_of_samples=3000, seed=21)
    clean.calculate_radial_velocities()
    for i in range(2):
        want = orc.multi_epoch_rv(t_obs[i], clean.time[i], clean.period[i], clean.eccentricity[i], clean.mass_ratio[i],
                                  clean.primary_mass[i], clean.inclination[i], clean.orbit_rotation[i])
        K = orc.semi_amplitude(clean.mass_ratio[i], clean.primary_mass[i], clean.period[i], clean.eccentricity[i],
                               clean.inclination[i])
        assert np.max(np.abs(clean.radial_velocities[i] - want) / np.maximum(np.abs(want), K)) < 1e-9
    assert det[0].mean() > det[1].mean() > 0.05          # 90 epochs catch more binaries than 3


def test_biascorr_constructor_errors_and_modes():
    with pytest.raises(ValueError, match="number of masses"):
        BiasCorr([DAYS6, DAYS6], masses=[1e31])
    with pytest.raises(ValueError, match="shape of mass distributions"):
        BiasCorr([DAYS6], mass_dist=np.ones((2, 5)), number_of_samples=5)
    with pytest.raises(ValueError, match="shape of period distribution"):
        BiasCorr([DAYS6], period_dist=np.ones((1, 6)), number_of_samples=5)
    with pytest.raises(ValueError, match="radial velocity uncertainties"):
        BiasCorr([DAYS6], rv_errors=[1.0, 2.0])
    with pytest.raises(ZeroDivisionError):
        BiasCorr([DAYS6], mass_ratio_kappa=-1)
    bc = BiasCorr([DAYS6], number_of_samples=100, seed=1)       # no rv_errors
    bc.calculate_radial_velocities()
    with pytest.raises(TypeError):
        bc.get_detected_binaries()
    bc.generate_single_delta_rv()
    assert np.all(bc.delta_rv_single == 0)
    with pytest.raises(ValueError, match="does not match the number of epochs"):
        b = BiasCorr([DAYS6], rv_errors=[[1.0, 2.0]], number_of_samples=10)
        b.calculate_radial_velocities()
    with pytest.raises(ZeroDivisionError):
        b = BiasCorr([DAYS6], rv_errors=1.0, number_of_samples=10, fbin=0)
        b.generate_single_delta_rv()
    # masses at face value: primary_mass has shape (nstars, 1) (RVMC_bias_corr.py:190)
    m = np.array([20.0, 40.0]) * 2e30
    b = BiasCorr([DAYS6, DAYS6], masses=m, number_of_samples=1000, rv_errors=1.0, seed=3)
    assert b.primary_mass.shape == (2, 1) and b.min_mass == 20.0 and b.max_mass == 40.0
    b.calculate_radial_velocities()
    assert b.delta_rv.shape == (2, 1000)
    b = BiasCorr([DAYS6, DAYS6], masses=m, mass_errors=m * 0.1, number_of_samples=4000, rv_errors=1.0, seed=3)
    assert 10 < b.min_mass < 20 and 40 < b.max_mass < 70
    per = np.full((1, 50), 10 * 86400.0)
    b = BiasCorr([DAYS6], period_dist=per, number_of_samples=50, rv_errors=1.0, seed=3)
    assert b.min_mass == 10.0 and not hasattr(b, "min_period")
    b.calculate_radial_velocities()
    assert np.all((b.eccentricity >= 1e-5) & (b.eccentricity <= 0.9)) and np.isfinite(b.delta_rv).all()
