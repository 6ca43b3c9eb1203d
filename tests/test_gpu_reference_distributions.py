"""GPU parity, part 5: output DISTRIBUTIONS against samples produced by the unmodified reference
(tests/golden/reference_distributions.npz, made by oracle/make_golden.py with the reference's own
NumPy RNG).  north_star: two-sample KS at p > 0.01 with independent RNG."""
import numpy as np
import pytest
from scipy import stats as sps

from rvmc_b200.bias_corr import BiasCorr
from rvmc_b200.rvmc import RVMC
from tests import gpu_util as g

pytestmark = pytest.mark.gpu

CASES = {"fbin070_pmin14": dict(fbin=0.7, min_period=1.4, max_period=3500.),
         "fbin028_pmin14": dict(fbin=0.28, min_period=1.4, max_period=3500.),
         "fbin070_pmin30": dict(fbin=0.7, min_period=30.0, max_period=3500.)}


@pytest.fixture(scope="module")
def ref():
    z, _ = g.golden("reference_distributions")
    return z


@pytest.mark.parametrize("tag", sorted(CASES))
def test_sigma1d_distributions_vs_reference(ref, tag):
    m17 = ref["m17_errors"]
    # infinite-pool fused kernel vs the reference's bootstrap from a 10^6-star pool
    c = RVMC(number_of_stars=1000, seed=101, **CASES[tag])
    fused = c.subsample_fused(sample_size=12, measured_errors=m17, number_of_samples=100000)
    ks = sps.ks_2samp(fused.astype(np.float64), ref["sigma_" + tag].astype(np.float64))
    assert ks.pvalue > 0.01, (tag, ks)
    fw = c.subsample_fused(sample_size=12, measured_errors=m17, number_of_samples=100000, weighted=True)
    ks = sps.ks_2samp(fw.astype(np.float64), ref["sigma_w_" + tag].astype(np.float64))
    assert ks.pvalue > 0.01, (tag, "weighted", ks)
    # finite-pool path (the reference's own semantics) and the per-orbit RV distribution
    c2 = RVMC(number_of_stars=10 ** 6, seed=202, **CASES[tag])
    rvb = c2.calculate_binary_velocities()
    ks = sps.ks_2samp(rvb[:200000], ref["rv_binary_" + tag].astype(np.float64))
    assert ks.pvalue > 0.01, (tag, "rv_binary", ks)
    pool = c2.subsample(sample_size=12, measured_errors=m17, number_of_samples=100000)
    ks = sps.ks_2samp(pool, ref["sigma_" + tag].astype(np.float64))
    assert ks.pvalue > 0.01, (tag, "pool", ks)
    assert abs(np.median(fused) / np.median(ref["sigma_" + tag]) - 1) < 0.02


def test_delta_rv_and_detection_vs_reference(ref):
    days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
    errs = [2.0, list(ref["biascorr_rv_errors_star1"])]
    bc = BiasCorr([days, days], number_of_samples=200000, rv_errors=errs, seed=303)
    bc.calculate_radial_velocities()
    det = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
    drv = bc.delta_rv
    for i in range(2):
        ks = sps.ks_2samp(drv[i].astype(np.float64), ref["biascorr_delta_rv"][i].astype(np.float64))
        assert ks.pvalue > 0.01, (i, ks)
        p_ref = float(ref["biascorr_detected_fraction"][i])
        se = np.sqrt(p_ref * (1 - p_ref) * (1 / 20000 + 1 / 200000))
        assert abs(det[i].mean() - p_ref) < 4 * se, (i, det[i].mean(), p_ref)


def test_grid_surface_and_best_fit_vs_reference_grid():
    """The reference's grid-fit semantics run with its own RVMC class (tests/golden/
    reference_grid_surface.npz) vs rvmc_grid_run: statistic surfaces agree within the reference's
    pool + Monte-Carlo noise, and the best-fit cell for an observed sigma lands within one cell
    (neighbouring cells differ by less than the reference's own noise)."""
    from rvmc_b200 import grid
    z, _ = g.golden("reference_grid_surface")
    periods, fbins = z["periods"], z["fbins"]
    F, P = np.meshgrid(fbins, periods, indexing="ij")
    pops = grid.points_array(F.size, fbin=F.ravel(), min_period=P.ravel(), max_period=3500.0, min_mass=6.0,
                             max_mass=20.0, sigma_dyn=2.0)
    res = grid.run_points(pops, grid.point_seeds(17, F.size), z["m17_errors"], 12, 100000, bin=0.5, ddof=0,
                          distributed=False)
    st = res["stats"].reshape(F.shape)
    for k, tol in (("median", 0.04), ("mean", 0.04), ("p16", 0.05), ("p84", 0.05), ("p05", 0.07), ("p95", 0.07)):
        rel = np.abs(st[k] / z[k] - 1)
        assert rel.max() < tol, (k, rel.max())
    assert np.median(np.abs(st["median"] / z["median"] - 1)) < 0.01
    # the mode of a 0.5-km/s histogram of 2e4 values is noisy where the distribution is broad
    # (~200 +- 14 counts per bin around the top): compare on the scale of the distribution's width
    assert np.all(np.abs(st["mode"] - z["mode"]) <= 1.0 + 0.25 * (z["p84"] - z["p16"]))
    # best fit along the P_min axis for each binary fraction, observed sigma = the reference's median
    for a in range(len(fbins)):
        ours = {k: list(st[k][a]) for k in ("mode", "median", "p05", "p16", "p84", "p95")}
        for j in (1, 3, 5):
            obs = float(z["median"][a, j])
            row = grid.find_best_Pmin(obs, 6, 20, 12, ours["mode"], ours["median"], ours["p05"], ours["p16"],
                                      ours["p84"], ours["p95"], None, None, list(periods), write=False)
            jj = int(np.argmin(np.abs(periods - row[0])))
            assert abs(jj - j) <= 1, (a, j, jj)
