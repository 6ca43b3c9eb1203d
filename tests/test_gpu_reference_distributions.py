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
