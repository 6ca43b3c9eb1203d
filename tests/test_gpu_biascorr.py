"""GPU parity, part 3: the multi-epoch Delta-RV / detection path (K5) through the C ABI."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats as sps

from oracle import rvmc_oracle as orc
from rvmc_b200 import _abi
from rvmc_b200 import _device as dv
from tests import gpu_util as g
from tests.oracle_from_draws import biascorr_item, biascorr_noise_from_draws, near_regime_edge, orbits_from_draws

pytestmark = pytest.mark.gpu

DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
ORBIT = ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination", "orbit_rotation")


def _tables(t_obs, rv_errors):
    nstars = len(t_obs)
    nep = np.array([len(t) for t in t_obs], dtype=np.int32)
    max_ep = int(nep.max())
    tob = np.zeros((nstars, max_ep))
    err = np.zeros((nstars, max_ep), dtype=np.float32)
    for i in range(nstars):
        tob[i, :nep[i]] = t_obs[i]
        if rv_errors is not None:
            err[i, :nep[i]] = rv_errors[i] if isinstance(rv_errors, (list, tuple)) else rv_errors
    return nep, max_ep, g.up(nep, "int32"), g.up(tob, "float64"), (g.up(err, "float32") if rv_errors is not None else None)


@pytest.mark.parametrize("name,mode", [("biascorr_masses_ragged", "masses"), ("biascorr_imf_6epochs", "imf")])
def test_replay_of_reference_arrays(name, mode):
    """The reference's exported BiasCorr arrays -> GPU multi-epoch RV (fp64) vs the reference's own
    radial_velocities / delta_rv / detected, noise replayed from the recorded normals."""
    z, d = g.golden(name)
    nstars, ns = int(z["nstars"]), int(z["nsamples"])
    t_obs = [z["t_obs_%d" % i] for i in range(nstars)]
    errs = [z["rv_errors_%d" % i] for i in range(nstars)] if mode == "masses" else [2.0] * nstars
    nep, max_ep, d_nep, d_tobs, d_err = _tables(t_obs, errs)
    tens = {k: g.up(z[k], "float64") for k in ORBIT}
    rv = g.empty((nstars, max_ep, ns), "float64")
    drv = g.empty((nstars, ns), "float64")
    g.ok(g.lib().rvmc_biascorr_replay_f64(nstars, ns, max_ep, dv.ptr(d_nep), dv.ptr(d_tobs), dv.soa(**tens), ns, None,
                                          0, 0, 1, dv.ptr(rv), dv.ptr(drv), g.stream()))
    rv = g.host(rv)
    first_noise = int(z["ctor_draws"])
    K = orc.semi_amplitude(z["mass_ratio"], z["primary_mass"], z["period"], z["eccentricity"], z["inclination"])
    noisy = np.zeros_like(rv)
    for i in range(nstars):
        want = z["radial_velocities_%d" % i] - d[first_noise + i]            # the reference's noiseless RVs
        assert g.rel_err(rv[i, :nep[i]], want, K[i]).max() < 1e-9
        noisy[i, :nep[i]] = z["radial_velocities_%d" % i]
    # get_detected_binaries on the reference's noisy RVs
    det = g.empty((nstars, ns), "uint8")
    g.ok(g.lib().rvmc_biascorr_detect(nstars, ns, max_ep, dv.ptr(d_nep), dv.ptr(g.up(noisy, "float64")), dv.ptr(d_err),
                                      20.0, 4.0, dv.ptr(det), g.stream()))
    np.testing.assert_array_equal(g.host(det).astype(bool), z["detected"])
    # delta_rv of the noiseless pass
    for i in range(nstars):
        want = z["radial_velocities_%d" % i] - d[first_noise + i]
        np.testing.assert_allclose(g.host(drv)[i], want.max(axis=0) - want.min(axis=0), rtol=1e-8, atol=1e-8)


def _fused(p, seed, noise_seed, t_obs, rv_errors, ns, first=0, mass_mode=0, masses=None, mass_errors=None, quirk=1,
           dRV=20.0, nsig=4.0, want_rv=True):
    nstars = len(t_obs)
    nep, max_ep, d_nep, d_tobs, d_err = _tables(t_obs, rv_errors)
    drv = g.empty((nstars, ns), "float32")
    det = g.empty((nstars, ns), "uint8")
    rv = g.empty((nstars, max_ep, ns), "float32") if want_rv else None
    cnt = g.zeros((4,), "int64")
    per = g.zeros((nstars,), "int32")
    m = None if masses is None else g.up(masses, "float64")
    me = None if mass_errors is None else g.up(mass_errors, "float64")
    g.ok(g.lib().rvmc_biascorr_fused(C.byref(p), seed, noise_seed, 0, nstars, max_ep, dv.ptr(d_nep), dv.ptr(d_tobs),
                                     dv.ptr(d_err), mass_mode, dv.ptr(m), dv.ptr(me), quirk, dRV, nsig, first, ns, ns,
                                     dv.ptr(drv), dv.ptr(det), dv.ptr(rv), dv.ptr(cnt), dv.ptr(per), g.stream()))
    return dict(drv=g.host(drv), det=g.host(det).astype(bool), rv=None if rv is None else g.host(rv), cnt=g.host(cnt),
                per=g.host(per), nep=nep, dRV=dRV)


@pytest.mark.parametrize("quirk", [1, 0])
@pytest.mark.parametrize("variant", [{}, dict(period_pi=-0.3, mass_ratio_kappa=-0.2, e_power=-0.4)])
def test_fused_multi_epoch_rv_matches_oracle_within_1e5(variant, quirk):
    """Per-orbit, per-epoch RVs of the fused fp32 path vs the oracle's fp64 restatement of
    RVMC_bias_corr.py:406-419 on the same Philox draws (noise off)."""
    p = _abi.PopParams.defaults(**variant)
    nstars, ns, seed, first = 3, 20000, 77, 1000
    rng = np.random.RandomState(1)
    t_obs = [DAYS6, np.sort(rng.uniform(0, np.pi * 1e7, 5)), np.sort(rng.uniform(0, np.pi * 1e7, 8))]
    res = _fused(p, seed, 0, t_obs, None, ns, first=first, quirk=quirk)
    for i in range(nstars):
        o = orbits_from_draws(p, seed, biascorr_item(i, np.arange(first, first + ns)))
        if quirk:
            want = orc.multi_epoch_rv(t_obs[i], o["time"], o["period"], o["eccentricity"], o["mass_ratio"],
                                      o["primary_mass"], o["inclination"], o["orbit_rotation"])
        else:
            tt = np.reshape(t_obs[i], (-1, 1)) + o["time"]
            M = 2 * np.pi * ((tt / o["period"]) % 1.0)
            E = orc.kepler_exact(M, o["eccentricity"])
            a = orc.semi_major_axis(o["mass_ratio"], o["primary_mass"], o["period"])
            vm = orc.max_orbital_velocity(o["mass_ratio"], o["primary_mass"], o["eccentricity"], a)
            want = orc.binary_radial_velocity(vm, o["eccentricity"], o["inclination"], o["orbit_rotation"], E)
        K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"], o["inclination"])
        keep = ~near_regime_edge(o, 1e-9)
        got = res["rv"][i, :len(t_obs[i])]
        assert g.rel_err(got, want, K)[:, keep].max() < 1e-5
        np.testing.assert_allclose(res["drv"][i][keep], (want.max(axis=0) - want.min(axis=0))[keep], rtol=0,
                                   atol=2e-5 * K[keep].max())
    assert not res["det"].any()                      # no errors given -> no detection flags
    assert res["cnt"][0] == ns * sum(len(t) for t in t_obs) and res["cnt"][3] == 0


@pytest.mark.parametrize("quirk", [1, 0])
@pytest.mark.parametrize("variant", [{}, dict(period_pi=-0.3, mass_ratio_kappa=-0.2, e_power=-0.4), dict(period_pi=0.37),
                                     dict(period_pi=-0.95), dict(period_pi=0.0)])
def test_compile_time_instantiation_matches_oracle_within_1e5(variant, quirk):
    """The call-free instantiation the headline runs (errors given, one Halley step) with the per-epoch
    RVs written out: vanishing errors make the noise term exactly negligible, so its orbital RVs face
    the oracle directly -- including a general period-law exponent, which this instantiation takes
    through pow_bounded_f64 (the phase needs the period to ~1e-9 relative here)."""
    p = _abi.PopParams.defaults(**variant)
    nstars, ns, seed, first = 2, 30000, 4242, 77
    rng = np.random.RandomState(5)
    t_obs = [DAYS6, np.sort(rng.uniform(0, np.pi * 1e7, 13))]          # 6 and 13 epochs: both unrolled variants
    res = _fused(p, seed, 9, t_obs, [1e-30, 1e-30], ns, first=first, quirk=quirk)
    for i in range(nstars):
        o = orbits_from_draws(p, seed, biascorr_item(i, np.arange(first, first + ns)))
        if quirk:
            want = orc.multi_epoch_rv(t_obs[i], o["time"], o["period"], o["eccentricity"], o["mass_ratio"],
                                      o["primary_mass"], o["inclination"], o["orbit_rotation"])
        else:
            tt = np.reshape(t_obs[i], (-1, 1)) + o["time"]
            E = orc.kepler_exact(2 * np.pi * ((tt / o["period"]) % 1.0), o["eccentricity"])
            a = orc.semi_major_axis(o["mass_ratio"], o["primary_mass"], o["period"])
            vm = orc.max_orbital_velocity(o["mass_ratio"], o["primary_mass"], o["eccentricity"], a)
            want = orc.binary_radial_velocity(vm, o["eccentricity"], o["inclination"], o["orbit_rotation"], E)
        K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"], o["inclination"])
        keep = ~near_regime_edge(o, 1e-9)
        assert g.rel_err(res["rv"][i, :len(t_obs[i])], want, K)[:, keep].max() < 1e-5
    assert res["cnt"][0] == ns * sum(len(t) for t in t_obs) and res["cnt"][2] == 0 and res["cnt"][3] == 0


@pytest.mark.parametrize("knobs", [dict(kepler_iters=2), dict(guard_amp=2.0), dict(kepler_iters=3, guard_amp=5.0)])
@pytest.mark.parametrize("noisy", [False, True])
def test_general_instantiation_halley_count_and_fp64_guard(knobs, noisy):
    """Configurations the compile-time instantiations do not take -- a run-time Halley count, or an fp64
    guard the host cannot prove unreachable (threshold lowered so that it engages at ordinary
    eccentricities) -- run the general instantiation of the same source: same parity bound."""
    p = _abi.PopParams.defaults(**knobs)
    nstars, ns, seed = 2, 20000, 31337
    t_obs = [DAYS6, DAYS6[:4] + 1234.5]
    res = _fused(p, seed, 5, t_obs, [1e-30, 1e-30] if noisy else None, ns)
    for i in range(nstars):
        o = orbits_from_draws(p, seed, biascorr_item(i, np.arange(ns)))
        want = orc.multi_epoch_rv(t_obs[i], o["time"], o["period"], o["eccentricity"], o["mass_ratio"],
                                  o["primary_mass"], o["inclination"], o["orbit_rotation"])
        K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"], o["inclination"])
        keep = ~near_regime_edge(o, 1e-9)
        assert g.rel_err(res["rv"][i, :len(t_obs[i])], want, K)[:, keep].max() < 1e-5
    assert res["cnt"][0] == ns * 10 and res["cnt"][3] == 0
    assert (res["cnt"][2] > 1000) == ("guard_amp" in knobs)          # guard lanes counted iff it can engage


@pytest.mark.parametrize("errors", ["scalar", "per_epoch"])
def test_fused_noise_and_detection_match_oracle(errors):
    p = _abi.PopParams.defaults()
    nstars, ns, seed, nseed = 2, 30000, 5, 6
    t_obs = [DAYS6, DAYS6[:4]]
    rv_errors = [2.0, 2.0] if errors == "scalar" else [np.array([0.5, 3.0, 1.0, 2.0, 0.7, 1.5]),
                                                      np.array([1.0, 1.0, 4.0, 0.3])]
    clean = _fused(p, seed, nseed, t_obs, None, ns)
    res = _fused(p, seed, nseed, t_obs, rv_errors, ns)
    total = 0
    for i in range(nstars):
        nep = len(t_obs[i])
        noise = biascorr_noise_from_draws(nseed, i, np.arange(ns), nep, rv_errors[i])
        np.testing.assert_allclose(res["rv"][i, :nep] - clean["rv"][i, :nep], noise, rtol=0, atol=3e-4)
        # detection rule (RVMC_bias_corr.py:459-474) applied by the oracle to the SAME noisy RVs
        want = orc.detected_binaries(res["rv"][i, :nep].astype(np.float64), rv_errors[i], 20, 4)
        diff = want != res["det"][i]
        if diff.any():          # only values within fp32 rounding of a threshold may differ
            assert diff.sum() <= 3
        np.testing.assert_allclose(res["drv"][i], res["rv"][i, :nep].max(axis=0) - res["rv"][i, :nep].min(axis=0),
                                   rtol=0, atol=1e-4)
        assert res["per"][i] == res["det"][i].sum()
        total += res["det"][i].sum()
    assert res["cnt"][1] == total
    frac = res["det"][0].mean()
    assert 0.3 < frac < 0.8          # the docstring-like setup detects roughly half


def test_fused_mass_modes_and_partition_invariance():
    p = _abi.PopParams.defaults()
    t_obs = [DAYS6, DAYS6]
    masses = np.array([20.0, 35.0]) * 2e30
    merr = masses * 0.1
    full = _fused(p, 9, 10, t_obs, [1.0, 2.0], 5000, mass_mode=1, masses=masses, mass_errors=merr)
    a = _fused(p, 9, 10, t_obs, [1.0, 2.0], 1234, first=0, mass_mode=1, masses=masses, mass_errors=merr)
    b = _fused(p, 9, 10, t_obs, [1.0, 2.0], 5000 - 1234, first=1234, mass_mode=1, masses=masses, mass_errors=merr)
    np.testing.assert_array_equal(np.concatenate([a["drv"], b["drv"]], axis=1), full["drv"])
    np.testing.assert_array_equal(np.concatenate([a["det"], b["det"]], axis=1), full["det"])
    # materialised attribute arrays use the same draws: replay them in fp64 and compare
    ns = 5000
    tens = {k: g.empty((2, ns), "float64") for k in ORBIT}
    g.ok(g.lib().rvmc_biascorr_sample(C.byref(p), 9, 2, 1, dv.ptr(g.up(masses, "float64")), dv.ptr(g.up(merr, "float64")),
                                      0, ns, 0, dv.soa(**tens), g.stream()))
    pm = g.host(tens["primary_mass"])
    assert abs(pm[0].mean() / masses[0] - 1) < 0.01 and abs(pm[1].std() / merr[1] - 1) < 0.05
    nep, max_ep, d_nep, d_tobs, d_err = _tables(t_obs, [1.0, 2.0])
    rv = g.empty((2, max_ep, ns), "float64")
    g.ok(g.lib().rvmc_biascorr_replay_f64(2, ns, max_ep, dv.ptr(d_nep), dv.ptr(d_tobs), dv.soa(**tens), ns,
                                          dv.ptr(d_err), 10, 0, 1, dv.ptr(rv), None, g.stream()))
    K = orc.semi_amplitude(*(g.host(tens[k]) for k in ("mass_ratio", "primary_mass", "period", "eccentricity",
                                                       "inclination")))
    assert g.rel_err(full["rv"], g.host(rv), np.maximum(K, 1.0)[:, None, :]).max() < 2e-5
    # face-value masses (mode 2): K scales as M^(1/3) between the two stars for equal draws... just finite
    f2 = _fused(p, 9, 10, t_obs, [1.0, 2.0], 1000, mass_mode=2, masses=masses)
    assert np.isfinite(f2["drv"]).all()
    # non-positive mass draws give NaN like the reference, and never a detection
    bad = _fused(p, 9, 10, t_obs, [1.0, 2.0], 2000, mass_mode=1, masses=masses, mass_errors=masses * 2.0)
    assert np.isnan(bad["drv"]).any() and not bad["det"][np.isnan(bad["drv"])].any()


def test_many_epochs_runtime_loop_variant():
    """> 16 epochs use the kernel variant with run-time loops; same answers as the oracle."""
    p = _abi.PopParams.defaults()
    rng = np.random.RandomState(3)
    t_obs = [np.sort(rng.uniform(0, 3e7, 40)), np.sort(rng.uniform(0, 3e7, 11))]
    errs = [rng.uniform(0.5, 3, 40), rng.uniform(0.5, 3, 11)]
    ns = 3000
    res = _fused(p, 21, 22, t_obs, errs, ns)
    for i in range(2):
        nep = len(t_obs[i])
        want = orc.detected_binaries(res["rv"][i, :nep].astype(np.float64), errs[i], 20, 4)
        assert (want != res["det"][i]).sum() <= 3
    o = orbits_from_draws(p, 21, biascorr_item(0, np.arange(ns)))
    clean = _fused(p, 21, 22, t_obs, None, ns)
    want = orc.multi_epoch_rv(t_obs[0], o["time"], o["period"], o["eccentricity"], o["mass_ratio"], o["primary_mass"],
                              o["inclination"], o["orbit_rotation"])
    K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"], o["inclination"])
    keep = ~near_regime_edge(o, 1e-9)
    assert g.rel_err(clean["rv"][0, :40], want, K)[:, keep].max() < 1e-5


def test_more_stars_than_one_grid_dimension():
    """> 65535 stars are launched in several star ranges; first_star shards give identical values."""
    p = _abi.PopParams.defaults()
    nstars, ns = 70001, 3
    t_obs = [DAYS6[:3]] * nstars
    full = _fused(p, 8, 9, t_obs, 2.0, ns, want_rv=False)
    assert np.isfinite(full["drv"]).all() and full["cnt"][0] == nstars * ns * 3
    assert full["per"].sum() == full["det"].sum() == full["cnt"][1]
    # the same stars in two calls with first_star offsets
    nep, max_ep, d_nep, d_tobs, d_err = _tables(t_obs[:1000], 2.0)
    parts = []
    for s0, s1 in ((0, 400), (400, 1000)):
        drv = g.empty((s1 - s0, ns), "float32")
        g.ok(g.lib().rvmc_biascorr_fused(C.byref(p), 8, 9, s0, s1 - s0, max_ep, dv.ptr(d_nep[s0:s1]), dv.ptr(d_tobs[s0:s1]),
                                         dv.ptr(d_err[s0:s1]), 0, None, None, 1, 20.0, 4.0, 0, ns, ns, dv.ptr(drv), None,
                                         None, None, None, g.stream()))
        parts.append(g.host(drv))
    np.testing.assert_array_equal(np.concatenate(parts), full["drv"][:1000])


def test_singles_delta_rv():
    nstars, n = 2, 40000
    errs = [np.array([1.0, 2.0, 0.5, 1.5, 1.0, 3.0]), np.array([2.0, 2.0, 2.0])]
    nep, max_ep, d_nep, _, d_err = _tables([DAYS6, DAYS6[:3]], errs)
    out = g.empty((nstars, n), "float64")
    g.ok(g.lib().rvmc_biascorr_singles(3, nstars, max_ep, dv.ptr(d_nep), dv.ptr(d_err), 0, n, dv.ptr(out), g.stream()))
    out = g.host(out)
    from oracle import philox_ref as ph
    for i in range(nstars):
        noise = biascorr_noise_from_draws(3, i, np.arange(n), nep[i], errs[i], stream=ph.STREAM_SINGLE)
        np.testing.assert_allclose(out[i], noise.max(axis=0) - noise.min(axis=0), rtol=0, atol=1e-10)
    # range of 3 iid N(0,2): mean = 2 * 1.6926
    assert abs(out[1].mean() / (2 * 1.6926) - 1) < 0.02


def test_detection_probability_ks_vs_oracle_independent_rng():
    """north_star: Delta-RV / detection distributions agree with the reference model under a
    two-sample KS test with independent RNG (oracle = NumPy MT19937 + SciPy-style secant)."""
    p = _abi.PopParams.defaults()
    ns = 40000
    res = _fused(p, 1234, 4321, [DAYS6], [2.0], ns, want_rv=False)
    rng = np.random.RandomState(99)
    m1 = orc.masses_from_uniform(rng.random_sample(ns), 6, 20, -2.35)
    per = orc.period_from_uniform(rng.random_sample(ns), 10 ** 0.15, 10 ** 3.5, -0.5)
    t = orc.time_from_uniform(rng.random_sample(ns), per)
    q = orc.mass_ratio_from_uniform(rng.random_sample(ns), 0.1, 1.0, 0)
    e = orc.eccentricity_from_uniform_slotted(per, rng.random_sample(ns), -0.5)
    inc = orc.inclination_from_uniform(rng.random_sample(ns))
    om = orc.orbit_rotation_from_uniform(rng.random_sample(ns))
    rv = orc.multi_epoch_rv(DAYS6, t, per, e, q, m1, inc, om) + rng.normal(0, 2.0, (6, ns))
    want_drv = orc.delta_rv(rv)
    want_det = orc.detected_binaries(rv, 2.0, 20, 4)
    ks = sps.ks_2samp(res["drv"][0].astype(np.float64), want_drv)
    assert ks.pvalue > 0.01, ks
    pd, po = res["det"][0].mean(), want_det.mean()
    se = np.sqrt(po * (1 - po) * 2 / ns)
    assert abs(pd - po) < 4 * se, (pd, po)


def test_biascorr_errors():
    p = _abi.PopParams.defaults()
    nep, max_ep, d_nep, d_tobs, d_err = _tables([DAYS6], [2.0])
    L = g.lib()
    drv = g.empty((1, 8), "float32")
    rc = L.rvmc_biascorr_fused(C.byref(p), 1, 1, 0, 1, 70, dv.ptr(d_nep), dv.ptr(d_tobs), dv.ptr(d_err), 0, None, None, 1,
                               20.0, 4.0, 0, 8, 8, dv.ptr(drv), None, None, None, None, g.stream())
    assert rc == _abi.RVMC_ERR_UNSUPPORTED
    rc = L.rvmc_biascorr_fused(C.byref(p), 1, 1, 0, 1, 6, dv.ptr(d_nep), dv.ptr(d_tobs), dv.ptr(d_err), 1, None, None, 1,
                               20.0, 4.0, 0, 8, 8, dv.ptr(drv), None, None, None, None, g.stream())
    assert rc == _abi.RVMC_ERR_INVALID and b"masses" in L.rvmc_last_error()
    g.ok(L.rvmc_biascorr_fused(C.byref(p), 1, 1, 0, 1, 6, dv.ptr(d_nep), dv.ptr(d_tobs), dv.ptr(d_err), 0, None, None, 1,
                               20.0, 4.0, 0, 0, 0, None, None, None, None, None, g.stream()))


@pytest.mark.parametrize("trial", range(6))
def test_random_campaigns_fused_vs_oracle(trial):
    """Random observing campaigns (1-12 ragged epochs, scalar / per-star / per-epoch errors, small and
    odd sample counts, both phase conventions): the fused kernel's noiseless RVs against the oracle's
    restatement, and its detection flags against the oracle's rule on the kernel's own noisy RVs."""
    rng = np.random.RandomState(100 + trial)
    nstars = int(rng.randint(1, 6))
    ns = int(rng.choice([1, 31, 257, 1000, 4099]))
    neps = rng.randint(1, 13, nstars)
    t_obs = [np.sort(rng.uniform(0, rng.choice([3e5, 3e7, 2e8]), k)) for k in neps]
    style = trial % 3
    if style == 0:
        errs = float(rng.uniform(0.5, 5))
    elif style == 1:
        errs = [float(rng.uniform(0.5, 5)) for _ in range(nstars)]
    else:
        errs = [rng.uniform(0.3, 6, k) for k in neps]
    quirk = int(trial % 2 == 0)
    kw = dict(period_pi=float(rng.choice([-0.5, 0.0, 0.3])), mass_ratio_kappa=float(rng.uniform(-0.9, 0.9)),
              e_power=float(rng.choice([-0.5, 0.7])), min_period=float(rng.choice([10 ** 0.15, 1.4, 20.0])))
    p = _abi.PopParams.defaults(**kw)
    seed, nseed, first = 500 + trial, 900 + trial, int(rng.randint(0, 10 ** 6))
    clean = _fused(p, seed, nseed, t_obs, None, ns, first=first, quirk=quirk)
    noisy = _fused(p, seed, nseed, t_obs, errs, ns, first=first, quirk=quirk, dRV=float(rng.choice([5.0, 20.0])), nsig=3.0)
    dRV = None
    for i in range(nstars):
        nep = len(t_obs[i])
        o = orbits_from_draws(p, seed, biascorr_item(i, np.arange(first, first + ns)))
        if quirk:
            want = orc.multi_epoch_rv(t_obs[i], o["time"], o["period"], o["eccentricity"], o["mass_ratio"],
                                      o["primary_mass"], o["inclination"], o["orbit_rotation"])
        else:
            tt = np.reshape(t_obs[i], (-1, 1)) + o["time"]
            E = orc.kepler_exact(2 * np.pi * ((tt / o["period"]) % 1.0), o["eccentricity"])
            a = orc.semi_major_axis(o["mass_ratio"], o["primary_mass"], o["period"])
            vm = orc.max_orbital_velocity(o["mass_ratio"], o["primary_mass"], o["eccentricity"], a)
            want = orc.binary_radial_velocity(vm, o["eccentricity"], o["inclination"], o["orbit_rotation"], E)
        K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"], o["inclination"])
        keep = ~near_regime_edge(o, 1e-9)
        # the quirk phase frac(2 pi t / P) is discontinuous at the wrap: skip the (measure-zero) lanes whose
        # phase lies within 1e-9 of it, where fp64 rounding may legitimately pick the other side
        ph = (2 * np.pi * (np.reshape(t_obs[i], (-1, 1)) + o["time"]) / o["period"]) % 1.0 if quirk else None
        if ph is not None:
            keep = keep & ~((ph < 1e-9) | (ph > 1 - 1e-9)).any(axis=0)
        assert g.rel_err(clean["rv"][i, :nep], want, K)[:, keep].max() < 1e-5, (trial, i)
        e_i = errs if style == 0 else errs[i]
        if nep >= 2:
            want_det = orc.detected_binaries(noisy["rv"][i, :nep].astype(np.float64), e_i, noisy["dRV"], 3.0)
        else:
            want_det = np.zeros(ns, dtype=bool)
        assert (want_det != noisy["det"][i]).sum() <= max(1, ns // 2000), (trial, i)
    assert noisy["cnt"][0] == ns * int(neps.sum()) and noisy["cnt"][3] == 0


@pytest.mark.parametrize("ns", [1000, 4099, 32 * 300])
def test_packed_detection_flags_equal_the_byte_flags(ns):
    """rvmc_biascorr_fused_packed: bit j of word w of a star's row is the uint8 flag of sample 32 w + j
    (ragged last warps included), whole pass and sample shards starting at multiples of 32; the host
    helper widens them to the bool array of RVMC_bias_corr.py:476 and counts them per star."""
    p = _abi.PopParams.defaults()
    rng = np.random.RandomState(4)
    t_obs = [DAYS6, np.sort(rng.uniform(0, np.pi * 1e7, 5)), DAYS6[:4]]
    errs = [2.0, [1.0, 2.0, 3.0, 1.5, 0.5], 1.2]
    nstars = len(t_obs)
    nep, max_ep, d_nep, d_tobs, d_err = _tables(t_obs, errs)
    nw = (ns + 31) // 32 + 1                                   # a stride wider than the row
    det = g.empty((nstars, ns), "uint8")
    drv = g.empty((nstars, ns), "float32")
    bits = g.zeros((nstars, nw), "int32")
    cnt = g.zeros((4,), "int64")

    def run(c0, c1, det_t, bits_t, drv_t):
        g.ok(g.lib().rvmc_biascorr_fused_packed(
            C.byref(p), 11, 12, 0, nstars, max_ep, dv.ptr(d_nep), dv.ptr(d_tobs), dv.ptr(d_err), 0, None, None, 1, 20.0,
            4.0, c0, c1 - c0, ns, None if drv_t is None else drv_t.data_ptr() + 4 * c0,
            None if det_t is None else det_t.data_ptr() + c0, bits_t.data_ptr() + 4 * (c0 // 32), nw, None,
            dv.ptr(cnt), None, g.stream()))

    run(0, ns, det, bits, drv)
    want = g.host(det).astype(bool)
    assert 0.05 < want.mean() < 0.95
    hb = g.host(bits).view(np.uint32)
    got = np.unpackbits(hb.view(np.uint8).reshape(nstars, -1), axis=1, bitorder="little")[:, :ns].astype(bool)
    np.testing.assert_array_equal(got, want)
    assert int(g.host(cnt)[1]) == int(want.sum())
    # bits past the end of a ragged row are zero (invalid lanes vote 0)
    full = np.unpackbits(hb.view(np.uint8).reshape(nstars, -1), axis=1, bitorder="little")
    assert not full[:, ns:32 * ((ns + 31) // 32)].any()
    # the library's host helper: widening and per-row counts
    out = np.full((nstars, ns), 9, dtype=np.uint8)
    hb = np.ascontiguousarray(hb)
    g.ok(g.lib().rvmc_unpack_bits_host(hb.ctypes.data, nw, nstars, ns, out.ctypes.data, ns, 3))
    np.testing.assert_array_equal(out.view(np.bool_), want)
    counts = np.zeros(nstars, dtype=np.int64)
    g.ok(g.lib().rvmc_count_bits_host(hb.ctypes.data, nw, nstars, ns, counts.ctypes.data, 2))
    np.testing.assert_array_equal(counts, want.sum(axis=1))
    # sample shards (the e2e path launches these on two streams): same bits, same Delta-RV
    bits2 = g.zeros((nstars, nw), "int32")
    drv2 = g.empty((nstars, ns), "float32")
    cuts = [0, 32 * 7, 32 * 20, ns]
    for a, b in zip(cuts[:-1], cuts[1:]):
        run(a, b, None, bits2, drv2)
    np.testing.assert_array_equal(g.host(bits2), g.host(bits))
    np.testing.assert_array_equal(g.host(drv2), g.host(drv))


@pytest.mark.parametrize("transfer", ["bits", "bytes"])
def test_biascorr_class_flag_transfer_modes_agree(transfer, monkeypatch):
    """get_detected_binaries() through the class: the bit-packed route (default) and the byte route give
    the same bool array; large enough that the pass is launched in overlapped ranges."""
    from rvmc_b200.bias_corr import BiasCorr
    res = {}
    for mode in ("bits", "bytes", "hybrid"):
        monkeypatch.setattr(BiasCorr, "flag_transfer", mode)
        for shape in ((40, 110_000), (2, 2_100_000)):
            bc = BiasCorr([DAYS6] * shape[0], number_of_samples=shape[1], rv_errors=2.0, seed=5)
            bc.calculate_radial_velocities()
            d = bc.get_detected_binaries()
            assert d.dtype == np.bool_ and d.shape == shape
            res[(mode, shape)] = d
            per_star, counters = bc.detection_counts()
            np.testing.assert_array_equal(per_star, d.sum(axis=1))
            d2 = bc.get_detected_binaries(delta_RV=30, n_sigma_RV=3)        # relaunch for another key
            assert d2.shape == shape and d2.sum() < d.sum()
    for shape in ((40, 110_000), (2, 2_100_000)):
        np.testing.assert_array_equal(res[("bits", shape)], res[("bytes", shape)])
        np.testing.assert_array_equal(res[("hybrid", shape)], res[("bytes", shape)])


def test_editing_rv_errors_after_the_pass_changes_only_the_thresholds():
    """RVMC_bias_corr.py:459-466 reads self.rv_errors when get_detected_binaries() runs: the stored RVs
    (noise from the OLD errors) are tested against thresholds from the NEW errors."""
    from rvmc_b200.bias_corr import BiasCorr
    bc = BiasCorr([DAYS6] * 3, number_of_samples=4000, rv_errors=2.0, seed=9)
    bc.calculate_radial_velocities()
    d_old = bc.get_detected_binaries()
    rv = [np.array(r) for r in bc.radial_velocities]
    bc2 = BiasCorr([DAYS6] * 3, number_of_samples=4000, rv_errors=2.0, seed=9)
    bc2.calculate_radial_velocities()
    bc2.rv_errors = 8.0
    d_new = bc2.get_detected_binaries()
    want = np.stack([orc.detected_binaries(rv[i], 8.0, 20, 4) for i in range(3)])
    assert (d_new != want).sum() <= 2                       # fp32 RVs at the threshold
    assert d_new.sum() < d_old.sum()
