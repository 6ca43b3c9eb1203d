"""Generate tests/golden/*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py            # rewrites tests/golden/

How the vectors are made: the reference modules are imported from /root/reference with the two
shims of oracle/shims/ on sys.path (astropy.constants.G = 6.6743e-11, empty matplotlib.pyplot).
NumPy's global generator is seeded and its four entry points the reference calls
(`uniform`, `random`, `normal`, `choice`) are wrapped by recorders that (a) forward to the real
generator so the reference sees exactly what it would see, and (b) log the *underlying* draw:
the U[0,1) variate behind `uniform(low, high)`, the returned array for `random`/`normal`, and
the index array behind `choice`.  Every array the reference then computes is stored next to
the draw log.  tests/test_oracle_golden.py replays the logged draws through
oracle/rvmc_oracle.py and compares.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RVMC_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

M17_ERRORS = np.array([0.8, 2.3, 1.0, 1.8, 0.8, 1.0, 1.4, 2.5, 2.0, 0.4, 1.5, 0.6])  # findBestDistributions.py:292


class DrawRecorder:
    """Wraps np.random.{uniform,random,normal,choice}; see module docstring."""

    def __init__(self):
        self.log = []
        self._orig = {}

    def __enter__(self):
        rs = np.random
        for name in ("uniform", "random", "normal", "choice"):
            self._orig[name] = getattr(rs, name)
        orig = self._orig

        def uniform(low=0.0, high=1.0, size=None):
            state = rs.get_state()
            out = orig["uniform"](low, high, size=size)
            after = rs.get_state()
            rs.set_state(state)
            u = rs.random_sample(size)
            assert np.array_equal(low + (high - low) * u, out), "uniform != low+(high-low)*u"
            rs.set_state(after)
            self.log.append(("uniform", np.asarray(u)))
            return out

        def random(size=None):
            out = orig["random"](size=size)
            self.log.append(("random", np.asarray(out)))
            return out

        def normal(loc=0.0, scale=1.0, size=None):
            out = orig["normal"](loc, scale, size=size)
            self.log.append(("normal", np.asarray(out)))
            return out

        def choice(a, size=None, replace=True, p=None):
            state = rs.get_state()
            out = orig["choice"](a, size=size, replace=replace, p=p)
            after = rs.get_state()
            rs.set_state(state)
            idx = orig["choice"](len(a), size=size, replace=replace, p=p)
            assert np.array_equal(np.asarray(a)[idx], out), "choice(a) != a[choice(len(a))]"
            rs.set_state(after)
            self.log.append(("choice", np.asarray(idx)))
            return out

        rs.uniform, rs.random, rs.normal, rs.choice = uniform, random, normal, choice
        return self

    def __exit__(self, *exc):
        for name, fn in self._orig.items():
            setattr(np.random, name, fn)

    def as_dict(self, prefix="draw"):
        return {"%s_%03d_%s" % (prefix, i, kind): arr for i, (kind, arr) in enumerate(self.log)}


def _import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import RVMC as ref_rvmc
        import RVMC_bias_corr as ref_bc
    assert os.path.dirname(os.path.abspath(ref_rvmc.__file__)) == os.path.abspath(REF)
    return ref_rvmc, ref_bc


def golden_rvmc(ref_rvmc, name, seed, n, ctor_kwargs, n_real, sample_size):
    np.random.seed(seed)
    with DrawRecorder() as rec:
        c = ref_rvmc.RVMC(number_of_stars=n, **ctor_kwargs)
        ctor_draws = len(rec.log)
        attrs = {k: np.array(getattr(c, k)) for k in
                 ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "eccentric_anomaly",
                  "inclination", "orbit_rotation", "cluster_velocities")}
        semi_major = c.calculate_semi_major_axis()
        v_max = c.max_orbital_velocity(semi_major)
        rv_orbit = c.binary_radial_velocity(v_max)              # noiseless orbital RV
        rv_binary = np.array(c.calculate_binary_velocities())   # + N(0, sigma_dyn), RVMC.py:224
        c.calculate_radial_velocity()
        radial_velocities = np.array(c.radial_velocities)
        mix_draws = len(rec.log)
        sig_unw = np.array(c.subsample(sample_size=sample_size, measured_errors=1.2,
                                       number_of_samples=n_real))
        unw_draws = len(rec.log)
        sig_w = np.array(c.subsample(sample_size=12, measured_errors=M17_ERRORS,
                                     number_of_samples=n_real, weighted=True))
    out = dict(attrs)
    out.update(rec.as_dict())
    out.update(semi_major_axis=semi_major, v_max=v_max, rv_orbit=rv_orbit, rv_binary=rv_binary,
               radial_velocities=radial_velocities, sigma_1d_unweighted=sig_unw,
               sigma_1d_weighted=sig_w, seed=seed, number_of_stars=n, n_real=n_real,
               sample_size=sample_size, ctor_draws=ctor_draws, mix_draws=mix_draws,
               unw_draws=unw_draws, m17_errors=M17_ERRORS)
    for k, v in ctor_kwargs.items():
        out["kw_" + k] = v
    out["kw_names"] = np.array(sorted(ctor_kwargs.keys()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "draws:", len(rec.log), "median sigma:", np.median(sig_unw))


def golden_biascorr(ref_bc, name, seed, nsamples, mode):
    np.random.seed(seed)
    rng = np.random.RandomState(seed + 1000)      # inputs, drawn outside the recorded stream
    if mode == "masses":
        nep = [5, 6, 4]                           # ragged epochs: the reference allows it
        t_obs = [np.sort(rng.uniform(0, np.pi * 10 ** 7, k)) for k in nep]
        masses = rng.uniform(10, 50, len(nep)) * 2e30
        kwargs = dict(masses=masses, mass_errors=masses * 0.1,
                      rv_errors=[list(rng.uniform(0.5, 3.0, k)) for k in nep])
    else:
        days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
        t_obs = [days.copy() for _ in range(4)]
        kwargs = dict(rv_errors=2.0, period_pi=-0.3, mass_ratio_kappa=-0.2, e_power=-0.4)
    with DrawRecorder() as rec:
        bc = ref_bc.BiasCorr(t_obs, number_of_samples=nsamples, **kwargs)
        ctor_draws = len(rec.log)
        attrs = {k: np.array(getattr(bc, k)) for k in
                 ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination",
                  "orbit_rotation")}
        bc.calculate_radial_velocities()
        rv_draws = len(rec.log)
        bc.generate_single_delta_rv()
        detected = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
        all_delta = bc.get_all_delta_rv()
    out = dict(attrs)
    out.update(rec.as_dict())
    for i, rv in enumerate(bc.radial_velocities):
        out["radial_velocities_%d" % i] = np.array(rv)
        out["t_obs_%d" % i] = np.array(t_obs[i])
        if mode == "masses":
            out["rv_errors_%d" % i] = np.array(kwargs["rv_errors"][i])
    out.update(delta_rv=np.array(bc.delta_rv), delta_rv_single=np.array(bc.delta_rv_single),
               detected=detected, all_delta_rv=all_delta, semi_major_axis=bc.semi_major_axis,
               max_orbital_velocity=bc.max_orbital_velocity, nstars=len(t_obs), nsamples=nsamples,
               ctor_draws=ctor_draws, rv_draws=rv_draws, seed=seed, fbin=bc.fbin)
    if mode == "masses":
        out.update(masses=kwargs["masses"], mass_errors=kwargs["mass_errors"])
    else:
        out.update(rv_errors_scalar=2.0, kw_period_pi=-0.3, kw_mass_ratio_kappa=-0.2, kw_e_power=-0.4)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "draws:", len(rec.log), "detected frac:", detected.mean())


def golden_distributions(ref_rvmc, ref_bc, name="reference_distributions"):
    """Samples of the reference's OUTPUT DISTRIBUTIONS (its own RNG, nothing recorded) for the
    two-sample KS legs of the parity protocol: sigma_1D of 12-star clusters bootstrapped from a
    10^6-star pool (three binary fractions / period cut-offs), and Delta-RV + detection flags of the
    6-epoch campaign.  float32 to keep the fixture small."""
    out = {}
    np.random.seed(20240607)
    cases = [("fbin070_pmin14", dict(fbin=0.7, min_period=1.4, max_period=3500.)),
             ("fbin028_pmin14", dict(fbin=0.28, min_period=1.4, max_period=3500.)),
             ("fbin070_pmin30", dict(fbin=0.7, min_period=30.0, max_period=3500.))]
    for tag, kw in cases:
        c = ref_rvmc.RVMC(number_of_stars=10 ** 6, **kw)
        c.calculate_binary_velocities()
        out["sigma_" + tag] = np.asarray(c.subsample(sample_size=12, measured_errors=M17_ERRORS,
                                                     number_of_samples=20000), dtype=np.float32)
        out["sigma_w_" + tag] = np.asarray(c.subsample(sample_size=12, measured_errors=M17_ERRORS,
                                                       number_of_samples=20000, weighted=True), dtype=np.float32)
        out["rv_binary_" + tag] = np.asarray(c.rv_binary[:50000], dtype=np.float32)
    days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
    bc = ref_bc.BiasCorr([days, days], number_of_samples=20000, rv_errors=[2.0, [0.5, 3.0, 1.0, 2.0, 0.7, 1.5]])
    bc.calculate_radial_velocities()
    det = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
    out["biascorr_delta_rv"] = np.asarray(bc.delta_rv, dtype=np.float32)
    out["biascorr_detected_fraction"] = det.mean(axis=1)
    out["biascorr_rv_errors_star1"] = np.array([0.5, 3.0, 1.0, 2.0, 0.7, 1.5])
    out["m17_errors"] = M17_ERRORS
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: (float(np.median(v)) if v.ndim else float(v)) for k, v in out.items() if k.startswith(("sigma_f", "biascorr_det"))})


def golden_grid_surface(ref_rvmc, name="reference_grid_surface"):
    """The reference's own grid-fit semantics (findBestDistributions.py:21-32,40,78,47-60) evaluated
    with the UNMODIFIED RVMC class standing in for the missing `main_RVMC.synthetic_RV_distribution`
    (SURVEY.md 8c): for each (P_min, f_bin) cell a 10^5-star pool, 2*10^4 bootstraps of 12 stars with
    the M17 errors, np.std with ddof=0, and the statistics of :47-60."""
    np.random.seed(424242)
    periods = np.logspace(np.log10(1.4), np.log10(3500), 8)
    fbins = np.linspace(0.1, 1.0, 6)
    keys = ("mode", "mean", "median", "p05", "p16", "p84", "p95")
    surf = {k: np.zeros((len(fbins), len(periods))) for k in keys}
    for a, fbin in enumerate(fbins):
        for b, pmin in enumerate(periods):
            c = ref_rvmc.RVMC(number_of_stars=10 ** 5, min_mass=6, max_mass=20, fbin=fbin, min_period=pmin,
                              max_period=3500, sigma_dyn=2.0)
            c.calculate_binary_velocities()
            c.calculate_radial_velocity()
            rv_dist = c.radial_velocities
            sig1D = np.std(np.random.choice(rv_dist, (20000, 12)) + np.random.normal(0, M17_ERRORS, (20000, 12)),
                           axis=1)                                                    # ddof = 0 (:29-30)
            bins = np.arange(0, max(sig1D) + 0.5, 0.5)
            values, bins = np.histogram(sig1D, bins=bins, density=True)
            index = int(np.where(values == np.max(values))[0][0])
            surf["mode"][a, b] = bins[index] + (bins[1] - bins[0]) / 2.
            surf["mean"][a, b] = np.mean(sig1D)
            surf["median"][a, b] = np.median(sig1D)
            for q, k in ((5, "p05"), (16, "p16"), (84, "p84"), (95, "p95")):
                surf[k][a, b] = np.percentile(sig1D, q)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), periods=periods, fbins=fbins, m17_errors=M17_ERRORS,
                        n_real=20000, **surf)
    print(name, "median surface corners:", surf["median"][0, 0], surf["median"][-1, 0], surf["median"][-1, -1])


def golden_multiseed(ref_rvmc, ref_bc, name="reference_multiseed", seeds=(11, 12, 13, 14, 15)):
    """the KS leg against the reference repeated over several INDEPENDENT reference
    runs (its own NumPy RNG, one seed each), so that the acceptance can be Bonferroni-aware instead of
    resting on one seed pair.  Per seed: sigma_1D of 10^4 clusters of 12 stars bootstrapped from a
    10^7-star pool (pool large enough that the reference's two-level sampling noise -- reference vs
    reference fails p > 0.01 one time in three at a 10^5 pool -- stays below the KS resolution at 10^4
    realizations), two populations; and Delta-RV + detected fraction of 10^4 binaries of the 6-epoch
    campaign.  float32, ~0.7 MB."""
    out = {"seeds": np.array(seeds), "m17_errors": M17_ERRORS}
    cases = [("fbin070_pmin14", dict(fbin=0.7, min_period=1.4, max_period=3500.)),
             ("fbin028_pmin14", dict(fbin=0.28, min_period=1.4, max_period=3500.))]
    days = np.array([0, 1, 7, 30, 180, 365.0]) * 86400
    for tag, kw in cases:
        rows = []
        for sd in seeds:
            np.random.seed(sd)
            c = ref_rvmc.RVMC(number_of_stars=10 ** 7, **kw)
            c.calculate_binary_velocities()
            rows.append(np.asarray(c.subsample(sample_size=12, measured_errors=M17_ERRORS, number_of_samples=10000),
                                   dtype=np.float32))
            del c
        out["sigma_" + tag] = np.stack(rows)
    drv, frac = [], []
    for sd in seeds:
        np.random.seed(1000 + sd)
        bc = ref_bc.BiasCorr([days], number_of_samples=10000, rv_errors=2.0)
        bc.calculate_radial_velocities()
        frac.append(bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4).mean())
        drv.append(np.asarray(bc.delta_rv[0], dtype=np.float32))
    out["biascorr_delta_rv"] = np.stack(drv)
    out["biascorr_detected_fraction"] = np.array(frac)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: np.round(np.median(v, axis=-1), 3).tolist() for k, v in out.items() if k.startswith(("sigma_", "biascorr_delta"))},
          out["biascorr_detected_fraction"])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_rvmc, ref_bc = _import_reference()
    import warnings
    warnings.simplefilter("ignore")
    if "--only-multiseed" in sys.argv:
        golden_multiseed(ref_rvmc, ref_bc)
        return
    golden_rvmc(ref_rvmc, "rvmc_default_n2048", seed=1234, n=2048, ctor_kwargs={}, n_real=96,
                sample_size=20)
    golden_rvmc(ref_rvmc, "rvmc_variant_n1536", seed=77, n=1536,
                ctor_kwargs=dict(min_mass=15., max_mass=60., min_period=1.4, max_period=3500.,
                                 period_pi=0.2, mass_ratio_kappa=-0.3, mass_power=-1.9, min_q=0.2,
                                 max_q=0.95, e_power=0.5, sigma_dyn=3.5, fbin=0.28),
                n_real=64, sample_size=12)
    golden_biascorr(ref_bc, "biascorr_masses_ragged", seed=5, nsamples=192, mode="masses")
    golden_biascorr(ref_bc, "biascorr_imf_6epochs", seed=9, nsamples=256, mode="imf")
    golden_distributions(ref_rvmc, ref_bc)
    golden_grid_surface(ref_rvmc)
    golden_multiseed(ref_rvmc, ref_bc)


if __name__ == "__main__":
    main()
