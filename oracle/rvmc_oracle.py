"""CPU oracle for the RVMC Monte-Carlo hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy (float64) restatement of the reference algorithm, function by function, each citing the
reference lines it follows (paths are relative to /root/reference).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import this module; `rvmc_b200/` never does.

Pinning status: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run
in the build container: `oracle/make_golden.py` imports the unmodified reference under the two
shims in `oracle/shims/`, with NumPy's global generator patched to *record* every draw, and
commits the recorded uniforms/normals/indices plus every array the reference computed from them
to `tests/golden/*.npz`.  `tests/test_oracle_golden.py` replays those draws through this module
and requires agreement to <= 1e-12 relative (bit-exact where the arithmetic is the same NumPy
expression).

Third-party arithmetic on the path whose source is not under /root/reference (no pins anywhere
in the reference -- it has no requirements file):
  * scipy.optimize.newton -> vectorised *secant* branch (no fprime is passed at RVMC.py:177-180 /
    RVMC_bias_corr.py:354-358).  Restated in `secant_array()` from SciPy's published algorithm
    (scipy 1.18.1 installed in the image, `scipy/optimize/_zeros_py.py::_array_newton`).
  * numpy.random legacy global generator (MT19937).  Not restated: every sampler below takes the
    uniform / normal / index draws as explicit arguments, in the reference's draw order.
  * astropy.constants.G -> pinned to 6.6743e-11 (CODATA 2018, astropy >= 4.0).
"""
import itertools

import numpy as np

G_VALUE = 6.6743e-11      # astropy.constants.G.value (RVMC.py:8,187,197)
MSUN_REF = 2.0e30         # the reference's solar mass in kg (RVMC.py:126)
DAY = 24 * 3600           # RVMC.py:134

# eccentricity regimes, hard-coded at RVMC.py:155-159
E_MIN = 1e-5
E_MAX_MID = 0.5
E_MAX_LONG = 0.9
P_CIRC_DAYS = 4
P_MID_DAYS = 6


# --------------------------------------------------------------------------------------------
# samplers (RVMC.py:103-168, twins RVMC_bias_corr.py:256-328); `u` are the U[0,1) draws
# --------------------------------------------------------------------------------------------
def powerlaw_from_uniform(u, min_val, max_val, power):
    """RVMC.py:103-112.  np.random.uniform(low, high) is low + (high-low)*random_sample()."""
    pmax = (max_val / min_val) ** (power + 1)
    p = pmax + (1 - pmax) * np.asarray(u, dtype=np.float64)
    return p ** (1. / (power + 1)) * min_val


def masses_from_uniform(u, min_mass, max_mass, mass_power):
    """RVMC.py:121-126: primary mass in kg."""
    return powerlaw_from_uniform(u, min_mass, max_mass, mass_power) * MSUN_REF


def period_from_uniform(u, min_period, max_period, period_pi):
    """RVMC.py:128-134: period in seconds; the power law is in log10(P/day)."""
    return DAY * 10 ** powerlaw_from_uniform(u, np.log10(min_period), np.log10(max_period), period_pi)


def time_from_uniform(u, period):
    """RVMC.py:163-168."""
    return np.asarray(u, dtype=np.float64) * period


def mass_ratio_from_uniform(u, min_q, max_q, kappa):
    """RVMC.py:136-140."""
    return powerlaw_from_uniform(u, min_q, max_q, kappa)


def eccentricity_regimes(period):
    """Masks of RVMC.py:156-157."""
    more_than_6 = period > (P_MID_DAYS * DAY)
    between_4_and_6 = (period > (P_CIRC_DAYS * DAY)) * np.logical_not(more_than_6)
    return more_than_6, between_4_and_6


def eccentricity_from_uniform_compact(period, u_long, u_mid, e_power):
    """RVMC.py:148-161 in the reference's draw order: first one uniform per P>6 d orbit, then one
    per 4 d<P<=6 d orbit (data-dependent draw counts)."""
    ecc = np.zeros(period.shape)
    m6, m46 = eccentricity_regimes(period)
    ecc[m6] = powerlaw_from_uniform(u_long, E_MIN, E_MAX_LONG, e_power)
    ecc[m46] = powerlaw_from_uniform(u_mid, E_MIN, E_MAX_MID, e_power)
    return ecc


def eccentricity_from_uniform_slotted(period, u, e_power):
    """Same map as RVMC.py:148-161 but with one uniform *slot* per orbit (the counter-based
    layout the GPU uses); the slot is simply unused when P <= 4 d."""
    u = np.asarray(u, dtype=np.float64)
    m6, m46 = eccentricity_regimes(period)
    ecc = np.zeros(period.shape)
    ecc[m6] = powerlaw_from_uniform(u[m6], E_MIN, E_MAX_LONG, e_power)
    ecc[m46] = powerlaw_from_uniform(u[m46], E_MIN, E_MAX_MID, e_power)
    return ecc


def inclination_from_uniform(u):
    """RVMC.py:114-119."""
    return np.arccos(np.asarray(u, dtype=np.float64))


def orbit_rotation_from_uniform(u):
    """RVMC.py:142-146."""
    return np.asarray(u, dtype=np.float64) * 2 * np.pi


# --------------------------------------------------------------------------------------------
# Kepler (RVMC.py:12-20,170-180; RVMC_bias_corr.py:11-19,348-358)
# --------------------------------------------------------------------------------------------
def anomaly_function(E, e, M):
    """RVMC.py:12-20."""
    return E - e * np.sin(E) - M


def secant_array(func, x0, args, tol, maxiter=50):
    """Restatement of SciPy's vectorised secant iteration (what `newton(func, x0, args=..., tol=...)`
    runs for array x0 without fprime).  Returns (root, converged, n_func_evals).

    Algorithm: second abscissa p1 = p(1+dx) +/- dx with dx = eps**0.33; repeat
    p_new = p1 - q1 (p1-p)/(q1-q0) on every lane with q1 != q0, midpoint on lanes whose secant
    slope vanishes for the first time; stop when *every* updated lane moved by < tol."""
    p = np.array(x0, dtype=np.float64, copy=True)
    dx = np.finfo(float).eps ** 0.33
    p1 = p * (1 + dx) + np.where(p >= 0, dx, -dx)
    q0 = np.asarray(func(p, *args))
    q1 = np.asarray(func(p1, *args))
    nev = 2
    failures = np.ones(p.shape, dtype=bool)
    active = np.ones(p.shape, dtype=bool)
    for _ in range(maxiter):
        nz = q1 != q0
        if not nz.any():
            p = (p1 + p) / 2.0
            break
        dp = (q1 * (p1 - p))[nz] / (q1 - q0)[nz]
        p[nz] = p1[nz] - dp
        first_flat = ~nz & active
        p[first_flat] = (p1 + p)[first_flat] / 2.0
        active &= nz
        failures[nz] = np.abs(dp) >= tol
        if not failures[nz].any():
            break
        p1, p = p, p1
        q0 = q1
        q1 = np.asarray(func(p1, *args))
        nev += 1
    return p, ~failures, nev


def eccentric_anomaly_single_epoch(time, period, eccentricity, tol=10 ** -6):
    """RVMC.py:170-180: M = 2 pi t / P, start at M."""
    M = np.pi * 2 * time / period
    root, _, _ = secant_array(anomaly_function, M, (eccentricity, M), tol)
    return root


def mean_anomaly_multi_epoch(time_deltas, period):
    """RVMC_bias_corr.py:357: `2 * np.pi * time_deltas % period / period` parses as
    ((2 pi t) % P) / P, a number in [0, 1) used *as radians* (SURVEY quirk Q1)."""
    return 2 * np.pi * time_deltas % period / period


def eccentric_anomaly_multi_epoch(time_deltas, period, eccentricity, tol=10 ** -6):
    """RVMC_bias_corr.py:348-358: start at the unwrapped 2 pi t / P, solve for the quirk-Q1 M."""
    x0 = np.pi * 2 * time_deltas / period
    M = mean_anomaly_multi_epoch(time_deltas, period)
    x0, e, M = np.broadcast_arrays(x0, eccentricity, M)
    root, _, _ = secant_array(anomaly_function, x0, (e, M), tol)
    return root


def kepler_exact(M, e, iters=60):
    """Independent check of the root (not reference code): bisection on the folded problem
    0 <= |M'| <= pi, then three Newton polishes; valid for every e < 1."""
    M = np.asarray(M, dtype=np.float64)
    e = np.broadcast_to(np.asarray(e, dtype=np.float64), M.shape)
    turns = np.floor(M / (2 * np.pi) + 0.5)
    Mr = M - turns * 2 * np.pi
    a = np.abs(Mr)
    lo = np.zeros_like(a)
    hi = np.full_like(a, np.pi)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        below = mid - e * np.sin(mid) - a < 0
        lo = np.where(below, mid, lo)
        hi = np.where(below, hi, mid)
    E = 0.5 * (lo + hi)
    for _ in range(3):
        E = E - (E - e * np.sin(E) - a) / (1 - e * np.cos(E))
    return np.where(Mr < 0, -E, E) + turns * 2 * np.pi


# --------------------------------------------------------------------------------------------
# orbit -> radial velocity (RVMC.py:182-226; RVMC_bias_corr.py:330-372)
# --------------------------------------------------------------------------------------------
def semi_major_axis(mass_ratio, primary_mass, period):
    """RVMC.py:182-187 (metres)."""
    return ((4 * np.pi ** 2) / (G_VALUE * (1 + mass_ratio) * primary_mass * period ** 2)) ** -(1. / 3.)


def max_orbital_velocity(mass_ratio, primary_mass, eccentricity, semi_major_axes):
    """RVMC.py:189-198 (km/s, primary at periastron)."""
    return (mass_ratio * ((G_VALUE * primary_mass * (1 + eccentricity)) /
                          ((1 + mass_ratio) * (1 - eccentricity) * semi_major_axes)) ** 0.5) * 0.001


def binary_radial_velocity(v_max, eccentricity, inclination, orbit_rotation, eccentric_anomaly):
    """RVMC.py:200-212."""
    return (1 / (1 + eccentricity)) * v_max * np.sin(inclination) * \
           (eccentricity * np.cos(orbit_rotation) +
            np.cos(2 * np.arctan((((1 + eccentricity) / (1 - eccentricity)) ** 0.5) *
                                 np.tan(0.5 * eccentric_anomaly)) + orbit_rotation))


def semi_amplitude(mass_ratio, primary_mass, period, eccentricity, inclination):
    """K = v_max/(1+e) sin i: the scale parity tolerances are quoted against."""
    a = semi_major_axis(mass_ratio, primary_mass, period)
    return max_orbital_velocity(mass_ratio, primary_mass, eccentricity, a) / (1 + eccentricity) * \
        np.sin(inclination)


def orbital_rv(primary_mass, period, mass_ratio, eccentricity, inclination, orbit_rotation,
               eccentric_anomaly):
    """RVMC.py:220-222 without the sigma_dyn term of :224."""
    a = semi_major_axis(mass_ratio, primary_mass, period)
    vmax = max_orbital_velocity(mass_ratio, primary_mass, eccentricity, a)
    return binary_radial_velocity(vmax, eccentricity, inclination, orbit_rotation, eccentric_anomaly)


# --------------------------------------------------------------------------------------------
# cluster mixing and dispersion (RVMC.py:228-273; findBestDistributions.py:29-30)
# --------------------------------------------------------------------------------------------
def mix_radial_velocities(rv_binary, cluster_velocities, fbin, binary_pick, single_pick):
    """RVMC.py:228-236 with the two `np.random.choice` results passed as index arrays:
    `binary_pick` = a without-replacement selection of int(N*fbin) indices, `single_pick` =
    N - int(N*fbin) indices drawn with replacement."""
    n = len(rv_binary)
    nb = int(n * fbin)
    out = np.zeros(n, dtype=float)
    out[:nb] = rv_binary[binary_pick]
    out[nb:] = cluster_velocities[single_pick]
    return out


def sigma1d_unweighted(rvs, ddof=1):
    """RVMC.py:270-271 (ddof=1) / findBestDistributions.py:29-30 (ddof=0, quirk Q10);
    `rvs` already holds choice(...) + normal(...), shape (number_of_samples, sample_size)."""
    return np.std(rvs, axis=1, ddof=ddof)


def sigma1d_weighted(rvs, measured_errors):
    """RVMC.py:262-268 (no Bessel term)."""
    weights = 1 / np.asarray(measured_errors, dtype=np.float64) ** 2
    rv_mean = np.sum(rvs * weights, axis=1) / np.sum(weights)
    return (np.sum(weights * (rvs.T - rv_mean).T ** 2, axis=1) / (np.sum(weights))) ** 0.5


# --------------------------------------------------------------------------------------------
# multi-epoch (RVMC_bias_corr.py:406-476)
# --------------------------------------------------------------------------------------------
def multi_epoch_rv(t_obs_i, time_i, period_i, ecc_i, mass_ratio_i, primary_mass_i, incl_i, omega_i):
    """One star of the loop at RVMC_bias_corr.py:414-419 -> (nepochs, nsamples) noiseless RVs."""
    time_deltas = np.reshape(t_obs_i, (-1, 1))
    virtual_times = time_deltas + time_i
    E = eccentric_anomaly_multi_epoch(virtual_times, period_i, ecc_i)
    a = semi_major_axis(mass_ratio_i, primary_mass_i, period_i)
    vmax = max_orbital_velocity(mass_ratio_i, primary_mass_i, ecc_i, a)
    return binary_radial_velocity(vmax, ecc_i, incl_i, omega_i, E)


def delta_rv(rvs):
    """RVMC_bias_corr.py:422."""
    return rvs.max(axis=0) - rvs.min(axis=0)


def needed_singles(number_of_samples, fbin):
    """RVMC_bias_corr.py:431."""
    return int((number_of_samples - fbin * number_of_samples) / fbin)


def detected_binaries(rvs, err, delta_RV=20, n_sigma_RV=4):
    """RVMC_bias_corr.py:459-474 for one star: `rvs` (nepochs, nsamples) noisy RVs, `err` a scalar
    or a length-nepochs sequence."""
    nep = len(rvs)
    if isinstance(err, (float, int)):
        err = [err] * nep
    else:
        err = np.ravel(err)
    det = np.zeros(rvs.shape[1], dtype=bool)
    for v1, v2 in itertools.combinations(range(nep), 2):
        d = np.abs(rvs[v1, :] - rvs[v2, :])
        det |= (d / np.sqrt(err[v1] ** 2 + err[v2] ** 2) > n_sigma_RV) & (d > delta_RV)
    return det


# --------------------------------------------------------------------------------------------
# grid statistics and best-fit selection (findBestDistributions.py:47-60,138-158)
# --------------------------------------------------------------------------------------------
def distribution_stats(sig1D, bin=0.5):
    """findBestDistributions.py:47-60.  `plt.hist(..., density=True)` is `np.histogram` with the
    same bins; mode = centre of the first maximal bin.  Returns a dict plus (values, bins)."""
    sig1D = np.asarray(sig1D, dtype=np.float64)
    bins = np.arange(0, max(sig1D) + bin, bin)
    values, bins = np.histogram(sig1D, bins=bins, density=True)
    index = int(np.where(values == np.max(values))[0][0])
    stats = dict(mode=bins[index] + (bins[1] - bins[0]) / 2.,
                 mean=np.mean(sig1D), median=np.median(sig1D),
                 p05=np.percentile(sig1D, 5), p16=np.percentile(sig1D, 16),
                 p84=np.percentile(sig1D, 84), p95=np.percentile(sig1D, 95))
    return stats, values, bins


def percentile_linear(sorted_vals, q):
    """NumPy's default ('linear') percentile written out, so the GPU select can be checked against
    the same interpolation rule: virtual index q/100*(n-1), lerp between the two neighbours."""
    a = np.asarray(sorted_vals, dtype=np.float64)
    n = a.size
    pos = (q / 100.0) * (n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    t = pos - lo
    lo_v, hi_v = a[lo], a[hi]
    diff = hi_v - lo_v
    out = lo_v + diff * t
    if t >= 0.5:                      # numpy's _lerp evaluates from the other end for t >= 0.5
        out = hi_v - diff * (1 - t)
    return out


def best_fit_indices(observed_sigma, mode, median, p05, p16, p84, p95, usemode=False):
    """findBestDistributions.py:129-158: first strict minimum of |stat - observed| (initial
    distance 100, index 0) for the central statistic and the four percentile curves."""
    central = mode if usemode else median
    out = []
    for curve in (central, p05, p16, p84, p95):
        best, idx = 100, 0
        for i, v in enumerate(curve):
            d = np.abs(v - observed_sigma)
            if d < best:
                best, idx = d, i
        out.append(idx)
    return tuple(out)     # (index, index05, index16, index84, index95)
