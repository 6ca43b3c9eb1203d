"""Drop-in for the reference's `RVMC` class (RVMC.py:23-328): same constructor, methods,
attributes, units and error behaviour; every method body is a call through the C ABI
(include/rvmc_b200.h) into sm_100a kernels.  There is no CPU path.

Differences a user can see (all additive):
  * `seed=` and `device=` keyword arguments.  Random numbers come from a counter-based Philox
    generator keyed by `seed` (OS entropy when omitted), not from NumPy's global MT19937 state,
    so `np.random.seed` has no effect here.
  * attribute arrays live on the GPU until they are read; reading one gives a NumPy float64 array
    (from then on the host copy is authoritative, so in-place edits are honoured).
  * `subsample_fused()` runs the infinite-pool limit of `subsample()` with per-star values kept
    on chip (the fast path; `subsample()` itself keeps the reference's finite-pool semantics).

One INTENTIONAL divergence: `subsample(weighted=True)` with a SCALAR `measured_errors`.  The reference
(RVMC.py:263-268) then divides by `np.sum(weights)` of a scalar -- the weight itself, not the sum over the
stars -- so its `rv_mean` is the SUM of the velocities and its result is sqrt(sum_j (v_j - sum_k v_k)^2), a
number its own docstring calls meaningless ("Only meaningful if an array of measured errors is provided",
RVMC.py:247).  Here a scalar error is what it means: equal weights, i.e. the ddof = 0 dispersion.  Pass the
errors as an array to get the reference's weighted formula bit for bit (tests/test_oracle_golden.py).
"""
import ctypes as C
import warnings

import numpy as np

from . import _abi
from . import _device as dv

STREAM_SIGMA_DYN = 2
STREAM_CLUSTER = 3

_POP_FROM_ATTR = dict(min_mass="min_mass", max_mass="max_mass", mass_power="mass_power",
                      min_period="min_period", max_period="max_period", period_pi="period_pi",
                      min_q="min_q", max_q="max_q", mass_ratio_kappa="mass_ratio_kappa",
                      e_power="e_power", sigma_dyn="sigma_dyn", fbin="fbin")


def _check_powers(**powers):
    """RVMC.py:112: `1. / (power + 1)` is a Python float division -> ZeroDivisionError."""
    for name, v in powers.items():
        if float(v) == -1.0:
            raise ZeroDivisionError("float division by zero")


class RVMC(dv.LazyArrays):
    """A tool for randomly generating radial velocity distributions of binary and single stars in
    clusters (GPU implementation of RVMC.py:23-328)."""

    _ARRAY_ATTRS = ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "eccentric_anomaly",
                    "inclination", "orbit_rotation", "rv_binary", "radial_velocities", "sigma_1d",
                    "cluster_velocities")
    _ORBIT_FIELDS = ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination",
                     "orbit_rotation", "eccentric_anomaly")

    def __init__(self, number_of_stars=10**5, min_mass=6., max_mass=20., min_period=10.**0.15,
                 max_period=10.**3.5, period_pi=-0.5, mass_ratio_kappa=0, mass_power=-2.35, min_q=0.1,
                 max_q=1.0, e_power=-0.5, mass_dist=None, period_dist=None, sigma_dyn=2.0, fbin=0.7,
                 seed=None, device=None, **physics):
        """`physics` (additive): the constants the reference hard-codes at RVMC.py:155-159
        -- e_min, e_max_mid, e_max_long, p_circ_days, p_mid_days -- and the solver knobs
        kepler_iters (Halley steps after the starter, default 1) and guard_amp."""
        unknown = set(physics) - {"e_min", "e_max_mid", "e_max_long", "p_circ_days", "p_mid_days", "kepler_iters",
                                  "guard_amp"}
        if unknown:
            raise TypeError("RVMC.__init__() got an unexpected keyword argument %r" % sorted(unknown)[0])
        self._physics = dict(physics)
        # RVMC.py:64-77
        self.fbin = fbin
        self.min_mass = min_mass
        self.max_mass = max_mass
        self.min_period = min_period
        self.max_period = max_period
        self.sigma_dyn = sigma_dyn
        self.number_of_stars = number_of_stars
        self.period_pi = period_pi
        self.mass_ratio_kappa = mass_ratio_kappa
        self.mass_power = mass_power
        self.min_q = min_q
        self.max_q = max_q
        self.e_power = e_power
        self.seed = dv.fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self._calls = 0
        self.nonconverged = 0

        n = int(number_of_stars)
        # the reference evaluates the samplers in this order and fails at the first bad index
        if mass_dist is None:
            _check_powers(mass_power=mass_power)
        if period_dist is None:
            _check_powers(period_pi=period_pi)
        _check_powers(mass_ratio_kappa=mass_ratio_kappa, e_power=e_power)
        for name, arr in (("mass_dist", mass_dist), ("period_dist", period_dist)):
            if arr is not None and np.shape(arr) != (n,):
                raise ValueError("%s has shape %s but number_of_stars is %d" % (name, np.shape(arr), n))

        self._lazy_init(dv.resolve_device(device))
        dev = self._device
        tens = {k: dv.empty((n,), "float64", dev) for k in self._ORBIT_FIELDS}
        given = 0
        if mass_dist is not None:
            tens["primary_mass"] = dv.to_device(mass_dist, "float64", dev)
            given |= 1
        if period_dist is not None:
            tens["period"] = dv.to_device(period_dist, "float64", dev)
            given |= 2
        pop = self._pop(lenient=given)
        nonconv = dv.zeros((1,), "int32", dev)
        cluster = dv.empty((n,), "float64", dev)
        seed0 = self._next_seed()
        with dv.DeviceGuard(dev):
            lib = _abi.lib()
            st = dv.stream_ptr(dev)
            _abi.check(lib.rvmc_sample_orbits(C.byref(pop), seed0, 0, n, given, dv.soa(**tens), dv.ptr(nonconv), st))
            # RVMC.py:101
            _abi.check(lib.rvmc_normal_f64(seed0, 0, n, STREAM_CLUSTER, 0, 0.0, float(sigma_dyn), 0,
                                           dv.ptr(cluster), st))
        for k, v in tens.items():
            self._set_device(k, v)
        self._set_device("cluster_velocities", cluster)
        self.rv_binary = None
        self._set_device("radial_velocities", dv.zeros((n,), "float64", dev))     # RVMC.py:98
        self.sigma_1d = None
        self._warn_nonconverged(nonconv)

    # ---- helpers --------------------------------------------------------------------------------
    def _next_seed(self):
        s = dv.call_seed(self.seed, self._calls)
        self._calls += 1
        return s

    def _pop(self, lenient=0, **overrides):
        """rvmc_pop_params from the CURRENT attribute values (the reference reads them at call time)."""
        kw = {k: float(getattr(self, a)) for k, a in _POP_FROM_ATTR.items()}
        kw.update(self._physics)
        kw.update(overrides)
        if lenient & 1:           # mass_dist given: the IMF law is unused, keep it well-formed
            kw.update(min_mass=6.0, max_mass=20.0, mass_power=-2.35)
        if lenient & 2:
            kw.update(min_period=10. ** 0.15, max_period=10. ** 3.5, period_pi=-0.5)
        return _abi.PopParams.defaults(**kw)

    def _warn_nonconverged(self, counter):
        bad = int(counter.item())
        self.nonconverged += bad
        if bad:
            # scipy.optimize.newton's message for array input (RVMC.py:177-180)
            warnings.warn("some failed to converge after 50 iterations (%d orbits)" % bad, RuntimeWarning)

    def _orbit_soa(self, names):
        tens = {k: self._on_device(k) for k in names}
        return tens, dv.soa(**tens)

    def _draw(self, field, given=0, size=None, pop=None):
        """Fresh draw of one attribute array (what each sample_* method of the reference returns)."""
        n = int(self.number_of_stars) if size is None else int(np.prod(size))
        dev = self._device
        out = dv.empty((n,), "float64", dev)
        tens = {field: out}
        if given & 2:
            tens["period"] = self._on_device("period")
            if tens["period"].numel() != n:
                raise ValueError("period has %d elements, expected %d" % (tens["period"].numel(), n))
        pop = self._pop() if pop is None else pop
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_sample_orbits(C.byref(pop), self._next_seed(), 0, n, given, dv.soa(**tens),
                                                     None, dv.stream_ptr(dev)))
        res = dv.to_host(out)
        return res if size is None or np.isscalar(size) else res.reshape(size)

    # ---- samplers (RVMC.py:103-168) -------------------------------------------------------------
    def sample_powerlaw(self, min_val, max_val, power, size=None):
        """Samples a powerlaw between min_val and max_val with index `power` (RVMC.py:103-112)."""
        _check_powers(power=power)
        pop = _abi.PopParams.defaults(min_q=float(min_val), max_q=float(max_val), mass_ratio_kappa=float(power))
        return self._draw("mass_ratio", size=size, pop=pop)

    def sample_inclination(self):
        return self._draw("inclination")

    def sample_masses(self):
        _check_powers(mass_power=self.mass_power)
        return self._draw("primary_mass")

    def sample_period(self):
        _check_powers(period_pi=self.period_pi)
        return self._draw("period")

    def sample_mass_ratio(self):
        _check_powers(mass_ratio_kappa=self.mass_ratio_kappa)
        return self._draw("mass_ratio")

    def sample_orbit_rotation(self):
        return self._draw("orbit_rotation")

    def sample_eccentricity(self):
        _check_powers(e_power=self.e_power)
        return self._draw("eccentricity", given=2, pop=self._pop(lenient=3))

    def sample_time(self):
        return self._draw("time", given=2, pop=self._pop(lenient=3))

    # ---- Kepler + RV (RVMC.py:170-226) ----------------------------------------------------------
    def calculate_eccentric_anomaly(self):
        """Solves Kepler's equation for the current time/period/eccentricity arrays (RVMC.py:170-180)."""
        dev = self._device
        tens, _ = self._orbit_soa(("time", "period", "eccentricity"))
        n = tens["time"].numel()
        out = dv.empty((n,), "float64", dev)
        nonconv = dv.zeros((1,), "int32", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_kepler_f64(n, dv.ptr(tens["time"]), dv.ptr(tens["period"]),
                                                  dv.ptr(tens["eccentricity"]), dv.ptr(out), dv.ptr(nonconv),
                                                  dv.stream_ptr(dev)))
        self._warn_nonconverged(nonconv)
        return dv.to_host(out)

    def calculate_semi_major_axis(self):
        """Total semi-major axis of all binaries in metres (RVMC.py:182-187)."""
        dev = self._device
        tens, s = self._orbit_soa(("mass_ratio", "primary_mass", "period"))
        n = tens["period"].numel()
        out = dv.empty((n,), "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_orbit_scales_f64(n, s, None, dv.ptr(out), None, dv.stream_ptr(dev)))
        return dv.to_host(out)

    def max_orbital_velocity(self, semi_major_axes):
        """Velocity of the primary at periastron in km/s (RVMC.py:189-198)."""
        dev = self._device
        tens, s = self._orbit_soa(("mass_ratio", "primary_mass", "period", "eccentricity"))
        n = tens["period"].numel()
        a = dv.to_device(semi_major_axes, "float64", dev)
        out = dv.empty((n,), "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_orbit_scales_f64(n, s, dv.ptr(a), None, dv.ptr(out), dv.stream_ptr(dev)))
        return dv.to_host(out)

    def binary_radial_velocity(self, v_max):
        """Radial velocity of the primary for a given periastron velocity array (RVMC.py:200-212)."""
        dev = self._device
        tens, s = self._orbit_soa(("eccentricity", "inclination", "orbit_rotation", "eccentric_anomaly"))
        n = tens["eccentricity"].numel()
        v = dv.to_device(v_max, "float64", dev)
        out = dv.empty((n,), "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_rv_from_vmax_f64(n, s, dv.ptr(v), dv.ptr(out), dv.stream_ptr(dev)))
        return dv.to_host(out)

    def calculate_binary_velocities(self, materialize=True):
        """Radial velocities of all binaries incl. the cluster dispersion (RVMC.py:214-226).
        `materialize=False` leaves the result on the device and returns None."""
        dev = self._device
        tens, s = self._orbit_soa(("primary_mass", "period", "mass_ratio", "eccentricity", "inclination",
                                   "orbit_rotation", "eccentric_anomaly"))
        n = tens["period"].numel()
        rv = dv.empty((n,), "float64", dev)
        with dv.DeviceGuard(dev):
            lib = _abi.lib()
            st = dv.stream_ptr(dev)
            _abi.check(lib.rvmc_rv_replay_f64(n, s, dv.ptr(rv), None, st))
            # RVMC.py:224
            _abi.check(lib.rvmc_normal_f64(self._next_seed(), 0, n, STREAM_SIGMA_DYN, 0, 0.0, float(self.sigma_dyn), 1,
                                           dv.ptr(rv), st))
        self._set_device("rv_binary", rv)
        return self.rv_binary if materialize else None

    # ---- pool mixing + bootstrap (RVMC.py:228-273) ----------------------------------------------
    def calculate_radial_velocity(self):
        """Combines single and binary stars into a pool with int(N*fbin) binaries (RVMC.py:228-236)."""
        if not self._has("rv_binary"):
            # np.random.choice(None, ...) in the reference
            raise ValueError("a must be 1-dimensional or an integer: call calculate_binary_velocities() first")
        dev = self._device
        rvb = self._on_device("rv_binary")
        cl = self._on_device("cluster_velocities")
        n = int(self.number_of_stars)
        if rvb.numel() != n or cl.numel() != n:
            raise ValueError("rv_binary / cluster_velocities must have number_of_stars = %d elements" % n)
        out = dv.empty((n,), "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_mix_pool(self._next_seed(), 0, n, float(self.fbin), dv.ptr(rvb), dv.ptr(cl),
                                                dv.ptr(out), dv.stream_ptr(dev)))
        self._set_device("radial_velocities", out)

    @staticmethod
    def _errors_vector(sample_size, measured_errors):
        """RVMC.py:251-254 and the broadcasting that follows at :264-271."""
        if not isinstance(measured_errors, float):
            if sample_size != len(measured_errors):          # TypeError for an int, like the reference (Q9)
                print("Changing the subsample size from %i to %i to match the measured errors!" %
                      (sample_size, len(measured_errors)))
                # ... but the reference does not change it: NumPy then fails to broadcast
                raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)"
                                 % (sample_size, len(measured_errors)))
            return np.ascontiguousarray(measured_errors, dtype=np.float64)
        return np.full(int(sample_size), measured_errors, dtype=np.float64)

    def subsample(self, sample_size=12, measured_errors=1.2, number_of_samples=10**5, fbin=None, weighted=False,
                  materialize=True, bessel=False):
        """Bootstraps `number_of_samples` clusters of `sample_size` stars from the pool and returns
        their velocity dispersions (RVMC.py:238-273; finite-pool semantics, ddof=1 or weighted).
        `bessel=True` (additive) applies the S/(S-1) correction the reference keeps commented
        out at RVMC.py:267 to the weighted dispersion."""
        err = self._errors_vector(sample_size, measured_errors)
        if weighted and isinstance(measured_errors, (list, tuple)):
            # `1 / measured_errors**2` (RVMC.py:263) needs an ndarray or a float
            raise TypeError("unsupported operand type(s) for ** or pow(): '%s' and 'int'"
                            % type(measured_errors).__name__)
        if fbin is not None:
            self.fbin = fbin                                  # persistent, like the reference (Q7)
        self.calculate_radial_velocity()
        dev = self._device
        pool = self._on_device("radial_velocities")
        ns = int(number_of_samples)
        out = dv.empty((ns,), "float64", dev)
        errd = dv.to_device(err, "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_sigma1d_pool(self._next_seed(), 0, dv.ptr(pool), pool.numel(), dv.ptr(errd),
                                                    int(sample_size), int(bool(weighted)),
                                                    (1 if bessel else 0) if weighted else 1, 0, ns, dv.ptr(out),
                                                    dv.stream_ptr(dev)))
        self._set_device("sigma_1d", out)
        return self.sigma_1d if materialize else None

    def subsample_fused(self, sample_size=12, measured_errors=1.2, number_of_samples=10**5, fbin=None,
                        weighted=False, ddof=None, first_realization=0, seed=None, return_counters=False,
                        materialize=True):
        """Infinite-pool limit of `subsample()`: every star of every realization is a fresh binary
        (probability fbin) or single; sampling, Kepler solve, RV and dispersion are fused in one
        kernel and per-star values never leave the chip.  Uses the CURRENT population parameters
        (not the materialised arrays).  Returns float32 dispersions."""
        err = self._errors_vector(sample_size, measured_errors)
        if ddof is None:
            ddof = 0 if weighted else 1           # RVMC.py:268 (no Bessel term) / :271 (ddof=1)
        if fbin is not None:
            self.fbin = fbin
        pop = self._pop()
        dev = self._device
        ns = int(number_of_samples)
        out = dv.empty((ns,), "float32", dev)
        errd = dv.to_device(err, "float32", dev)
        counters = dv.zeros((4,), "int64", dev)
        sd = self._next_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_sigma1d_fused(C.byref(pop), sd, dv.ptr(errd), int(sample_size),
                                                     int(bool(weighted)), int(ddof), int(first_realization), ns,
                                                     dv.ptr(out), dv.ptr(counters), dv.stream_ptr(dev)))
        self._set_device("sigma_1d", out)
        res = self.sigma_1d if materialize else None
        if return_counters:
            c = dv.to_host(counters)
            self.nonconverged += int(c[2])
            return res, dict(binary_rvs=int(c[0]), guard=int(c[1]), nonconverged=int(c[2]))
        return res

    # ---- plotting (out of scope of the GPU path; thin pass-through, matplotlib imported lazily) ---
    def check_initialization(self):
        """Cumulative histograms of the sampled parameters (RVMC.py:275-328)."""
        import matplotlib.pyplot as plt
        fig, axes = plt.subplots(3, 3, figsize=(6, 6))
        axes = axes.ravel()
        axes[7].remove()
        axes[8].remove()
        nbins = int(self.number_of_stars ** 0.333)
        panels = [(self.primary_mass / 2.e30, np.geomspace(self.min_mass, self.max_mass, nbins), "Primary mass", True),
                  (self.period / 86400., np.geomspace(self.min_period, self.max_period, nbins), "Period", True),
                  (self.mass_ratio, np.linspace(0, 1, nbins), "Mass ratio", False),
                  (self.eccentricity, np.linspace(0, 1, nbins), "Eccentricity", False),
                  (self.inclination, np.linspace(0, np.pi * 0.5, nbins), "Inclination", False),
                  (self.orbit_rotation, np.linspace(0, 2 * np.pi, nbins), "Orbit rotation", False),
                  (self.eccentric_anomaly, np.linspace(0, 2 * np.pi, nbins), "Eccentric anomaly", False)]
        for ax, (data, bins, title, logx) in zip(axes, panels):
            ax.hist(data, histtype="step", cumulative=True, density=True, bins=bins)
            ax.set_title(title)
            if logx:
                ax.set_xscale("log")
        plt.tight_layout()
        plt.show()


def plot_sigma_1d(sigma1d, labels, colors=None, hide_median=False):
    """Step histograms of several sigma_1D distributions on one axis, medians dashed (the figure of
    RVMC.py:331-365).  Plotting is outside the GPU path: this only hands the arrays to matplotlib and returns
    (fig, ax), or None after the reference's two complaints (more than ten curves without colours; label count)."""
    import matplotlib.pyplot as plt
    n = len(sigma1d)
    if colors is None and n > 10:
        print("Please specify your own colors if more than 10 distributions are to be plotted")
        return None
    if len(labels) != n:
        print("Warning! Number of labels does not match number of distributions")
        return None
    palette = list(colors) if colors is not None else plt.rcParams["axes.prop_cycle"].by_key()["color"]
    edges = np.linspace(0, 100, 200)
    fig, ax = plt.subplots(figsize=(6, 5))
    for dist, label, colour in zip(sigma1d, labels, palette):
        ax.hist(dist, bins=edges, density=True, histtype="step", lw=1.5, color=colour, label=label)
        if not hide_median:
            ax.axvline(np.median(dist), ls="--", color=colour, zorder=0.1)
    ax.set(xlabel=r"$\sigma_{\rm 1D}$ [km s$^{-1}$]", ylabel=r"Frequency [(km s$^{-1}$)$^{-1}$]", xlim=(0, 50))
    if n <= 10:
        ax.legend()
    fig.tight_layout()
    return fig, ax
